#!/bin/bash
# ncu --set full of the HBM-bound kernels inside the timed steps (cudaProfilerStart/Stop range of bench.py)
mkdir -p gpurun_out
GANMF_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"fused_adam|csr_gather|p_catchup|p_batch_adam|colsum|rowdot|splitk_reduce|gather_rows|sqdiff" -c 24 \
  -o gpurun_out/prof_hbm_v8 python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"; tail -2 gpurun_out/ncu_hbm.log
ncu -i gpurun_out/prof_hbm_v8.ncu-rep --page raw --csv > gpurun_out/prof_hbm_v8_raw.csv 2>/dev/null
ls -la gpurun_out/prof_hbm_v8*
