#!/bin/bash
# round 2 evidence: launch list of the timed steps (cfg5, N=1), full-set capture of the 13 GEMM launches of a step pair
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
GANMF_BENCH_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file gpurun_out/r02_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r02_launches_cfg5.out 2>&1
echo "launch list rc=$?"
GANMF_BENCH_PROFILER_RANGE=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tc_gemm -c 13 \
  -o gpurun_out/r02_tc_gemm_cfg5 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r02_tc_gemm_cfg5.out 2>&1
echo "full set rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_cfg5.csv
