#!/bin/bash
# what the driver runs at round end (GPU tests, smoke, bench) + ncu --set full of the HBM-bound kernels
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.log | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table_final.txt timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-users 2048"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_adam|csr_gather|p_catchup|p_batch_adam|colsum|splitk_reduce" -s 4000 -c 8 -o gpurun_out/prof_hbm_v8 $CMD > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"topk_rows|mask_seen|user_metrics" -c 6 -o gpurun_out/prof_eval_v8 $CMD > gpurun_out/ncu_eval.log 2>&1
echo "eval capture rc=$?"
for f in prof_hbm_v8 prof_eval_v8; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
print("value %.0f rows/s  ms/step %.3f  gemm %.1f TF/s (share %.2f)  e2e %.0f  launches %d" % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['e2e']['value'], d['gpu_launches']))
print("eval %.0f users/s" % d['eval']['value']); print("cpu", d.get('cpu_baseline', {}).get('value')); print("clocks", d['clocks'])
PY
