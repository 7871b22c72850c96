#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests (fused)"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -12
for nf in 1 0; do
  GANMF_NO_FUSED_ADAM=$nf GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table_nf$nf.txt timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nf$nf.log 2> gpurun_out/bench_nf$nf.err
  echo "== GANMF_NO_FUSED_ADAM=$nf"; tail -2 gpurun_out/bench_nf$nf.err
  python - $nf <<'PY'
import json, sys
d = json.loads([l for l in open('gpurun_out/bench_nf%s.log' % sys.argv[1]) if l.startswith('{')][-1])
print("value %.0f rows/s  ms/step %.3f  gemm %.1f TF/s (share %.2f)  launches %d loss %s" % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['gpu_launches'], d['loss_last']))
PY
  cat gpurun_out/gemm_table_nf$nf.txt
done
cp gpurun_out/bench_nf0.log gpurun_out/bench.log
