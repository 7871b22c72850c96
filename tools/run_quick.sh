#!/bin/bash
# quick visit: kernel + eval tests, then the bench (no CPU baseline)
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh "$@" > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -20
GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table.txt timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])
print("value %.0f rows/s  ms/step %.3f  gemm %.1f TF/s (share %.2f)  e2e %.0f" % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['e2e']['value']))
print("eval %.0f users/s  hbm_frac %.3f" % (d['eval']['value'], d['eval']['hbm_frac_4I_bytes_per_user']))
for k, v in d['hbm_kernels'].items():
    print("  %-28s %.0f GB/s  frac %.3f" % (k, v['achieved'], v['frac']))
print("clocks", d['clocks'])
PY
