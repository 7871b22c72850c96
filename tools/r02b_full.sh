#!/bin/bash
# full GPU test suite + smoke + the default bench line (N=1) + the reference arm + (optional) extra quick A/Bs
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02b_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02b_smoke.log
GANMF_BENCH_GEMM_TABLE=gpurun_out/r02b_gemm_table_n1 timeout 1200 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02b_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["eval"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("achieved_algorithmic"), d["clocks"])
print(d.get("cpu_baseline"))
for r in d.get("records", []):
    print({k: r.get(k) for k in ("metric", "value", "ms_per_step", "ms_per_step_pair")}, r.get("eval", {}).get("value"))
print(d.get("hbm_kernels"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02b_bench_reference_arm.json 2> gpurun_out/r02b_bench_reference_arm.err; echo "reference arm rc=$?"; head -c 700 gpurun_out/r02b_bench_reference_arm.json; echo
for v in "GANMF_SPARSE_REAL=1" "GANMF_SPARSE_REAL=0"; do
  env $v timeout 300 python bench.py --quick --workload cfg4 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 $v', d['value'], d['ms_per_step'], d['roofline']['routes'])"
done
