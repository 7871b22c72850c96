#!/bin/bash
# round 2: fused evaluator -- new tests, then the bench line (eval section) and the whole GPU suite
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_eval.py -x -q 2>&1 | tail -40 > gpurun_out/r02_pytest_fused.log
cat gpurun_out/r02_pytest_fused.log
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_2.log
cat gpurun_out/r02_pytest_gpu_2.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-hbm-kernels > gpurun_out/r02_bench_n1_eval.json 2> gpurun_out/r02_bench_n1_eval.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench_n1_eval.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1_eval.json").read().strip().splitlines()[-1])
print(d["value"], d["eval"])
print(d["records"][0]["value"], d["records"][0]["eval"])
PY
