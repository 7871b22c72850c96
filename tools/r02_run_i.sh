#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$|Ignoring|CUTOFF|Training stopped|EvaluatorHoldout|WARNING" | tail -60 > gpurun_out/r02_pytest_gpu_3.log
tail -45 gpurun_out/r02_pytest_gpu_3.log
python tools/quality_sweep.py --seeds 1337,1,2 > gpurun_out/r02_quality_sweep_v2.json 2> gpurun_out/r02_quality_sweep_v2.log
cat gpurun_out/r02_quality_sweep_v2.log
GANMF_BENCH_GEMM_TABLE=gpurun_out/r02_gemm_table_v2 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/r02_bench_n1_v2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_n1_v2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1_v2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["eval"], d["roofline"]["frac"], d["cpu_baseline"]["value"])
r = d["records"][0]
print({k: r[k] for k in ("value", "ms_per_step")}, r["e2e"]["value"], r["eval"], r["roofline"]["frac"], r["cpu_baseline"]["value"])
PY
