#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
GANMF_BENCH_GEMM_TABLE=gpurun_out/r02b_gemm_table_n1 timeout 1200 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/r02b_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["eval"]["value"], d["eval"]["hbm_frac_4I_bytes_per_user"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"].get("achieved_algorithmic"), d["clocks"])
for r in d.get("records", []):
    print({k: r.get(k) for k in ("metric", "value", "ms_per_step", "ms_per_step_pair", "unavailable")}, (r.get("eval") or {}).get("value"))
PY
