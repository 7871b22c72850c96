#!/bin/bash
# round 2 (second session): route tests, then same-box A/B at cfg5 (bench.py --quick: training throughput + GEMM table).
# usage: r02b_ab.sh "<VAR=val ...>;<VAR=val ...>;..."   (one bench run per ';'-separated environment, '-' = defaults)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_parity.py tests/test_gpu_tp.py tests/test_gpu_baseline_shapes.py tests/test_gpu_api.py tests/test_gpu_dist_api.py -m gpu -q -x --timeout 600 > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02b_tests.log
fi
IFS=';' read -ra VARIANTS <<< "${1:--}"
i=0
for v in "${VARIANTS[@]}"; do
  i=$((i+1)); tag="v$i"
  envs=""; [ "$v" != "-" ] && envs="$v"
  env $envs GANMF_BENCH_GEMM_TABLE=gpurun_out/r02b_gemm_$tag timeout 600 python bench.py --quick --steps 20 --warmup 5 \
     > gpurun_out/r02b_quick_$tag.json 2> gpurun_out/r02b_quick_$tag.err
  echo "[$tag] env: $envs rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02b_quick_$tag.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "gemm TF/s %.1f" % d["roofline"]["achieved"], "share %.3f" % d["roofline"]["gemm_share_of_step"], d["roofline"]["routes"], d["clocks"]["sm_mhz"], d["loss_last"])
except Exception as e:
    print("no line", e)
PY
  tail -c 400 gpurun_out/r02b_quick_$tag.err
  cat gpurun_out/r02b_gemm_$tag.cfg5
done
