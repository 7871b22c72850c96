#!/bin/bash
# round 2 (second session): new-route tests, then same-box A/B of the sparse real-profile encode and of the decoder-bias
# gradient from the residual GEMM's column sums at cfg5 (bench.py --quick: training throughput + GEMM table only)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_tp.py tests/test_gpu_baseline_shapes.py tests/test_gpu_api.py -m gpu -q -x --timeout 600 > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/r02b_tests.log
for v in "auto:1" "0:1" "auto:0"; do
  sr=${v%%:*}; cp=${v##*:}
  if [ "$sr" = auto ]; then unset GANMF_SPARSE_REAL; else export GANMF_SPARSE_REAL=$sr; fi
  GANMF_COLPART=$cp GANMF_BENCH_GEMM_TABLE=gpurun_out/r02b_gemm_sr${sr}_cp${cp} timeout 600 python bench.py --quick --steps 20 --warmup 5 \
     > gpurun_out/r02b_quick_sr${sr}_cp${cp}.json 2> gpurun_out/r02b_quick_sr${sr}_cp${cp}.err
  echo "sparse=$sr colpart=$cp rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02b_quick_sr${sr}_cp${cp}.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["achieved"], d["roofline"]["routes"], d["clocks"]["sm_mhz"], d["loss_last"])
except Exception as e:
    print("no line", e)
PY
  tail -c 600 gpurun_out/r02b_quick_sr${sr}_cp${cp}.err
done
cat gpurun_out/r02b_gemm_srauto_cp1.cfg5
