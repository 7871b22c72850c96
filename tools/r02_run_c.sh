#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_eval.py tests/test_gpu_baseline_shapes.py tests/test_gpu_dist_api.py tests/test_gpu_eval.py -x -q -s 2>&1 | tail -25
for shape in "200000 32768 0.001" "27000 32768 0.005" "200000 131072 0.001" "200000 250000 0.001" "27000 138000 0.005"; do
  set -- $shape
  python tools/eval_profile.py --items $1 --users $2 --density $3
done
python tools/quality_sweep.py --seeds 1337,1,2 > gpurun_out/r02_quality_sweep.json 2> gpurun_out/r02_quality_sweep.log
cat gpurun_out/r02_quality_sweep.log
