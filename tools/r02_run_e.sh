#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_eval.py tests/test_gpu_baseline_shapes.py tests/test_gpu_dist_api.py tests/test_gpu_eval.py -x -q -s 2>&1 | tail -30
for blk in 131072 65536; do
  python tools/eval_profile.py --items 200000 --users 131072 --block $blk
done
python tools/eval_profile.py --items 200000 --users 32768 --block 32768
python tools/eval_profile.py --items 200000 --users 32768
python tools/eval_profile.py --items 27000 --users 138000 --density 0.005 --block 138000
EVAL_PROFILE_RANGE=1 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:score_select -c 1 \
  -o gpurun_out/r02_score_select_full2 python tools/eval_profile.py --items 200000 --users 32768 --block 32768 --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
