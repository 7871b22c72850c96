#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for blk in 131072 65536 32768 16384; do
  python tools/eval_profile.py --items 200000 --users 131072 --block $blk
done
GANMF_EVAL_SEGS=1 python tools/eval_profile.py --items 200000 --users 131072 --block 32768
GANMF_EVAL_SEGS=2 python tools/eval_profile.py --items 200000 --users 131072 --block 32768
GANMF_EVAL_SEGS=1 python tools/eval_profile.py --items 200000 --users 32768 --block 32768
GANMF_EVAL_SEGS=8 python tools/eval_profile.py --items 200000 --users 32768 --block 32768
python tools/eval_profile.py --items 200000 --users 32768 --block 32768
EVAL_PROFILE_RANGE=1 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:score_select -c 1 \
  -o gpurun_out/r02_score_select_full python tools/eval_profile.py --items 200000 --users 32768 --block 32768 --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
