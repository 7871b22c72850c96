#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
SKIP_TESTS=1 tools/r02b_ab.sh "-" 2>&1 | grep -v "^ \|^#"
for b in 0 65536 32768 16384 8192; do
  timeout 300 python tools/eval_profile.py --items 27000 --users 131072 --density 0.005 --block $b 2>&1 | tail -1
done
for b in 0 65536 32768; do
  timeout 300 python tools/eval_profile.py --items 200000 --users 131072 --density 0.001 --block $b 2>&1 | tail -1
done
