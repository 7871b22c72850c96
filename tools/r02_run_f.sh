#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -x -q -s 2>&1 | grep -E "cfg3-item|free run|passed|failed|Error" | head -30
python tools/quality_sweep.py --seeds 1337,1,2,3,4,5,6,7,8,9 --runs DisGANMF_user_LastFM,DisGANMF_item_LastFM > gpurun_out/r02_quality_sweep_lastfm10.json 2> gpurun_out/r02_quality_sweep_lastfm10.log
cat gpurun_out/r02_quality_sweep_lastfm10.log
for gp in 1 2 3; do
  GANMF_GEMM_PATH=$gp python tools/quality_sweep.py --seeds 1337,1 --runs DisGANMF_item_1M > gpurun_out/r02_quality_item1M_path$gp.json 2> gpurun_out/r02_quality_item1M_path$gp.log
  echo "gemm path $gp"; cat gpurun_out/r02_quality_item1M_path$gp.log
done
for blk in 138000 65536 32768; do python tools/eval_profile.py --items 27000 --users 138000 --density 0.005 --block $blk; done
