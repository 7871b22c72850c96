#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python tools/diag_cfg3_item.py GEMM_AUTO 2>/dev/null | head -14
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_train_parity.py -x -q -s 2>&1 | grep -E "cfg3-item|free run|passed|failed|Error|assert" | head -40
python tools/quality_sweep.py --seeds 1337,1,2 --runs DisGANMF_item_hetrec2011 2>&1 >/dev/null | tail -3
