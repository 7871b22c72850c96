#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_tp.py -m gpu -q -x --timeout 800 -k "sparse_real_route or resident_generator or (item_sharded_steps_parity and 111) or cta_pairs" > gpurun_out/r02b_initcheck_steps.log 2>&1
echo "initcheck rc=$?"; grep -c "Uninitialized" gpurun_out/r02b_initcheck_steps.log; grep -A12 "Uninitialized" gpurun_out/r02b_initcheck_steps.log | head -60; tail -4 gpurun_out/r02b_initcheck_steps.log
