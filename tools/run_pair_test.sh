#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "pairs" --timeout 300 > gpurun_out/pair_test.log 2>&1
echo "pair test rc=$?"; tail -15 gpurun_out/pair_test.log
