#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_parity.py -q -m gpu --timeout 600 > gpurun_out/train_parity.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/train_parity.log
