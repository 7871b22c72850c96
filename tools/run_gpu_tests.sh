#!/bin/bash
# usage: tools/run_gpu_tests.sh [pytest args]; log lands in gpurun_out/pytest_gpu.log
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used --format=csv > gpurun_out/pytest_gpu.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 "$@" >> gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
