#!/bin/bash
# round 2, first GPU call: GPU tests (incl. the item-sharded step on one device) + the new default bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu_1.log
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1.log
cat gpurun_out/r02_pytest_gpu_1.log
GANMF_BENCH_GEMM_TABLE=gpurun_out/r02_gemm_table timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_first.json 2> gpurun_out/r02_bench_n1_first.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r02_bench_n1_first.err
head -c 6000 gpurun_out/r02_bench_n1_first.json
