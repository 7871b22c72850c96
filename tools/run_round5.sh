#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -20
GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table.txt timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
echo "==== bench"; cat gpurun_out/gemm_table.txt; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
bash tools/run_ncu_full.sh
