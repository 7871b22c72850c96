#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gemm_selftest2.sh gpurun_out/gemm_selftest3.log > /dev/null 2>&1
echo "==== selftest: $(grep -c ^PASS gpurun_out/gemm_selftest3.log) pass, $(grep -c ^FAIL gpurun_out/gemm_selftest3.log) fail"
grep -E "^FAIL|rc=124|rc=13[0-9]" gpurun_out/gemm_selftest3.log | head
sed -n '/throughput-sized/,$p' gpurun_out/gemm_selftest3.log | grep -E "PASS|FAIL"
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
echo "==== bench"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
