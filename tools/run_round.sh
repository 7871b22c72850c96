#!/bin/bash
# One GPU-box visit: tests, smoke, bench (both arms).  Logs under gpurun_out/.
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > /dev/null 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
echo "==== tests"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -20
echo "==== smoke"; tail -5 gpurun_out/smoke.log
echo "==== bench"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "==== ref"; cat gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
