#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/gemm_diag.log; : > $OUT
BIN=build/gemm_selftest
for epi in 0 1; do
  echo "== TC_DBG_EPI=$epi" >> $OUT
  for args in "2048 27000 1024 0 1 256 1 0" "27000 1024 2048 1 1 256 1 0" "1024 27000 1024 0 0 256 1 0" "2048 27000 1024 0 1 128 1 0" "1024 27000 2048 1 1 256 1 0" "2048 1024 27000 0 1 256 8 0" "27000 250 1024 1 1 256 1 0"; do
    TC_DBG_EPI=$epi timeout 120 $BIN $args 2>&1 | grep -E "PASS|FAIL" >> $OUT
  done
done
cat $OUT
bash tools/run_gemm_selftest2.sh gpurun_out/gemm_selftest3.log > /dev/null 2>&1
echo "==== selftest: $(grep -c ^PASS gpurun_out/gemm_selftest3.log) pass, $(grep -c ^FAIL gpurun_out/gemm_selftest3.log) fail (10 expected from the encoding-variant section)"
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -20
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
echo "==== bench"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
