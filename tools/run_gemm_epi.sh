#!/bin/bash
OUT=gpurun_out/gemm_epi.log; mkdir -p gpurun_out; : > $OUT
BIN=build/gemm_selftest
run() { timeout 120 $BIN "$@" 2>&1 | grep -E "PASS|FAIL|error|Error" >> $OUT; echo "   rc=$? TC_MT=$TC_MT args: $*" >> $OUT; }
for mt in 1 2; do export TC_MT=$mt
for amn in 0 1; do for bmn in 0 1; do
  run 333 517 1000 $amn $bmn 256 1 2
  run 512 1024 600 $amn $bmn 256 1 2
  run 640 768 300 $amn $bmn 128 1 2
  run 700 384 4100 $amn $bmn 256 3 2
done; done; done
export TC_MT=1
echo "== throughput with one addend + bias + sumsq (G3-like), sampled check" >> $OUT
run 2048 27000 1024 0 1 256 1 2
run 1024 27000 1024 0 1 256 1 2
run 1024 27000 1024 0 0 256 1 2
run 2048 27000 1024 0 1 256 1 0
cat $OUT
