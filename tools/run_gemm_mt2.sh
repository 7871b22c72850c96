#!/bin/bash
OUT=gpurun_out/gemm_mt2.log; mkdir -p gpurun_out; : > $OUT
BIN=build/gemm_selftest
run() { timeout 120 $BIN "$@" 2>&1 | grep -E "PASS|FAIL|error|Error" >> $OUT; echo "   rc=$? TC_MT=$TC_MT args: $*" >> $OUT; }
export TC_MT=2
echo "== correctness, 256-row CTA tiles" >> $OUT
for amn in 0 1; do for bmn in 0 1; do
  run 256 256 32 $amn $bmn 256 1 0
  run 200 300 100 $amn $bmn 256 1 0
  run 333 517 1000 $amn $bmn 256 1 1
  run 700 384 4100 $amn $bmn 256 3 1
  run 512 384 4100 $amn $bmn 256 5 0
done; done
run 24 3706 250 0 0 256 1 0
echo "== throughput (sampled check)" >> $OUT
for mt in 1 2; do
  export TC_MT=$mt
  run 2048 1024 27000 0 1 256 9 0
  run 2048 1024 27000 0 0 256 9 0
  run 2048 1024 27000 0 1 256 14 0
  run 1024 27000 2048 1 1 256 1 0
  run 27000 1024 2048 1 1 256 1 0
  run 1024 1024 27000 0 0 256 9 0
  run 2048 27000 1024 0 1 256 1 0
done
cat $OUT
