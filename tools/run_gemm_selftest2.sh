#!/bin/bash
# MN-major bring-up: primary encoding, then fallbacks, then the full matrix.
OUT=${1:-gpurun_out/gemm_selftest2.log}
mkdir -p "$(dirname "$OUT")"
BIN=build/gemm_selftest
: > "$OUT"
run() { timeout 120 $BIN "$@" >> "$OUT" 2>&1; echo "  rc=$? env[$TC_MN_LAYOUT $TC_MN_SBO $TC_MN_LBO $TC_MN_SWZ] args: $*" >> "$OUT"; }
echo "== primary (layout=1 BASE32B, sbo=512, lbo=4096, TMA 128B_ATOM_32B)" >> "$OUT"
for amn in 0 1; do for bmn in 0 1; do run 128 128 32 $amn $bmn 128 1 0; done; done
echo "== variants on 128x128x32 a_mn=1 b_mn=0" >> "$OUT"
for swz in 4 5 3; do for lay in 1 2; do for sbo in 512 1024; do
  TC_MN_SWZ=$swz TC_MN_LAYOUT=$lay TC_MN_SBO=$sbo run 128 128 32 1 0 128 1 0
done; done; done
echo "== full matrix with primary" >> "$OUT"
for amn in 0 1; do for bmn in 0 1; do for bn in 128 256; do
  run 200 300 100 $amn $bmn $bn 1 0
  run 333 517 1000 $amn $bmn $bn 1 1
done; done; done
for amn in 0 1; do for bmn in 0 1; do
  run 256 384 4100 $amn $bmn 128 5 0
  run 256 384 4100 $amn $bmn 256 3 1
done; done
run 64 4 1000 0 1 128 1 1
run 1000 1 70 0 0 128 1 0
run 24 3706 250 0 0 128 1 0
echo "== throughput-sized (sampled check)" >> "$OUT"
run 2048 27000 1024 0 1 128 1 0
run 2048 27000 1024 0 1 256 1 0
run 1024 27000 2048 1 1 256 1 0
run 1024 27000 1024 0 0 256 1 0
run 27000 1024 2048 1 1 256 1 0
run 2048 1024 27000 0 1 256 8 0
run 2048 1024 27000 0 0 256 8 0
cat "$OUT"
