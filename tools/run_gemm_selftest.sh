#!/bin/bash
# Runs the tcgen05 GEMM bring-up matrix on a GPU box; each case in its own
# process under a timeout so a trapped kernel cannot hang the box.
# usage: tools/run_gemm_selftest.sh [out_file]
OUT=${1:-gpurun_out/gemm_selftest.log}
mkdir -p "$(dirname "$OUT")"
BIN=build/gemm_selftest
: > "$OUT"
run() { timeout 60 $BIN "$@" >> "$OUT" 2>&1; echo "  rc=$? args: $*" >> "$OUT"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv >> "$OUT" 2>&1
# precision probe: does TMA/UMMA round or truncate fp32 -> tf32? (dtype 7 = FLOAT32, 11 = TFLOAT32)
run 128 128 64 0 0 128 1 0 1 7
run 128 128 64 0 0 128 1 0 1 11
# smallest case, each major-ness combination
for amn in 0 1; do for bmn in 0 1; do
  run 128 128 32 $amn $bmn 128 1 0
done; done
# multi k-block, ragged edges, both tile widths
for amn in 0 1; do for bmn in 0 1; do for bn in 128 256; do
  run 200 300 100 $amn $bmn $bn 1 0
  run 333 517 1000 $amn $bmn $bn 1 1
done; done; done
# split-K (deterministic reduce kernel) with and without epilogue
for amn in 0 1; do for bmn in 0 1; do
  run 256 384 4100 $amn $bmn 128 5 0
  run 256 384 4100 $amn $bmn 256 3 1
done; done
# tiny N / tiny M (DisGANMF-like skinny shapes ride on TMA zero fill)
run 64 4 1000 0 1 128 1 1
run 1000 1 70 0 0 128 1 0
# throughput-sized cases (cfg4-like): G3 (K-major x MN-major), G4 (MN x MN), G7 (K x K)
run 2048 27000 1024 0 1 128 1 0
run 2048 27000 1024 0 1 256 1 0
run 1024 27000 2048 1 1 256 1 0
run 1024 27000 1024 0 0 256 1 0
run 2048 1024 27000 0 1 256 8 0
cat "$OUT"
