#!/bin/bash
# source-level ncu capture of the tcgen05 GEMMs of steady-state steps (after the bench's warm epoch and the 3
# warm-up step pairs = 135*13 + 39 launches): 2 D steps (G1 F=Pb.V^T, G2 codes (split-K), G3 residual,
# G5 dH (split-K), G4 dWd+Adam, G6 dWe+Adam) and 1 G step (G1, G2, G3', G5', G7 dF, G8 dV, G9 dPb)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --quick"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 1794 -c 19 -o gpurun_out/prof_gemm_v8 $CMD > gpurun_out/ncu_gemm_v8.log 2>&1
echo "capture rc=$?"
ls -la gpurun_out/prof_gemm_v8.ncu-rep
