#!/bin/bash
OUT=gpurun_out/gemm_diag.log; mkdir -p gpurun_out; : > $OUT
BIN=build/gemm_selftest
for epi in 0 1 2; do
  echo "== TC_DBG_EPI=$epi" >> $OUT
  TC_DBG_EPI=$epi timeout 120 $BIN 2048 27000 1024 0 1 256 1 0 2>&1 | grep -E "PASS|FAIL" >> $OUT
  TC_DBG_EPI=$epi timeout 120 $BIN 27000 1024 2048 1 1 256 1 0 2>&1 | grep -E "PASS|FAIL" >> $OUT
  TC_DBG_EPI=$epi timeout 120 $BIN 1024 27000 1024 0 0 256 1 0 2>&1 | grep -E "PASS|FAIL" >> $OUT
  TC_DBG_EPI=$epi timeout 120 $BIN 2048 27000 1024 0 1 128 1 0 2>&1 | grep -E "PASS|FAIL" >> $OUT
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 1 -c 1 -o gpurun_out/gemm_g3 \
   $BIN 2048 27000 1024 0 1 256 1 0 > gpurun_out/ncu_gemm.log 2>&1
ncu -i gpurun_out/gemm_g3.ncu-rep --page raw --csv > gpurun_out/gemm_g3_raw.csv 2>/dev/null
ncu -i gpurun_out/gemm_g3.ncu-rep --page details > gpurun_out/gemm_g3_details.txt 2>/dev/null
cat $OUT
grep -E "Duration|DRAM Throughput|L2 Cache Throughput|Compute \(SM\)|Memory Throughput|Issue Slots|Registers|Warp Cycles Per Issued|No Eligible|Stall" gpurun_out/gemm_g3_details.txt | head -40
