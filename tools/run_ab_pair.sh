#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh > /dev/null 2>&1
echo "==== tests"; grep -E "passed|failed|^FAILED|Error" gpurun_out/pytest_gpu.log | tail -10
run() { # name, env...
  name=$1; shift
  env "$@" GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table_$name.txt timeout 600 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/ab_$name.log 2> gpurun_out/ab_$name.err
  echo "== $name rc=$?"; grep -E '^\{' gpurun_out/ab_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f rows/s  %.3f ms/step  gemm %.1f TF/s  clocks %s %s' % (d['value'], d['ms_per_step'], d['gemm_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
}
run pair1 GANMF_PAIR=1
run pair0 GANMF_PAIR=0
run pair1b GANMF_PAIR=1
paste gpurun_out/gemm_table_pair1.txt gpurun_out/gemm_table_pair0.txt | cut -c1-140
