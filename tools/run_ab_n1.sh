#!/bin/bash
# single-GPU A/B of the GEMM unit scheduler (static vs first-static/rest-dynamic) and the lazy user-factor Adam
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table_$name.txt timeout 600 python bench.py --steps 20 --warmup 3 --quick > gpurun_out/ab_$name.log 2> gpurun_out/ab_$name.err
  echo "== $name rc=$?"; grep -E '^\{' gpurun_out/ab_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f rows/s  %.3f ms/step  gemm %.1f TF/s  clocks %s %s' % (d['value'], d['ms_per_step'], d['gemm_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
}
run dyn A=1
run static GANMF_STATIC_SCHED=1
run dense GANMF_NO_LAZY_ADAM=1
run dyn2 A=1
paste gpurun_out/gemm_table_dyn.txt gpurun_out/gemm_table_static.txt | cut -c1-140
