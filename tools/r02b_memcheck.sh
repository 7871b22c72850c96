#!/bin/bash
# compute-sanitizer memcheck over the kernel tests, the item-sharded step tests (all route combinations) and the API
# tests (item mode: device CSR transposition); usage: r02b_memcheck.sh [kernels|steps]
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CS="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
if [ "${1:-kernels}" = kernels ]; then
  timeout 900 $CS python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 800 -k "gemm or resident or transposed or encode or gather" > gpurun_out/r02b_memcheck_kernels.log 2>&1
  echo "memcheck kernels rc=$?"; tail -4 gpurun_out/r02b_memcheck_kernels.log
else
  timeout 1200 $CS python -m pytest tests/test_gpu_tp.py tests/test_gpu_api.py tests/test_gpu_train_parity.py -m gpu -q -x --timeout 1100 -k "not lazy and not 100_steps" > gpurun_out/r02b_memcheck_steps.log 2>&1
  echo "memcheck steps rc=$?"; tail -4 gpurun_out/r02b_memcheck_steps.log
fi
