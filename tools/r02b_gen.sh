#!/bin/bash
# resident-A generator GEMM: kernel tests under a timeout first (a barrier bug must not hang the box), then the A/B
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "resident or transposed or encode" > gpurun_out/r02b_gen_tests.log 2>&1
rc=$?; echo "kernel tests rc=$rc"; tail -15 gpurun_out/r02b_gen_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_train_parity.py tests/test_gpu_api.py -m gpu -q -x --timeout 300 > gpurun_out/r02b_gen_tests2.log 2>&1
echo "parity tests rc=$?"; tail -4 gpurun_out/r02b_gen_tests2.log
SKIP_TESTS=1 tools/r02b_ab.sh "-;GANMF_GEN_RESIDENT=0" 2>&1 | grep -v "^  "
