#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
port=29560
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1
echo "dp parity rc=$?"; grep -E "world=|DP PARITY|Error|error" gpurun_out/dp_parity_n$N.log | head -24
run() { # name, env...
  name=$1; shift; port=$((port+1))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --quick > gpurun_out/abn${N}_$name.log 2> gpurun_out/abn${N}_$name.err
  echo "== $name rc=$?"; grep -E '^\{' gpurun_out/abn${N}_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f rows/s  %.3f ms/step  gemm %.1f TF/s  clocks %s %s' % (d['value'], d['ms_per_step'], d['gemm_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
}
run r0 GANMF_DP_RESERVE_SMS=0
run r32c32 GANMF_DP_RESERVE_SMS=32 GANMF_NCCL_MAX_CTAS=32
run r24c24 GANMF_DP_RESERVE_SMS=24 GANMF_NCCL_MAX_CTAS=24
run r16c16 GANMF_DP_RESERVE_SMS=16 GANMF_NCCL_MAX_CTAS=16
run r20c32 GANMF_DP_RESERVE_SMS=20 GANMF_NCCL_MAX_CTAS=32
run r32c0 GANMF_DP_RESERVE_SMS=32
run r40c40 GANMF_DP_RESERVE_SMS=40 GANMF_NCCL_MAX_CTAS=40
run r32c32b GANMF_DP_RESERVE_SMS=32 GANMF_NCCL_MAX_CTAS=32
