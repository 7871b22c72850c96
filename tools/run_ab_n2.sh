#!/bin/bash
# data-parallel A/B at N GPUs: parity first, then the bench under the scheduling variants
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1
echo "dp parity rc=$?"; grep -E "world=|rel err|DP PARITY|Error|error" gpurun_out/dp_parity_n$N.log | head -24
port=29520
run() { # name, env...
  name=$1; shift; port=$((port+1))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --quick > gpurun_out/abn${N}_$name.log 2> gpurun_out/abn${N}_$name.err
  echo "== $name rc=$?"; grep -E '^\{' gpurun_out/abn${N}_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f rows/s  %.3f ms/step  gemm %.1f TF/s  clocks %s %s' % (d['value'], d['ms_per_step'], d['gemm_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
}
run default A=1
run panels1 GANMF_DP_PANELS=1
run old GANMF_DP_PANELS=1 GANMF_STATIC_SCHED=1
run panels4 GANMF_DP_PANELS=4
run cta8 GANMF_NCCL_MAX_CTAS=8
run cta16 GANMF_NCCL_MAX_CTAS=16
run cta32 GANMF_NCCL_MAX_CTAS=32
