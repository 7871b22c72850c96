#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
port=29540
for cap in default 16 32; do
  port=$((port+1))
  if [ $cap = default ]; then envs="A=1"; else envs="GANMF_NCCL_MAX_CTAS=$cap"; fi
  env $envs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/nccl_probe.py 2>&1 | grep "N=" 
done
run() { # name, env...
  name=$1; shift; port=$((port+1))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 --quick > gpurun_out/abn${N}_$name.log 2> gpurun_out/abn${N}_$name.err
  echo "== $name rc=$?"; grep -E '^\{' gpurun_out/abn${N}_$name.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.0f rows/s  %.3f ms/step  gemm %.1f TF/s  clocks %s %s' % (d['value'], d['ms_per_step'], d['gemm_tflops'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
}
run p1 GANMF_DP_PANELS=1
run p1_r32c32 GANMF_DP_PANELS=1 GANMF_DP_RESERVE_SMS=32 GANMF_NCCL_MAX_CTAS=32
run p2_r32c32 GANMF_DP_PANELS=2 GANMF_DP_RESERVE_SMS=32 GANMF_NCCL_MAX_CTAS=32
run p2_r24c24 GANMF_DP_PANELS=2 GANMF_DP_RESERVE_SMS=24 GANMF_NCCL_MAX_CTAS=24
run p2_r16c16 GANMF_DP_PANELS=2 GANMF_DP_RESERVE_SMS=16 GANMF_NCCL_MAX_CTAS=16
run p2_r32 GANMF_DP_PANELS=2 GANMF_DP_RESERVE_SMS=32
run p4_r32c32 GANMF_DP_PANELS=4 GANMF_DP_RESERVE_SMS=32 GANMF_NCCL_MAX_CTAS=32
run p1_again GANMF_DP_PANELS=1
