#!/bin/bash
# round 2: item-sharded parity over NCCL + the default bench at N GPUs.  usage: r02_run_n.sh N [tag]
N=${1:-2}; TAG=${2:-a}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tools/tp_parity.py > gpurun_out/r02_tp_parity_n$N.log 2>&1
echo "tp_parity rc=$?"; grep -E "PARITY|->" gpurun_out/r02_tp_parity_n$N.log
GANMF_BENCH_GEMM_TABLE=gpurun_out/r02_gemm_table_n$N timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/r02_bench_n${N}_$TAG.json 2> gpurun_out/r02_bench_n${N}_$TAG.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_n${N}_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_n${N}_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, d["e2e"], d["eval"], d["roofline"]["achieved"], d["clocks"])
except Exception as e:
    print("no line", e)
PY
cat gpurun_out/r02_gemm_table_n$N.cfg5 2>/dev/null
