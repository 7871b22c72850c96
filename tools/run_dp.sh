#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1
echo "dp parity rc=$?"; grep -E "world=|rel err|DP PARITY|Error|error" gpurun_out/dp_parity_n$N.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; grep -E '^\{' gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err
