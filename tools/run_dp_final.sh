#!/bin/bash
# data-parallel evidence at N GPUs: parity vs the single-stream oracle, collectives alone, bench (full line) and the
# no-reservation A/B
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/dp_parity.py > gpurun_out/dp_parity_n$N.log 2>&1
echo "dp parity rc=$?"; grep -E "world=|DP PARITY|Error|error" gpurun_out/dp_parity_n$N.log | head -8
timeout 300 $TR --master-port 29512 tools/nccl_probe.py 2>&1 | grep "N=" | tee gpurun_out/nccl_probe_n$N.txt
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n$N.err
grep -E '^\{' gpurun_out/bench_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=%d %.0f rows/s  %.3f ms/step  e2e %.0f  eval %.0f users/s  clocks %s %s' % (d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['eval']['value'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
GANMF_DP_RESERVE_SMS=0 timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --quick > gpurun_out/bench_n${N}_r0.json 2> gpurun_out/bench_n${N}_r0.err
echo "r0 rc=$?"; grep -E '^\{' gpurun_out/bench_n${N}_r0.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('r0: %.0f rows/s  %.3f ms/step' % (d['value'], d['ms_per_step']))"
