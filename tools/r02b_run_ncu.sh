#!/bin/bash
# round 2 (second session) evidence at cfg5, N=1, shipping routes (sparse real codes, low-rank generator route):
# launch list of the timed steps; full-set capture of every kernel of one D+G step pair.  The report stays on the box
# (> 64 MiB); its raw page comes back as csv.
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
GANMF_BENCH_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
  --log-file gpurun_out/r02b_launches_cfg5.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/r02b_launches_cfg5.out 2>&1
echo "launch list rc=$?"
GANMF_BENCH_PROFILER_RANGE=1 ncu --set full --clock-control none --profile-from-start off -c 80 \
  -o /tmp/r02b_step_pair_cfg5 python bench.py --steps 1 --warmup 3 --quick > gpurun_out/r02b_step_pair_cfg5.out 2>&1
echo "full set rc=$?"
ncu -i /tmp/r02b_step_pair_cfg5.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r02b_step_pair_cfg5_raw.csv.gz
ls -la /tmp/*.ncu-rep gpurun_out/r02b_*
