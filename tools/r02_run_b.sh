#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_eval.py tests/test_gpu_baseline_shapes.py -x -q -s 2>&1 | tail -25
for shape in "200000 32768 0.001" "27000 32768 0.005" "200000 131072 0.001" "27000 138000 0.005"; do
  set -- $shape
  python tools/eval_profile.py --items $1 --users $2 --density $3
done
EVAL_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv \
  --log-file gpurun_out/r02_eval_launches_cfg5.csv python tools/eval_profile.py --items 200000 --users 32768 --reps 1 > /dev/null 2>&1
EVAL_PROFILE_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 200 --csv \
  --log-file gpurun_out/r02_eval_launches_cfg4.csv python tools/eval_profile.py --items 27000 --users 32768 --density 0.005 --reps 1 > /dev/null 2>&1
python - <<'PY'
import csv
for cfg in ("cfg5", "cfg4"):
    rows = [r for r in csv.reader(open("gpurun_out/r02_eval_launches_%s.csv" % cfg)) if len(r) > 10 and r[0].isdigit()]
    print(cfg)
    for r in rows:
        print("  %-60s %10.1f us" % (r[4][:60], float(r[-1].replace(",", "")) / 1e3))
PY
