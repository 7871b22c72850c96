#!/bin/bash
# ncu launch list (device time per launch) of a short bench run + one full capture of the top GEMM.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-users 2048 \
  > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv', errors='ignore')))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'].split('(')[0][:70]
    v = float(d['Metric Value'].replace(',', ''))
    unit = d['Metric Unit']
    v = v / 1e3 if unit in ('nsecond', 'ns') else (v * 1e3 if unit in ('msecond', 'ms') else v)   # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("total device time (us): %.0f" % tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6d launches %10.1f us  %5.1f%%  %s" % (a[0], a[1], 100 * a[1] / tot, k))
PY
