#!/bin/bash
mkdir -p gpurun_out
python tools/eval_diag.py 2>&1 | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/eval_launches.csv python tools/eval_diag.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/eval_launches.csv', errors='ignore')))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'].split('(')[0][:60] + " grid" + d['Grid Size'] + " blk" + d['Block Size']
    v = float(d['Metric Value'].replace(',', '')) / 1e3
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("total device time (us): %.0f" % tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("%6d launches %10.1f us  %5.1f%%  %s" % (a[0], a[1], 100 * a[1] / tot, k))
PY
