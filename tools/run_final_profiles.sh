#!/bin/bash
# round-end evidence: full bench line, reference arm, ncu launch list of the timed steps, cfg5 shard
mkdir -p gpurun_out
GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table_final.txt timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$?"
GANMF_BENCH_PROFILER_RANGE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_timed.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu list rc=$?"
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_timed.csv', errors='ignore')))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'].split('(')[0][:80]
    v = float(d['Metric Value'].replace(',', ''))
    unit = d['Metric Unit']
    v = v / 1e3 if unit in ('nsecond', 'ns') else (v * 1e3 if unit in ('msecond', 'ms') else v)   # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/launch_list_timed_summary.txt', 'w') as f:
    f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none: the 2 timed D+G step pairs of\n")
    f.write("# bench.py --steps 2 --warmup 3 --quick (cudaProfilerStart/Stop around the timed region); cold-cache, serialised\n")
    f.write("total device time (us): %.0f\n" % tot)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%6d launches %10.1f us  %5.1f%%  %s\n" % (a[0], a[1], 100 * a[1] / tot, k))
print(open('gpurun_out/launch_list_timed_summary.txt').read())
PY
gzip -f gpurun_out/launches_timed.csv
timeout 900 python tools/run_cfg5.py > gpurun_out/cfg5_shard.json 2> gpurun_out/cfg5_shard.err
echo "cfg5 rc=$?"; tail -3 gpurun_out/cfg5_shard.json | cut -c1-1500
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1])
print("value %.0f rows/s  ms/step %.3f  gemm %.1f TF/s (share %.2f)  e2e %.0f  launches %d" % (d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['gemm_share_of_step'], d['e2e']['value'], d['gpu_launches']))
print("eval %.0f users/s  hbm_frac %.3f" % (d['eval']['value'], d['eval']['hbm_frac_4I_bytes_per_user']))
for k, v in d['hbm_kernels'].items():
    print("  %-28s %.0f GB/s  frac %.3f" % (k, v['achieved'], v['frac']))
print("cpu", d.get('cpu_baseline')); print("clocks", d['clocks'])
PY
