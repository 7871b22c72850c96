#!/bin/bash
# ncu evidence for the judged kernels: (1) launch list of a short bench run, (2) --set full captures of
# the tcgen05 GEMM (several launches = different GEMMs of the step), Adam, CSR gather and top-k.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-users 2048"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# skip the warm-up GEMMs (3 warm-up step pairs x 13 GEMMs) so the captured ones are steady-state
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 39 -c 13 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_adam|csr_gather|topk_rows" -s 12 -c 6 -o gpurun_out/prof_hbm $CMD > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
METRICS='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum$|dram__bytes_write.sum$|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|launch__registers_per_thread|launch__grid_size|sm__throughput.avg.pct|lts__t_bytes.sum '
for f in prof_gemm prof_hbm; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  python - gpurun_out/${f}_raw.csv > gpurun_out/${f}_summary.txt <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = rows[0]
units = rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
tens = [h for h in hdr if "tensor" in h][:12]
print("# columns:", [w for w, _ in idx])
print("# units  :", [units[i] for _, i in idx])
for r in rows[2:]:
    if len(r) != len(hdr): continue
    print(" | ".join((r[i][:60] if w == "Kernel Name" else r[i]) for w, i in idx))
print("# tensor-related metric names available:", tens)
for r in rows[2:3]:
    print("# first row tensor metrics:", {h: r[hdr.index(h)] for h in tens})
PY
  echo "== $f"; cat gpurun_out/${f}_summary.txt | cut -c1-400
done
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches.csv', errors='ignore')))
hdr = None; agg = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum': continue
    name = d['Kernel Name'].split('(')[0][:70]
    v = float(d['Metric Value'].replace(',', '')); unit = d['Metric Unit']
    v = v / 1e3 if unit in ('nsecond', 'ns') else (v * 1e3 if unit in ('msecond', 'ms') else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open('gpurun_out/launch_list_summary.txt', 'w') as f:
    f.write("total device time (us): %.0f  [bench.py --steps 2 --warmup 3: 5 D+G step pairs, eval of 1024+2048 users, kernel micro-benches]\n" % tot)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%6d launches %10.1f us  %5.1f%%  %s\n" % (a[0], a[1], 100 * a[1] / tot, k))
print(open('gpurun_out/launch_list_summary.txt').read())
PY
