#!/usr/bin/env python
"""Summaries of the two ncu captures tools/r02b_run_ncu.sh brings back (cfg5, N=1, shipping routes):

    python tools/ncu_step_pair_summary.py gpurun_out/r02b_step_pair_cfg5_raw.csv.gz gpurun_out/r02b_launches_cfg5.csv profiles/

writes  <out>/r02b_ncu_full_step_pair_cfg5.txt      every kernel of one D+G step pair (`ncu --set full`, raw page)
        <out>/r02b_tc_gemm_dram_traffic_cfg5.json   mean DRAM bytes per tensor-core GEMM launch (bench.py roofline.traffic)
        <out>/r02b_ncu_launch_list_cfg5.txt         kernel shares of the timed steps (gpu__time_duration.sum)
"""
import csv
import gzip
import io
import json
import os
import re
import sys

HBM_PEAK = 6552.0       # GB/s, MEASURED_PEAKS.json

# the tensor-core launches of a step pair on the shipping routes, in launch order: label, (M, N, K)
B, I, K_, E = 1024, 200000, 250, 1024
GEMMS = [("G1 F=Pb.V^T", (B, I, K_)), ("M1=V^T.We", (K_, E, I)), ("Hf=Pb.M1+be", (B, E, K_)),
         ("G3 Res2=H2.Wd+bd-X2", (2 * B, I, E)), ("G5 dH2=Res2.Wd^T", (2 * B, E, I)), ("G4 dWd -> Adam(Wd)", (E, I, 2 * B)),
         ("G6 dWe -> Adam(We)", (I, E, 2 * B)),
         ("G1 F (G step)", (B, I, K_)), ("M1 (G step)", (K_, E, I)), ("Hf (G step)", (B, E, K_)),
         ("G3' Res_f", (B, I, E)), ("G5' dHf", (B, E, I)), ("dPb-a -c1.Res_f.V", (B, K_, I)),
         ("dPb-b +dHf.M1^T", (B, K_, E)), ("dV-a -c1.Res_f^T.Pb", (I, K_, B)), ("T1t=Pb^T.dHf", (K_, E, B)),
         ("dV-b +We.T1t^T", (I, K_, E))]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("ganmf::", "")


def read_raw(path):
    txt = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
    rd = csv.reader(io.StringIO(txt[txt.index('"ID"'):]))
    hdr, units = next(rd), next(rd)
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
             "ms": 1.0, "msecond": 1.0}

    def val(r, k):
        try:
            return float(r[col[k]].replace(",", "")) * scale.get(units[col[k]], 1.0)
        except ValueError:
            return float("nan")
    out = []
    for r in rd:
        if not r or not r[0].isdigit():
            continue
        out.append((short(r[col["Kernel Name"]]), val(r, "gpu__time_duration.sum"),
                    val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")))
    return out


def step_pair(raw_path, out_dir):
    rows = read_raw(raw_path)
    is_gemm = lambda n: n.startswith("tc_gemm_kernel") or n.startswith("resident_a_gemm_kernel")
    gemms = [r for r in rows if is_gemm(r[0])]
    out = ["# ncu --set full --clock-control none: every kernel of one D+G step pair at cfg5 (2M x 200k, B=1024, k=250, E=1024), 1 GPU,",
           "# shipping routes (DESIGN.md section 5).  tools/r02b_run_ncu.sh + tools/ncu_step_pair_summary.py; raw page:",
           "# r02b_step_pair_cfg5_raw.csv.gz.  tensor% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed;",
           "# DRAM = dram__bytes_read.sum + dram__bytes_write.sum; copy peak %.0f GB/s (MEASURED_PEAKS.json)." % HBM_PEAK,
           "# (ncu serialises the kernels and runs them cold: shares, not absolutes.)", "#",
           "# --- the %d tensor-core GEMM launches" % len(gemms),
           "# %-42s %-38s %8s %8s %9s %9s %7s %7s" % ("GEMM", "kernel", "ms", "tensor%", "TFLOP/s", "DRAM MB", "GB/s", "of HBM")]
    tot_ms = tot_d = tot_f = 0.0
    labelled = len(gemms) == len(GEMMS)
    for i, (name, ms, tp, d) in enumerate(gemms):
        lab, shape = GEMMS[i] if labelled else ("launch %d" % i, None)
        fl = 2.0 * shape[0] * shape[1] * shape[2] if shape else float("nan")
        out.append("  %-42s %-38s %8.4f %8.1f %9.0f %9.0f %7.0f %6.0f%%" %
                   ("%s %s" % (lab, shape if shape else ""), name[:38], ms, tp, fl / ms / 1e9, d / 1e6, d / ms / 1e6,
                    100 * d / ms / 1e6 / HBM_PEAK))
        tot_ms += ms
        tot_d += d
        tot_f += fl if shape else 0.0
    out.append("# GEMM total %.3f ms, %.1f TFLOP/s executed, mean DRAM bytes per launch %.0f" %
               (tot_ms, tot_f / tot_ms / 1e9, tot_d / max(len(gemms), 1)))
    out += ["#", "# --- every other kernel of the pair (in launch order)"]
    for name, ms, tp, d in rows:
        if not is_gemm(name):
            out.append("  %-40s %8.4f ms  DRAM %8.1f MB %7.0f GB/s" % (name[:40], ms, d / 1e6, d / ms / 1e6))
    allms = sum(r[1] for r in rows)
    out.append("# all kernels %.3f ms; GEMM share %.1f %%" % (allms, 100 * tot_ms / allms))
    out.append("# (p_catchup_kernel is long in the D step of this capture because --steps 1 follows a whole epoch; in the G step")
    out.append("#  of the same minibatch the rows are already current.  In the step it runs on the side stream.)")
    open(os.path.join(out_dir, "r02b_ncu_full_step_pair_cfg5.txt"), "w").write("\n".join(out) + "\n")
    json.dump({"dram_bytes_per_launch_mean": tot_d / max(len(gemms), 1), "launches": len(gemms),
               "note": "%d consecutive tensor-core GEMM launches = one D+G step pair at cfg5 on the shipping routes (ncu --set "
                       "full --clock-control none, profiles/r02b_ncu_full_step_pair_cfg5.txt), dram__bytes_read.sum + "
                       "dram__bytes_write.sum per launch, mean" % len(gemms)},
              open(os.path.join(out_dir, "r02b_tc_gemm_dram_traffic_cfg5.json"), "w"), indent=1)
    print("\n".join(out))


def launch_list(path, out_dir):
    txt = open(path).read()
    rd = csv.DictReader(io.StringIO(txt[txt.index('"ID"'):]))
    agg, tot, n = {}, 0.0, 0
    for row in rd:
        if not row["ID"].isdigit():
            continue
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
        n += 1
    out = ["# ncu launch list of the timed region: bench.py --steps 2 --quick (2 D + 2 G steps), cfg5 2M x 200k, 1 GPU, shipping",
           "# routes; cold-cache, serialised: shares, not absolutes (tools/r02b_run_ncu.sh)",
           "# %d launches, %.1f us total" % (n, tot), "# %-52s %6s %12s %7s" % ("kernel", "calls", "total us", "share")]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-54s %6d %12.1f %6.1f%%" % (k[:54], c, t, 100 * t / tot))
    open(os.path.join(out_dir, "r02b_ncu_launch_list_cfg5.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    step_pair(sys.argv[1], sys.argv[3])
    launch_list(sys.argv[2], sys.argv[3])
