#!/bin/bash
mkdir -p gpurun_out
for run in "GANMF user 1M" "GANMF item 1M" "GANMF user hetrec2011" "GANMF item hetrec2011" "GANMF user LastFM" "GANMF item LastFM" "DisGANMF user hetrec2011" "DisGANMF item hetrec2011" "DisGANMF user 1M" "DisGANMF item LastFM"; do
  f="gpurun_out/quality_$(echo $run | tr ' ' '_').json"
  timeout 900 python tools/quality_run.py $run > "$f" 2> gpurun_out/quality.err
  rc=$?
  python - "$f" "$rc" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    m = d["metrics"]
    print("%-28s train %6.1fs %8.0f rows/s | eval %5.2fs | worst |rel| P/R/NDCG@5-20 %.4f | P@5 %+.2f%% NDCG@10 %+.2f%% R@20 %+.2f%%" % (
        d["run"], d["train_s"], d["rows_per_s"], d["eval_s"], d["worst_rel_diff_P_R_NDCG_5_20"],
        100 * m["PRECISION@5"]["rel_diff"], 100 * m["NDCG@10"]["rel_diff"], 100 * m["RECALL@20"]["rel_diff"]))
except Exception as e:
    print(sys.argv[1], "rc", sys.argv[2], "no result:", e)
PY
done
tail -3 gpurun_out/quality.err
