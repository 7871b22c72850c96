#!/bin/bash
mkdir -p gpurun_out
GANMF_BENCH_GEMM_TABLE=gpurun_out/gemm_table.txt timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err
cat gpurun_out/gemm_table.txt
python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]); print('value', d['value'], 'ms/step', d['ms_per_step'], 'gemm TF/s', d['roofline']['achieved'], 'eval users/s', d['eval']['value'])"
for run in "GANMF user 1M" "GANMF item LastFM"; do
  timeout 1200 python tools/quality_run.py $run > "gpurun_out/quality_$(echo $run | tr ' ' '_').json" 2> gpurun_out/quality.err
  echo "quality $run rc=$?"; tail -3 gpurun_out/quality.err
  python - "$run" <<'PY'
import json, sys
f = "gpurun_out/quality_%s.json" % sys.argv[1].replace(" ", "_")
try:
    d = json.load(open(f))
    print(d["run"], "train_s %.1f rows/s %.0f eval_s %.2f users/s %.0f worst rel diff %.4f" % (d["train_s"], d["rows_per_s"], d["eval_s"], d["users_per_s"], d["worst_rel_diff_P_R_NDCG_5_20"]))
    for k, v in d["metrics"].items():
        print("   %-14s got %.6f ref %.6f  %+.2f%%" % (k, v["got"], v["ref"], 100 * v["rel_diff"]))
except Exception as e:
    print("no result", e)
PY
done
