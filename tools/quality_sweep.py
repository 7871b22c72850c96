#!/usr/bin/env python
"""All 12 committed end-to-end runs (RunBestParameters.py equivalent) x several seeds: relative difference of
PRECISION / RECALL / NDCG @ 5, 10, 20 against the reference's stored test_results, per seed and for the seed mean.

    python tools/quality_sweep.py [--seeds 1337,1,2] [--runs GANMF_user_1M,...] > profiles/r02_quality_sweep.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import load_quality_targets, load_split  # noqa: E402

DS = {"1M": "Movielens1M", "hetrec2011": "Movielenshetrec2011", "LastFM": "LastFM"}
KEYS = [(m, c) for c in (5, 10, 20) for m in ("PRECISION", "RECALL", "NDCG")]


def one_run(run, seed):
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.GANRec.GANMF import GANMF
    algo, mode, ds = run.split("_")
    tgt = load_quality_targets()[run]
    bp = dict(tgt["best_params"])
    for k in ("epochs", "num_factors", "batch_size", "emb_dim", "d_layers", "d_nodes"):
        if k in bp:
            bp[k] = int(bp[k])
    split = load_split(DS[ds])
    np.random.seed(seed)                                      # RunBestParameters.py:81
    model = (GANMF if algo == "GANMF" else DisGANMF)(split["train"], mode=mode, seed=seed, is_experiment=True)
    t0 = time.time()
    model.fit(validation_set=None, sample_every=None, validation_evaluator=None, **bp)
    t_train = time.time() - t0
    res, _ = EvaluatorHoldout(split["test"], [5, 10, 20, 50], exclude_seen=True).evaluateRecommender(model)
    model._engine.close()
    return {"%s@%d" % (m, c): float(res[c][m]) for m, c in KEYS}, t_train


def main():
    seeds = [1337, 1, 2]
    runs = sorted(load_quality_targets())
    if "--seeds" in sys.argv:
        seeds = [int(x) for x in sys.argv[sys.argv.index("--seeds") + 1].split(",")]
    if "--runs" in sys.argv:
        runs = sys.argv[sys.argv.index("--runs") + 1].split(",")
    out = {}
    for run in runs:
        ref = load_quality_targets()[run]["results"]
        per_seed, times = [], []
        for s in seeds:
            vals, t = one_run(run, s)
            per_seed.append(vals)
            times.append(t)
        rec = {"seeds": seeds, "train_s": times, "metrics": {}}
        worst_mean, worst_any, inside = 0.0, 0.0, True
        for m, c in KEYS:
            key = "%s@%d" % (m, c)
            want = ref[str(c)][m]
            got = [v[key] for v in per_seed]
            mean = float(np.mean(got))
            rec["metrics"][key] = {"ref": want, "seeds": got, "mean_rel": mean / want - 1.0,
                                   "ref_inside_seed_spread": bool(min(got) <= want <= max(got))}
            worst_mean = max(worst_mean, abs(mean / want - 1.0))
            worst_any = max(worst_any, max(abs(g / want - 1.0) for g in got))
            inside = inside and rec["metrics"][key]["ref_inside_seed_spread"]
        rec.update(worst_abs_rel_of_seed_mean=worst_mean, worst_abs_rel_any_seed=worst_any, ref_inside_spread_all=inside)
        out[run] = rec
        sys.stderr.write("%-28s mean worst %.4f  any-seed worst %.4f  ref inside spread (all 9): %s  train %s s\n" %
                         (run, worst_mean, worst_any, inside, ["%.1f" % t for t in times]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
