#!/usr/bin/env python
"""End-to-end quality parity (RunBestParameters.py equivalent): train with the committed best
hyper-parameters on the committed split, evaluate at cutoffs 5/10/20/50, compare with the
reference's stored test_results (tests/golden/quality_targets.json).

    python tools/quality_run.py GANMF user 1M [--epochs N]
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


def main():
    algo, mode, ds = sys.argv[1:4]
    run = "%s_%s_%s" % (algo, mode, ds)
    tgt = load_quality_targets()[run]
    bp = dict(tgt["best_params"])
    for k in ("epochs", "num_factors", "batch_size", "emb_dim", "d_layers", "d_nodes"):
        if k in bp:
            bp[k] = int(bp[k])
    if "--epochs" in sys.argv:
        bp["epochs"] = int(sys.argv[sys.argv.index("--epochs") + 1])
    split = load_split(DS[ds])
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.GANRec.GANMF import GANMF
    seed = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 1337
    np.random.seed(seed)                                      # RunBestParameters.py:81 (1337)
    cls = GANMF if algo == "GANMF" else DisGANMF
    model = cls(split["train"], mode=mode, seed=seed, is_experiment=True)
    t0 = time.time()
    model.fit(validation_set=None, sample_every=None, validation_evaluator=None, **bp)
    t_train = time.time() - t0
    ev = EvaluatorHoldout(split["test"], [5, 10, 20, 50], exclude_seen=True)
    t0 = time.time()
    res, _ = ev.evaluateRecommender(model)
    t_eval = time.time() - t0
    rows = model.num_users * bp["epochs"]
    out = {"run": run, "train_s": t_train, "rows_per_s": rows / t_train, "eval_s": t_eval,
           "users_per_s": len(ev.usersToEvaluate) / t_eval, "d_loss_last": model.train_d_loss[-1],
           "g_loss_last": model.train_g_loss[-1], "metrics": {}}
    worst = 0.0
    for c in (5, 10, 20):
        for m in ("PRECISION", "RECALL", "NDCG", "MAP"):
            got, want = float(res[c][m]), tgt["results"][str(c)][m]
            rel = got / want - 1.0
            out["metrics"]["%s@%d" % (m, c)] = {"got": got, "ref": want, "rel_diff": rel}
            if m != "MAP":
                worst = max(worst, abs(rel))
    out["worst_rel_diff_P_R_NDCG_5_20"] = worst
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
