"""Generates the committed golden fixtures from the UNMODIFIED reference code.

Run ONLY in the build container (needs /root/reference):  python tests/golden/make_golden.py
Nothing at test / bench time reads /root/reference; the .npz / .json files written next to this
script are what travels.

Fixtures
  eval_small_*.npz   random score matrices pushed through the reference's own
                     Base/BaseRecommender.recommend + Base/Evaluation/Evaluator.EvaluatorHoldout
                     (numpy>=2 shim: np.int=int, np.float=float, np.bool=bool) -> all 19 metrics
                     per cutoff + the recommendation lists.  eval_small_ignore: the same with
                     EvaluatorHoldout(ignore_items=...) (40 of 320 items); eval_small_ignore_users:
                     EvaluatorHoldout(ignore_users=...) (23 of 140 users).
  lastfm_kat.npz     the one surviving reference checkpoint
                     (feature_matching/GANMF_item_LastFM_00/GANMF_item_LastFM/GANMF_item.data-*):
                     generator factors + committed LastFM train/test split + the stored
                     test_results.pkl (19 metrics x 4 cutoffs).
  splits_*.npz       committed dataset splits (experiments/datasets/*.npz), indices only
                     (all ratings are 1.0), for the end-to-end quality runs.
  quality_targets.json   test_results/<run>/test_results.pkl + experiments/<run>/best_params.pkl
"""
import json
import os
import pickle
import sys

import numpy as np
import scipy.sparse as sps

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    np.int, np.float, np.bool = int, float, bool          # removed aliases the reference still uses
    sys.path.insert(0, REF)
    from Base.BaseRecommender import BaseRecommender
    from Base.Evaluation.Evaluator import EvaluatorHoldout
    return BaseRecommender, EvaluatorHoldout


def make_eval_fixture(name, n_users, n_items, train_density, test_density, cutoffs, seed, ratings,
                      short_rows=False, n_ignore=0, n_ignore_users=0):
    BaseRecommender, EvaluatorHoldout = load_reference()
    rs = np.random.RandomState(seed)
    train = sps.random(n_users, n_items, train_density, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1.0
    test = sps.random(n_users, n_items, test_density, format="csr", dtype=np.float32, random_state=rs)
    test = test - test.multiply(train.astype(bool))       # disjoint from train
    test = sps.csr_matrix(test)
    test.eliminate_zeros()
    if ratings:
        test.data[:] = rs.randint(1, 6, size=test.nnz).astype(np.float32)
    else:
        test.data[:] = 1.0
    if short_rows:
        # a few users who have seen almost everything: lists shorter than the cutoff
        train = train.tolil()
        for u in range(0, n_users, 17):
            keep = rs.choice(n_items, size=3 + (u % 5), replace=False)
            row = np.ones(n_items, dtype=np.float32)
            row[keep] = 0
            row[test[u].indices] = 0
            train[u] = row
        train = sps.csr_matrix(train, dtype=np.float32)
    # distinct values per row (a shuffled grid), so the reference's unspecified tie order never matters
    scores = np.stack([rs.permutation(n_items) for _ in range(n_users)]).astype(np.float32)
    scores = (scores * np.float32(0.0078125) - np.float32(1.5)).astype(np.float32)
    assert all(len(np.unique(r)) == n_items for r in scores)

    class FixedScores(BaseRecommender):
        RECOMMENDER_NAME = "FixedScores"

        def __init__(self, urm):
            self.URM_train = urm

        def _compute_item_score(self, user_id_array, items_to_compute=None):
            return scores[user_id_array].copy()

    rec = FixedScores(train)
    # ignore_items (Evaluator.py:128-134,369-370,410-411): the evaluator masks them through
    # set_items_to_ignore / remove_CustomItems_flag and shrinks COVERAGE_ITEM's denominator
    ignore = np.sort(rs.choice(n_items, size=n_ignore, replace=False)).astype(np.int64) if n_ignore else None
    # ignore_users (Evaluator.py:171-176): dropped from usersToEvaluate, Coverage_User's denominator shrinks
    ign_users = np.sort(rs.choice(n_users, size=n_ignore_users, replace=False)).astype(np.int64) \
        if n_ignore_users else None
    ev = EvaluatorHoldout(test, cutoff_list=list(cutoffs), exclude_seen=True, ignore_items=ignore,
                          ignore_users=ign_users)
    results, _ = ev.evaluateRecommender(rec)
    users = np.array(sorted(ev.usersToEvaluate))
    assert list(ev.usersToEvaluate) == sorted(ev.usersToEvaluate)   # the set difference iterated in ascending order
    if n_ignore:
        rec.set_items_to_ignore(ignore)
    lists, _ = rec.recommend(users, cutoff=max(cutoffs), remove_seen_flag=True, return_scores=True,
                             remove_CustomItems_flag=bool(n_ignore))
    flat = np.full((len(users), max(cutoffs)), -1, dtype=np.int32)
    for i, l in enumerate(lists):
        flat[i, :len(l)] = l
    metric_names = sorted(results[cutoffs[0]].keys())
    table = np.array([[float(results[c][m]) for m in metric_names] for c in cutoffs], dtype=np.float64)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        scores=scores, train_indptr=train.indptr, train_indices=train.indices,
        test_indptr=test.indptr, test_indices=test.indices, test_data=test.data,
        shape=np.array([n_users, n_items]), cutoffs=np.array(cutoffs), users=users, lists=flat,
        metric_names=np.array(metric_names), results=table,
        ignore_items=ignore if n_ignore else np.zeros(0, dtype=np.int64),
        ignore_users=ign_users if n_ignore_users else np.zeros(0, dtype=np.int64))
    print(name, "users evaluated", len(users), "P@%d" % cutoffs[0], results[cutoffs[0]]["PRECISION"])


def read_tf_bundle_floats(path, shapes):
    """TF bundle data file = raw little-endian fp32 tensors concatenated in alphabetical name order
    (verified in SURVEY.md section 4)."""
    raw = np.fromfile(path, dtype="<f4")
    out, off = {}, 0
    for name in sorted(shapes):
        n = int(np.prod(shapes[name]))
        out[name] = raw[off:off + n].reshape(shapes[name]).copy()
        off += n
    assert off == raw.size, (off, raw.size)
    return out


def make_lastfm_kat():
    d = os.path.join(REF, "feature_matching/GANMF_item_LastFM_00/GANMF_item_LastFM")
    bp = pickle.load(open(os.path.join(d, "build_params.pkl"), "rb"))
    k, E = bp["num_factors"], bp["emb_dim"]
    train = sps.load_npz(os.path.join(REF, "experiments/datasets/LastFM_URM_train.npz")).tocsr()
    test = sps.load_npz(os.path.join(REF, "experiments/datasets/LastFM_URM_test.npz")).tocsr()
    n_users, n_items = train.shape            # item mode: rows of the model = items
    shapes = {"autoencoder/decoding/bias": (n_users,), "autoencoder/decoding/kernel": (E, n_users),
              "autoencoder/encoding/bias": (E,), "autoencoder/encoding/kernel": (n_users, E),
              "generator/item_embeddings": (n_users, k), "generator/user_embeddings": (n_items, k)}
    t = read_tf_bundle_floats(os.path.join(d, "GANMF_item.data-00000-of-00001"), shapes)
    res = pickle.load(open(os.path.join(d, "test_results.pkl"), "rb"))
    cutoffs = sorted(res.keys())
    names = sorted(res[cutoffs[0]].keys())
    table = np.array([[float(res[c][m]) for m in names] for c in cutoffs])
    assert np.all(train.data == 1.0) and np.all(test.data == 1.0)
    np.savez_compressed(
        os.path.join(HERE, "lastfm_kat.npz"),
        user_embeddings=t["generator/user_embeddings"], item_embeddings=t["generator/item_embeddings"],
        train_indptr=train.indptr, train_indices=train.indices.astype(np.int32),
        test_indptr=test.indptr, test_indices=test.indices.astype(np.int32),
        shape=np.array(train.shape), cutoffs=np.array(cutoffs), metric_names=np.array(names),
        results=table, num_factors=k, emb_dim=E)
    print("lastfm_kat: P@5 stored", res[5]["PRECISION"])


def make_splits():
    for ds in ("Movielens1M", "LastFM", "Movielenshetrec2011"):
        out = {}
        for part in ("train", "test"):
            m = sps.load_npz(os.path.join(REF, "experiments/datasets/%s_URM_%s.npz" % (ds, part))).tocsr()
            m.sort_indices()
            assert np.all(m.data == 1.0)
            out[part + "_indptr"] = m.indptr.astype(np.int32)
            out[part + "_indices"] = m.indices.astype(np.uint16 if m.shape[1] < 65536 else np.int32)
            out["shape"] = np.array(m.shape)
        np.savez_compressed(os.path.join(HERE, "splits_%s.npz" % ds), **out)
        print("splits", ds, out["shape"])


def make_quality_targets():
    tgt = {}
    for algo in ("GANMF", "DisGANMF"):
        for mode in ("user", "item"):
            for ds in ("1M", "hetrec2011", "LastFM"):
                run = "%s_%s_%s" % (algo, mode, ds)
                try:
                    bp = pickle.load(open(os.path.join(REF, "experiments", run, "best_params.pkl"), "rb"))
                    res = pickle.load(open(os.path.join(REF, "test_results", run, "test_results.pkl"), "rb"))
                except Exception as e:        # noqa
                    print("skip", run, e)
                    continue
                tgt[run] = {"best_params": {k: (v if isinstance(v, str) else float(v)) for k, v in bp.items()},
                            "results": {str(c): {m: float(v) for m, v in res[c].items()} for c in res}}
    json.dump(tgt, open(os.path.join(HERE, "quality_targets.json"), "w"), indent=1, sort_keys=True)
    print("quality targets:", sorted(tgt))


if __name__ == "__main__":
    if "--only-ignore" in sys.argv:
        make_eval_fixture("eval_small_ignore", 130, 320, 0.06, 0.04, (5, 10, 20), 17, ratings=True, n_ignore=40)
        make_eval_fixture("eval_small_ignore_users", 140, 200, 0.06, 0.05, (5, 10), 19, ratings=False,
                          n_ignore_users=23)
        sys.exit(0)
    make_eval_fixture("eval_small_ignore", 130, 320, 0.06, 0.04, (5, 10, 20), 17, ratings=True, n_ignore=40)
    make_eval_fixture("eval_small_ignore_users", 140, 200, 0.06, 0.05, (5, 10), 19, ratings=False, n_ignore_users=23)
    make_eval_fixture("eval_small_implicit", 150, 400, 0.05, 0.03, (5, 10, 20, 50), 7, ratings=False)
    make_eval_fixture("eval_small_ratings", 120, 300, 0.08, 0.04, (1, 5, 10), 11, ratings=True)
    make_eval_fixture("eval_small_shortlists", 100, 60, 0.10, 0.08, (5, 20, 50), 13, ratings=True,
                      short_rows=True)
    make_lastfm_kat()
    make_splits()
    make_quality_targets()
