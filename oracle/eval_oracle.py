"""NumPy restatement of the reference's recommend -> evaluate path (TEST INFRASTRUCTURE).

Follows
  Base/BaseRecommender.py:93-100,155-247          recommend(): seen mask, top-cutoff, drop -inf
  Base/Evaluation/Evaluator.py:119-179,234-414    EvaluatorHoldout block loop, averaging, F1
  Base/Evaluation/metrics.py:30-570,576-722       per-user metrics and count-based objects
One deliberate tightening: ties in the score are broken by LOWEST item index (the reference's
argpartition + quicksort order among equal scores is unspecified, BaseRecommender.py:214-221).

`promotion` selects the scalar dtype rules the reference code runs under:
  "legacy"  numpy 1.16.2, the version the reference pins (conda_requirements.txt): a numpy
            float32 scalar combined with a Python int/float gives float64, so PRECISION,
            RECALL, PRECISION_RECALL_MIN_DEN and ROC_AUC are float64 per user and every
            running sum is float64.  THIS is what the device evaluator implements.
  "nep50"   numpy >= 2 (this container): the same expressions stay float32 and the running
            sums of those metrics are float32.  Used only to pin this restatement bit for bit
            against the unmodified reference code executed here (tests/golden/make_golden.py).
Pinned by: tests/golden/eval_*.npz (reference run here), metrics_Test.py vectors,
the LastFM checkpoint KAT (tests/golden/lastfm_kat.npz).
"""
import numpy as np
import scipy.sparse as sps

METRIC_NAMES = ["ROC_AUC", "PRECISION", "PRECISION_RECALL_MIN_DEN", "RECALL", "MAP", "MRR", "NDCG",
                "F1", "HIT_RATE", "ARHR", "RMSE", "NOVELTY", "AVERAGE_POPULARITY",
                "DIVERSITY_MEAN_INTER_LIST", "DIVERSITY_HERFINDAHL", "COVERAGE_ITEM",
                "COVERAGE_USER", "DIVERSITY_GINI", "SHANNON_ENTROPY"]


# ------------------------------------------------------------------ recommend()
def remove_seen(scores, urm_train, user_ids):
    """scores[u, seen(u)] = -inf from the CSR row of URM_train (BaseRecommender.py:93-100,189-194)."""
    scores = np.array(scores, dtype=np.float32, copy=True)
    indptr, indices = urm_train.indptr, urm_train.indices
    for r, u in enumerate(user_ids):
        scores[r, indices[indptr[u]:indptr[u + 1]]] = -np.inf
    return scores


def topk_lowest_index(scores, k):
    """Top-k item ids by descending score, ties -> lowest index; returns (idx[n,k], val[n,k])."""
    order = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return order.astype(np.int32), np.take_along_axis(scores, order, axis=1)


def recommend(scores, urm_train, user_ids, cutoff, remove_seen_flag=True):
    """BaseRecommender.recommend (:155-247) given the raw score rows of `user_ids`.
    Returns (list of id lists with -inf entries dropped, masked scores)."""
    if remove_seen_flag:
        scores = remove_seen(scores, urm_train, user_ids)
    else:
        scores = np.array(scores, dtype=np.float32, copy=True)
    idx, val = topk_lowest_index(scores, cutoff)
    lists = [idx[r][np.logical_not(np.isinf(val[r]))].tolist() for r in range(len(user_ids))]
    return lists, scores


# ------------------------------------------------------------------ per-user metrics
def _sum_f32(x):
    return np.sum(x, dtype=np.float32)


def precision(is_rel, legacy=True):                       # metrics.py:612-620
    if len(is_rel) == 0:
        return 0.0
    h = _sum_f32(is_rel)
    return np.float64(h) / len(is_rel) if legacy else h / np.float32(len(is_rel))


def precision_recall_min_denominator(is_rel, n_test, legacy=True):   # :623-631
    if len(is_rel) == 0:
        return 0.0
    h = _sum_f32(is_rel)
    d = min(n_test, len(is_rel))
    return np.float64(h) / d if legacy else h / np.float32(d)


def recall(is_rel, n_test, legacy=True):                  # :662-667
    h = _sum_f32(is_rel)
    return np.float64(h) / n_test if legacy else h / np.float32(n_test)


def rr(is_rel):                                           # :670-678
    ranks = np.arange(1, len(is_rel) + 1)[is_rel]
    return 1.0 / ranks[0] if len(ranks) > 0 else 0.0


def arhr(is_rel):                                         # :596-608
    p = 1 / np.arange(1, len(is_rel) + 1, 1.0, dtype=np.float64)
    return is_rel.dot(p)


def average_precision(is_rel, n_test):                    # :681-690
    if len(is_rel) == 0:
        return 0.0
    p_at_k = is_rel * np.cumsum(is_rel, dtype=np.float32) / (1 + np.arange(is_rel.shape[0]))
    return np.sum(p_at_k) / np.min([n_test, is_rel.shape[0]])


def roc_auc(is_rel, legacy=True):                         # :576-592
    ranks = np.arange(len(is_rel))
    pos, neg = ranks[is_rel], ranks[~is_rel]
    if len(neg) == 0:
        return 1.0
    if len(pos) == 0:
        return 0.0
    s = np.float64(0.0) if legacy else 0.0
    for p in pos:
        s = s + _sum_f32(p < neg)
    d = pos.shape[0] * neg.shape[0]
    return s / d if legacy else s / np.float32(d)


def dcg(scores):                                          # :720-722 (all float32)
    return np.sum(np.divide(np.power(2, scores) - 1,
                            np.log(np.arange(scores.shape[0], dtype=np.float32) + 2)), dtype=np.float32)


def ndcg(ranked, pos_items, relevance):                   # :693-717 (ranked already cut to `at`)
    it2rel = {it: r for it, r in zip(pos_items, relevance)}
    rank_scores = np.asarray([it2rel.get(it, 0.0) for it in ranked], dtype=np.float32)
    ideal = dcg(np.sort(relevance)[::-1][:len(ranked)])
    got = dcg(rank_scores)
    if got == 0.0:
        return 0.0
    return got / ideal


def rmse(all_scores, relevant_items, ratings):            # :634-659
    err = (all_scores[relevant_items] - ratings) ** 2
    finite = np.isfinite(err)
    if finite.sum() == 0:
        return np.nan
    return np.sqrt(np.sum(err[finite]) / finite.sum())


# ------------------------------------------------------------------ EvaluatorHoldout
def users_to_evaluate(urm_test, min_ratings=1):
    """Evaluator.__init__ (:152-179): rows of URM_test with >= min_ratings entries, ascending."""
    urm_test = sps.csr_matrix(urm_test)
    return np.arange(urm_test.shape[0])[np.ediff1d(urm_test.indptr) >= min_ratings]


def finalize_count_metrics(counts, n_eval, cutoff, n_items, ignore_items=None):
    """get_metric_value() of the histogram-based objects (metrics.py:30-55,139-295,463-551)
    from the per-item recommendation counts of one cutoff.  ignore_items are never recommended (count 0):
    they shrink Coverage_Item's denominator (metrics.py:36-46) and are DELETED from the histogram before
    Herfindahl's sum (:213-217; the values are unchanged but numpy's pairwise summation groups a shorter array
    differently -- 1 ulp); the Gini / Shannon masks drop every zero-count item anyway (:163-167,267-271)."""
    counts = np.asarray(counts, dtype=np.float64)
    n_ignore = 0 if ignore_items is None else len(ignore_items)
    out = {}
    out["COVERAGE_ITEM"] = (counts > 0).sum() / (n_items - n_ignore)
    nz = counts[counts != 0]
    n = len(nz)
    srt = np.sort(nz)
    index = np.arange(1, n + 1)
    out["DIVERSITY_GINI"] = 2 * np.sum((n + 1 - index) / (n + 1) * srt / np.sum(srt))
    kept = counts if not n_ignore else np.delete(counts, np.asarray(ignore_items, dtype=np.int64))
    tot = kept.sum()
    out["DIVERSITY_HERFINDAHL"] = 1 - np.sum((kept / tot) ** 2) if tot != 0 else np.nan
    prob = nz / nz.sum()
    out["SHANNON_ENTROPY"] = -np.sum(prob * np.log2(prob))
    if n_eval == 0:
        out["DIVERSITY_MEAN_INTER_LIST"] = 1.0
    else:
        cooc = np.sum(counts ** 2) - n_eval * cutoff
        couples = n_eval ** 2 - n_eval
        out["DIVERSITY_MEAN_INTER_LIST"] = (couples - cooc / cutoff) / couples
    return out


def evaluate(score_fn, urm_train, urm_test, cutoff_list, exclude_seen=True, min_ratings=1,
             promotion="legacy", block_size=None, return_lists=False, ignore_items=None, ignore_users=None):
    """EvaluatorHoldout.evaluateRecommender (Evaluator.py:234-414).

    score_fn(user_id_array) -> float32 [n, n_items] raw scores (== _compute_item_score).
    Returns (results_dict {cutoff: {metric: value}}, n_users_evaluated[, all lists])."""
    legacy = promotion == "legacy"
    urm_train = sps.csr_matrix(urm_train)
    urm_test = sps.csr_matrix(urm_test)
    n_users, n_items = urm_test.shape
    users = users_to_evaluate(urm_test, min_ratings)
    n_ignore_users = 0
    if ignore_users is not None:                                         # Evaluator.py:171-176
        n_ignore_users = len(ignore_users)
        users = np.array(sorted(set(users.tolist()) - set(int(u) for u in ignore_users)), dtype=users.dtype)
    max_cutoff = max(cutoff_list)
    if block_size is None:
        block_size = min(1000, int(1e8 / n_items))

    pop = np.ediff1d(sps.csc_matrix(urm_train).indptr)                  # Novelty/AveragePopularity
    n_inter = pop.sum()
    pop_norm = pop / pop.max()

    # nep50: the running sum starts as a Python float and turns float32 at the first numpy
    # float32 addend, exactly as `results[...] += metric(...)` does under numpy >= 2
    f32_acc = (lambda: np.float64(0.0)) if legacy else (lambda: 0.0)
    acc = {c: {"ROC_AUC": f32_acc(), "PRECISION": f32_acc(), "PRECISION_RECALL_MIN_DEN": f32_acc(),
               "RECALL": f32_acc(), "NDCG": f32_acc(), "MAP": 0.0, "MRR": 0.0, "HIT_RATE": 0.0,
               "ARHR": 0.0, "RMSE": 0.0, "NOVELTY": 0.0, "AVERAGE_POPULARITY": 0.0,
               "counts": np.zeros(n_items, dtype=np.int64), "covered_users": 0}
           for c in cutoff_list}
    n_eval = 0
    all_lists = []
    for s in range(0, len(users), block_size):
        batch = users[s:s + block_size]
        raw = np.asarray(score_fn(batch), dtype=np.float32)
        if ignore_items is not None and len(ignore_items):
            # Evaluator.py:369-370 + BaseRecommender.py:103-106,210-211: -inf on the custom items
            raw = raw.copy()
            raw[:, np.asarray(ignore_items, dtype=np.int64)] = -np.inf
        lists, scores = recommend(raw, urm_train, batch, max_cutoff, exclude_seen)
        if return_lists:
            all_lists.extend(lists)
        for r, u in enumerate(batch):
            rel_items = urm_test.indices[urm_test.indptr[u]:urm_test.indptr[u + 1]]
            rel_ratings = urm_test.data[urm_test.indptr[u]:urm_test.indptr[u + 1]]
            user_rmse = rmse(scores[r], rel_items, rel_ratings)
            rec = np.asarray(lists[r], dtype=np.int64)
            is_rel = np.isin(rec, rel_items, assume_unique=True)
            n_eval += 1
            for c in cutoff_list:
                a = acc[c]
                ir, rc = is_rel[:c], rec[:c]
                a["ROC_AUC"] = a["ROC_AUC"] + roc_auc(ir, legacy)
                a["PRECISION"] = a["PRECISION"] + precision(ir, legacy)
                a["PRECISION_RECALL_MIN_DEN"] = a["PRECISION_RECALL_MIN_DEN"] + \
                    precision_recall_min_denominator(ir, len(rel_items), legacy)
                a["RECALL"] = a["RECALL"] + recall(ir, len(rel_items), legacy)
                a["NDCG"] = a["NDCG"] + ndcg(rc, rel_items, rel_ratings)
                a["HIT_RATE"] += ir.sum()
                a["ARHR"] += arhr(ir)
                a["RMSE"] += user_rmse
                a["MRR"] += rr(ir)
                a["MAP"] += average_precision(ir, len(rel_items))
                if len(rc) > 0:
                    p = pop[rc] / n_inter
                    p = p[p != 0]
                    a["NOVELTY"] += np.sum(-np.log2(p) / n_items)
                    a["AVERAGE_POPULARITY"] += np.sum(pop_norm[rc]) / len(rc)
                    a["counts"][rc] += 1
                    a["covered_users"] += 1

    results = {}
    for c in cutoff_list:
        a = acc[c]
        res = {}
        if n_eval > 0:
            for k in ("ROC_AUC", "PRECISION", "PRECISION_RECALL_MIN_DEN", "RECALL", "MAP", "MRR", "NDCG",
                      "HIT_RATE", "ARHR", "RMSE", "NOVELTY", "AVERAGE_POPULARITY"):
                res[k] = a[k] / n_eval
            res.update(finalize_count_metrics(a["counts"], n_eval, c, n_items, ignore_items))
            res["COVERAGE_USER"] = a["covered_users"] / (n_users - n_ignore_users)     # metrics.py:57-80
            p_, r_ = res["PRECISION"], res["RECALL"]
            res["F1"] = 2 * (p_ * r_) / (p_ + r_) if p_ + r_ != 0 else 0.0
        results[c] = {k: res[k] for k in METRIC_NAMES if k in res}
        results[c]["_sums"] = {k: a[k] for k in a if k != "counts"}
        results[c]["_counts"] = a["counts"]
    if return_lists:
        return results, n_eval, all_lists
    return results, n_eval


def get_result_string(results, n_decimals=7):
    """Evaluator.get_result_string (:95-110)."""
    out = ""
    for cutoff, res in results.items():
        out += "CUTOFF: {} - ".format(cutoff)
        for metric, v in res.items():
            if metric.startswith("_"):
                continue
            out += "{}: {:.{n}f}, ".format(metric, v, n=n_decimals)
        out += "\n"
    return out


# ------------------------------------------------------------------ EarlyStoppingScheduler
class EarlyStoppingOracle:
    """Utils_.py:25-88 state machine (model is any object with stop_fit/load_model/save_current_model)."""

    def __init__(self, model, evaluate_fn, metrics=("MAP",), freq=1, allow_worse=5, after=0):
        self.model, self.evaluate_fn = model, evaluate_fn
        self.metrics, self.freq, self.after = list(metrics), freq, after
        self.best = np.zeros(len(self.metrics))
        self.allow_worse = self.worse_left = allow_worse
        self.scores = []

    def __call__(self, epoch):
        if epoch > self.after and epoch % self.freq == 0:
            res = self.evaluate_fn(self.model)
            curr = np.array([res[5][m] for m in self.metrics])      # cutoff 5 is hard-coded (:64)
            self.scores.append(curr)
            if np.all(np.less_equal(curr, self.best)):
                if self.worse_left > 0:
                    self.worse_left -= 1
                else:
                    self.model.stop_fit()
                    self.model.load_model()
            else:
                self.best = curr
                self.worse_left = self.allow_worse
                self.model.save_current_model()
