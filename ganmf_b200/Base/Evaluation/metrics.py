"""Host-side finishing of the evaluator metrics (reference: Base/Evaluation/metrics.py).

The per-user arithmetic (precision :612, recall :662, average_precision :681, ndcg/dcg :693-722,
rr :670, arhr :596, roc_auc :576, rmse :634, Novelty :298, AveragePopularity :355) runs in
csrc/eval_kernels.cuh.  What is left here are the closed-form values of the histogram-based
objects, computed once per cutoff from the integer per-item recommendation counts the device
returns: Coverage_Item :30, Gini_Diversity :139, Diversity_Herfindahl :188, Shannon_Entropy :235,
Diversity_MeanInterList :463."""
import numpy as np


def finalize_count_metrics(counts, n_eval, cutoff, n_items, ignore_items=None):
    """ignore_items are never recommended (count 0): they shrink Coverage_Item's denominator (metrics.py:36-46)
    and are deleted from the histogram Herfindahl sums over (:213-217: same values, but numpy's pairwise
    summation of the shorter array can differ in the last bit); Gini / Shannon drop zero-count items anyway."""
    counts = np.asarray(counts, dtype=np.float64)
    n_ignore = 0 if ignore_items is None else len(ignore_items)
    out = {}
    out["COVERAGE_ITEM"] = (counts > 0).sum() / (n_items - n_ignore)
    nz = counts[counts != 0]
    n = len(nz)
    srt = np.sort(nz)
    index = np.arange(1, n + 1)
    out["DIVERSITY_GINI"] = 2 * np.sum((n + 1 - index) / (n + 1) * srt / np.sum(srt)) if n else 0.0
    kept = counts if not n_ignore else np.delete(counts, np.asarray(ignore_items, dtype=np.int64))
    tot = kept.sum()
    out["DIVERSITY_HERFINDAHL"] = 1 - np.sum((kept / tot) ** 2) if tot != 0 else np.nan
    prob = nz / nz.sum() if n else nz
    out["SHANNON_ENTROPY"] = -np.sum(prob * np.log2(prob)) if n else 0.0
    if n_eval == 0:
        out["DIVERSITY_MEAN_INTER_LIST"] = 1.0
    else:
        cooc = np.sum(counts ** 2) - n_eval * cutoff
        couples = n_eval ** 2 - n_eval
        out["DIVERSITY_MEAN_INTER_LIST"] = (couples - cooc / cutoff) / couples if couples else 1.0
    return out
