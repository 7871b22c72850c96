"""Pins oracle/eval_oracle.py against the reference: golden runs of the unmodified reference
evaluator (tests/golden/make_golden.py), the reference's metrics_Test.py vectors, and the
surviving LastFM checkpoint with its stored test_results.pkl."""
import numpy as np
import pytest

from oracle import eval_oracle as eo
from tests.helpers import load_eval_fixture, load_lastfm_kat


@pytest.mark.parametrize("name", ["eval_small_implicit", "eval_small_ratings", "eval_small_shortlists",
                                  "eval_small_ignore", "eval_small_ignore_users"])
def test_oracle_matches_reference_run_bit_for_bit(name):
    fx = load_eval_fixture(name)
    res, n_eval, lists = eo.evaluate(lambda u: fx["scores"][u], fx["train"], fx["test"], fx["cutoffs"],
                                     promotion="nep50", return_lists=True, ignore_items=fx["ignore_items"],
                                     ignore_users=fx["ignore_users"])
    assert n_eval == len(fx["users"])
    for i, l in enumerate(lists):                      # same recommendation lists
        want = fx["lists"][i]
        assert l == want[want >= 0].tolist()
    for ci, c in enumerate(fx["cutoffs"]):
        for mi, m in enumerate(fx["metric_names"]):
            want, got = fx["results"][ci, mi], float(res[c][m])
            assert (np.isnan(want) and np.isnan(got)) or want == got, (c, m, want, got)


@pytest.mark.parametrize("name", ["eval_small_implicit", "eval_small_ratings", "eval_small_shortlists"])
def test_legacy_promotion_only_moves_last_bits(name):
    fx = load_eval_fixture(name)
    res, _ = eo.evaluate(lambda u: fx["scores"][u], fx["train"], fx["test"], fx["cutoffs"], promotion="legacy")
    for ci, c in enumerate(fx["cutoffs"]):
        for mi, m in enumerate(fx["metric_names"]):
            want, got = fx["results"][ci, mi], float(res[c][m])
            if np.isnan(want):
                assert np.isnan(got)
            else:
                assert got == pytest.approx(want, rel=2e-6, abs=1e-9), (c, m)


def test_lastfm_checkpoint_kat():
    """Score -> seen mask -> top-k -> metrics in item mode reproduces the reference's stored
    test_results.pkl (GANMF.py:288-290: predictions.transpose()[user_id_array])."""
    k = load_lastfm_kat()
    P, V = k["user_embeddings"], k["item_embeddings"]     # rows = items (17632), cols = users (1884)
    res, n_eval = eo.evaluate(lambda u: (P @ V[u].T).T.astype(np.float32), k["train"], k["test"], k["cutoffs"],
                              promotion="legacy")
    assert n_eval > 0
    for ci, c in enumerate(k["cutoffs"]):
        for mi, m in enumerate(k["metric_names"]):
            assert float(res[c][m]) == pytest.approx(k["results"][ci, mi], rel=1e-6, abs=1e-9), (c, m)


def test_reference_metric_vectors():
    """Base/Evaluation/metrics_Test.py:157-308 expected values."""
    pos = np.asarray([2, 4, 5, 10])
    l1, l2, l3 = np.asarray([1, 2, 3, 4, 5]), np.asarray([10, 5, 2, 4, 3]), np.asarray([1, 3, 6, 7, 8])
    l4 = np.asarray([11, 12, 13, 14, 15, 16, 2, 4, 5, 10])
    l5 = np.asarray([2, 11, 12, 13, 14, 15, 4, 5, 10, 16])
    rel = lambda l, p=pos: np.isin(l, p, assume_unique=True)
    assert np.allclose(eo.roc_auc(rel(l1, np.asarray([2, 4]))), (2. / 3 + 1. / 3) / 2)
    for legacy in (True, False):
        assert np.allclose(eo.recall(rel(l1), 4, legacy), 3. / 4)
        assert np.allclose(eo.recall(rel(l2), 4, legacy), 1.0)
        assert np.allclose(eo.recall(rel(l3), 4, legacy), 0.0)
        assert np.allclose(eo.precision(rel(l1), legacy), 3. / 5)
        assert np.allclose(eo.precision(rel(l2), legacy), 4. / 5)
        assert np.allclose(eo.precision(rel(l3), legacy), 0.0)
    assert np.allclose(eo.rr(rel(l1)), 1. / 2) and np.allclose(eo.rr(rel(l2)), 1.) and eo.rr(rel(l3)) == 0.0
    assert np.allclose(eo.average_precision(rel(l1), 4), (1. / 2 + 2. / 4 + 3. / 5) / 4)
    assert np.allclose(eo.average_precision(rel(l2), 4), 1.0)
    assert np.allclose(eo.average_precision(rel(l3), 4), 0.0)
    assert np.allclose(eo.average_precision(rel(l4), 4), (1. / 7 + 2. / 8 + 3. / 9 + 4. / 10) / 4)
    assert np.allclose(eo.average_precision(rel(l5), 4), (1. + 2. / 7 + 3. / 8 + 4. / 9) / 4)
    relv = np.asarray([5, 4, 3, 2])
    idcg = ((2 ** 5 - 1) / np.log(2) + (2 ** 4 - 1) / np.log(3) + (2 ** 3 - 1) / np.log(4) + (2 ** 2 - 1) / np.log(5))
    assert np.allclose(eo.dcg(np.sort(relv)[::-1]), idcg)
    assert np.allclose(eo.ndcg(l1, pos, relv),
                       ((2 ** 5 - 1) / np.log(3) + (2 ** 4 - 1) / np.log(5) + (2 ** 3 - 1) / np.log(6)) / idcg)
    assert np.allclose(eo.ndcg(l2, pos, relv),
                       ((2 ** 2 - 1) / np.log(2) + (2 ** 3 - 1) / np.log(3) + (2 ** 5 - 1) / np.log(4) +
                        (2 ** 4 - 1) / np.log(5)) / idcg)
    assert np.allclose(eo.ndcg(l3, pos, relv), 0.0)


def test_topk_ties_lowest_index_first():
    s = np.array([[1.0, 3.0, 3.0, -np.inf, 3.0, 0.5]], dtype=np.float32)
    idx, val = eo.topk_lowest_index(s, 4)
    assert idx.tolist() == [[1, 2, 4, 0]]
