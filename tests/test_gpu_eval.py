"""Evaluator parity on the B200.  Contract (north_star): given the same fp32 score matrix, top-k
item indices (ties -> lowest index) and metric sums are bit-exact against the oracle, which is itself
pinned to the unmodified reference evaluator (tests/test_oracle_eval.py)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import eval_oracle as eo
from tests.helpers import load_eval_fixture, load_lastfm_kat

pytestmark = pytest.mark.gpu

EXACT = ["PRECISION", "RECALL", "PRECISION_RECALL_MIN_DEN", "MAP", "NDCG", "MRR", "ROC_AUC", "HIT_RATE", "NOVELTY",
         "AVERAGE_POPULARITY"]


def scores_engine(n_users, n_items, train, test):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    e = Engine(L.KIND_GANMF, n_users, n_items, 1, emb_dim=1, max_batch=1)
    e.set_csr(L.CSR_SEEN, train, with_data=False)
    e.set_test(test, train)
    return e


@pytest.mark.parametrize("name", ["eval_small_implicit", "eval_small_ratings", "eval_small_shortlists"])
def test_mask_topk_and_metric_sums_bit_exact(name):
    fx = load_eval_fixture(name)
    n_users, n_items = fx["train"].shape
    users = fx["users"].astype(np.int32)
    K = max(fx["cutoffs"])
    e = scores_engine(n_users, n_items, fx["train"], fx["test"])
    scores = np.ascontiguousarray(fx["scores"][users])
    idx, val = e.mask_topk(scores, min(K, 128), users=users, remove_seen=True, write_back=True)
    # (1) same lists as the reference produced (golden), -1 marks dropped -inf entries
    want = fx["lists"][:, :idx.shape[1]]
    assert np.array_equal(idx, want)
    # (2) masked scores written back == reference's scores_batch
    assert np.array_equal(scores, eo.remove_seen(fx["scores"][users], fx["train"], users))
    # (3) metric sums: bit-exact vs the oracle in the reference's pinned-numpy arithmetic
    res, n_eval = eo.evaluate(lambda u: fx["scores"][u], fx["train"], fx["test"], fx["cutoffs"], promotion="legacy")
    sums, counts, per = e.metrics_from_topk(idx, users, fx["cutoffs"], want_per_user=True)
    from ganmf_b200._lib import MC_NAMES
    for ci, c in enumerate(fx["cutoffs"]):
        osum = res[c]["_sums"]
        for m in EXACT:
            assert sums[ci, MC_NAMES.index(m)] == float(osum[m]), (c, m, sums[ci, MC_NAMES.index(m)], float(osum[m]))
        assert sums[ci, MC_NAMES.index("ARHR")] == pytest.approx(float(osum["ARHR"]), rel=1e-14)
        assert sums[ci, MC_NAMES.index("COVERED")] == osum["covered_users"]
        assert np.array_equal(counts[ci], res[c]["_counts"])
    e.close()


def test_lastfm_checkpoint_kat_full_device_path():
    """Reference-trained factors (the surviving TF checkpoint) through the WHOLE device path in item
    mode: score GEMM -> seen mask -> top-k -> metrics == the reference's stored test_results.pkl."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    from ganmf_b200.GANRec.GANMF import GANMF
    k = load_lastfm_kat()
    rec = GANMF(k["train"], mode="item", is_experiment=True)
    rec.build(num_factors=k["num_factors"], emb_dim=k["emb_dim"])
    rec._build_engine(batch_size=8)
    rec._engine.set_param("generator/user_embeddings", k["user_embeddings"])
    rec._engine.set_param("generator/item_embeddings", k["item_embeddings"])
    rec._finish_fit()
    ev = EvaluatorHoldout(k["test"], cutoff_list=k["cutoffs"], exclude_seen=True)
    res, txt = ev.evaluateRecommender(rec)
    for ci, c in enumerate(k["cutoffs"]):
        for mi, m in enumerate(k["metric_names"]):
            assert float(res[c][m]) == pytest.approx(k["results"][ci, mi], rel=2e-5, abs=1e-9), (c, m)
    assert txt.startswith("CUTOFF: 5 - ROC_AUC: ")
    # and bit-exact sums against the oracle fed with the device's own scores
    eng = rec._engine
    ores, _ = eo.evaluate(lambda u: eng.score(u), k["train"], k["test"], k["cutoffs"], promotion="legacy")
    users = eo.users_to_evaluate(k["test"]).astype(np.int32)
    sums, counts = eng.evaluate(users, k["cutoffs"], remove_seen=True)
    for ci, c in enumerate(k["cutoffs"]):
        for m in EXACT:
            assert sums[ci, L.MC_NAMES.index(m)] == float(ores[c]["_sums"][m]), (c, m)
        assert sums[ci, L.MC_NAMES.index("RMSE")] == pytest.approx(float(ores[c]["_sums"]["RMSE"]), rel=1e-12)
        assert np.array_equal(counts[ci], ores[c]["_counts"])


def test_score_matches_oracle_both_modes():
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(0)
    for item_mode in (False, True):
        n_rows, width, kf = 700, 450, 33
        e = Engine(L.KIND_GANMF, n_rows, width, kf, emb_dim=8, max_batch=8, item_mode=item_mode)
        P = rs.standard_normal((n_rows, kf)).astype(np.float32)
        V = rs.standard_normal((width, kf)).astype(np.float32)
        e.set_param("generator/user_embeddings", P)
        e.set_param("generator/item_embeddings", V)
        n_users = width if item_mode else n_rows
        users = rs.permutation(n_users)[:100].astype(np.int32)
        got = e.score(users)
        full64 = P.astype(np.float64) @ V.astype(np.float64).T
        full64 = full64.T if item_mode else full64            # GANMF.py:288-292
        want = full64[users]
        assert got.shape == want.shape
        # split-TF32 scorer: fp32-accurate (a single-pass TF32 product would miss this by two orders of magnitude)
        assert np.max(np.abs(got - want)) <= 1e-5 * np.max(np.abs(want))
        e.close()


def test_large_scale_properties():
    """Full-size shape (cfg4 width, 27k items): size-independent properties of mask + top-k."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    rs = np.random.RandomState(3)
    n_users, n_items, kf, n = 4096, 27000, 64, 512
    train = sps.random(n_users, n_items, 0.005, format="csr", dtype=np.float32, random_state=rs)
    train.data[:] = 1
    e = Engine(L.KIND_GANMF, n_users, n_items, kf, emb_dim=8, max_batch=8)
    e.set_csr(L.CSR_SEEN, train, with_data=False)
    e.set_param("generator/user_embeddings", rs.standard_normal((n_users, kf)).astype(np.float32))
    e.set_param("generator/item_embeddings", rs.standard_normal((n_items, kf)).astype(np.float32))
    users = np.sort(rs.permutation(n_users)[:n]).astype(np.int32)
    idx, val, sc = e.recommend(users, 50, remove_seen=True, return_scores=True)
    assert np.all(np.diff(val, axis=1) <= 0)                              # sorted descending
    for r, u in enumerate(users):
        seen = train.indices[train.indptr[u]:train.indptr[u + 1]]
        assert not np.intersect1d(idx[r], seen).size                       # no seen item recommended
        assert np.all(np.isneginf(sc[r, seen]))
        assert np.array_equal(val[r], sc[r, idx[r]])                       # values are the masked scores
        assert val[r, -1] >= np.partition(sc[r], -50)[-50]                 # nothing better was left behind
    idx2, _, _ = e.recommend(users, 50, remove_seen=True)                   # idempotent
    assert np.array_equal(idx, idx2)
    e.close()


@pytest.mark.parametrize("name", ["eval_small_implicit", "eval_small_ratings", "eval_small_shortlists",
                                  "eval_small_ignore", "eval_small_ignore_users"])
def test_evaluator_with_foreign_recommender_matches_reference_run(name):
    """EvaluatorHoldout on a recommender that only exposes _compute_item_score / get_URM_train (as the
    reference's baselines do): all 19 metrics equal the unmodified reference evaluator's output (golden),
    up to the float32-vs-float64 running sums of numpy >= 2 (see oracle/eval_oracle.py), and are bit-exact
    against the oracle in the reference's pinned-numpy arithmetic."""
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    fx = load_eval_fixture(name)

    class Fixed(object):
        RECOMMENDER_NAME = "Fixed"

        def get_URM_train(self):
            return fx["train"].copy()

        def _compute_item_score(self, user_id_array, items_to_compute=None):
            return fx["scores"][user_id_array].copy()

    # (eval_small_ignore: the reference run used EvaluatorHoldout(ignore_items=40 of 320 items))
    ev = EvaluatorHoldout(fx["test"], cutoff_list=fx["cutoffs"], exclude_seen=True, ignore_items=fx["ignore_items"],
                          ignore_users=fx["ignore_users"])
    res, txt = ev.evaluateRecommender(Fixed())
    ores, n_eval = eo.evaluate(lambda u: fx["scores"][u], fx["train"], fx["test"], fx["cutoffs"], promotion="legacy",
                               ignore_items=fx["ignore_items"], ignore_users=fx["ignore_users"])
    for ci, c in enumerate(fx["cutoffs"]):
        for mi, m in enumerate(fx["metric_names"]):
            want = fx["results"][ci, mi]
            got = float(res[c][m])
            if np.isnan(want):
                assert np.isnan(got)
                continue
            assert got == pytest.approx(want, rel=2e-6, abs=1e-9), (c, m)
            if m in ("ARHR", "RMSE"):          # BLAS ddot / float32 pairwise sum orders: equal to ~1 ulp
                assert got == pytest.approx(float(ores[c][m]), rel=1e-12)
            else:
                assert got == float(ores[c][m]), (c, m, got, float(ores[c][m]))
