"""The host side of GANMF / DisGANMF .fit() (ganmf_b200/GANRec/_gan_base.py) on the CPU: a stand-in replaces the
device engine (which cannot exist without a GPU), everything else is the product code -- minibatch id stream,
reference return values, EarlyStoppingScheduler wiring, the item-mode orientation of URM_train around evaluations,
saveModel / loadModel files.  Reference: GANRec/GANMF.py:142-244,246-255,309-342, Utils_.py:25-88."""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sps

from oracle import train_oracle as to


class StandInEngine(object):
    """Records what the host asks of the device."""
    instances = []

    def __init__(self, kind, n_rows, width, max_batch=32, item_mode=False, device=0, gemm_path=0, **kw):
        self.kind, self.n_rows, self.width, self.max_batch, self.item_mode, self.kw = kind, n_rows, width, max_batch, \
            item_mode, kw
        self.lib = object()
        self.epochs, self.snapshots, self.restores, self.csr = [], 0, 0, {}
        self.p = {"generator/user_embeddings": np.zeros((n_rows, kw["num_factors"]), np.float32),
                  "generator/item_embeddings": np.ones((width, kw["num_factors"]), np.float32)}
        StandInEngine.instances.append(self)

    def set_csr(self, which, m, with_data=True):
        self.csr[which] = sps.csr_matrix(m).shape

    def set_csr_transposed(self, which, m, with_data=True):      # the device holds m.T
        self.csr[which] = sps.csr_matrix(m).shape[::-1]

    def init_params(self, seed):
        self.seed = seed

    def param_infos(self):
        return [("autoencoder/encoding/kernel", self.width, 4, 0), ("generator/user_embeddings", self.n_rows, 2, 1),
                ("generator/item_embeddings", self.width, 2, 1)]

    def train_epoch(self, perm, batch, d_steps, g_steps, *hp):
        self.epochs.append((np.array(perm).copy(), batch, d_steps, g_steps, hp))
        nb = (len(perm) + batch - 1) // batch
        e = len(self.epochs)
        return np.full(nb * d_steps, 1.0 / e, np.float32), np.full(nb * g_steps, 2.0 / e, np.float32)

    def snapshot(self):
        self.snapshots += 1

    def restore(self):
        self.restores += 1

    def get_params(self):
        return dict(self.p)

    def set_params(self, p):
        self.p = {k: np.array(v) for k, v in p.items()}

    def close(self):
        pass


class SeqEvaluator(object):
    """validation_evaluator stand-in: MAP@5 follows a script; records the orientation of URM_train it sees."""

    def __init__(self, seq):
        self.seq, self.i, self.shapes = list(seq), 0, []

    def evaluateRecommender(self, model):
        self.shapes.append(model.URM_train.shape)
        v = self.seq[min(self.i, len(self.seq) - 1)]
        self.i += 1
        return {5: {"MAP": v}}, "MAP@5 %.3f" % v


@pytest.fixture()
def stand_in(monkeypatch):
    from ganmf_b200.GANRec import _gan_base
    StandInEngine.instances = []
    monkeypatch.setattr(_gan_base, "Engine", StandInEngine)
    return StandInEngine


def urm(n_users=37, n_items=53):
    m = sps.random(n_users, n_items, 0.1, format="csr", dtype=np.float32, random_state=np.random.RandomState(0))
    m.data[:] = 1.0
    return m


def test_fit_feeds_the_reference_minibatch_stream_and_returns_epochs_plus_one(stand_in):
    from ganmf_b200.GANRec.GANMF import GANMF
    np.random.seed(1337)                                      # RunBestParameters.py:81
    rec = GANMF(urm(), mode="user", seed=7, is_experiment=True)
    last = rec.fit(num_factors=2, emb_dim=4, epochs=5, batch_size=16, d_lr=1e-3, g_lr=2e-3, d_steps=2, g_steps=3,
                   d_reg=1e-4, g_reg=0.0, m=10, recon_coefficient=0.05)
    assert last == 6                                          # GANMF.py:244: epochs + 1 when never stopped
    eng = stand_in.instances[-1]
    assert (eng.n_rows, eng.width, eng.max_batch, eng.item_mode, eng.seed) == (37, 53, 16, False, 7)
    assert eng.kw == {"num_factors": 2, "emb_dim": 4}
    # one cumulative in-place shuffle of arange(num_users) per epoch from numpy's global stream (GANMF.py:156,175)
    want = [np.concatenate(b) for _, b in to.epoch_index_stream(37, 16, 5, seed=1337)]
    assert len(eng.epochs) == 5
    for (perm, batch, d_steps, g_steps, hp), w in zip(eng.epochs, want):
        assert np.array_equal(perm, w) and (batch, d_steps, g_steps) == (16, 2, 3)
        assert hp == (1e-3, 2e-3, 1e-4, 0.0, 10.0, 0.05)
    assert rec.train_d_loss == pytest.approx([1.0, 0.5, 1 / 3, 0.25, 0.2])       # per-epoch means (GANMF.py:205-209)
    assert rec.config["epochs"] == 5 and rec.config["m"] == 10 and "self" not in rec.config
    assert rec.params == {"D": ["autoencoder/encoding/kernel"],
                          "G": ["generator/user_embeddings", "generator/item_embeddings"]}


def test_early_stopping_return_value_and_snapshot_calls(stand_in):
    from ganmf_b200.GANRec.GANMF import GANMF
    rec = GANMF(urm(), mode="user", is_experiment=True)
    ev = SeqEvaluator([0.10, 0.20, 0.15, 0.15, 0.12, 0.30])    # evaluated at epochs 2, 4, 6, 8, 10
    last = rec.fit(num_factors=2, emb_dim=4, epochs=50, batch_size=16, allow_worse=2, freq=2, after=0,
                   metrics=["MAP"], validation_evaluator=ev, validation_set=None, sample_every=None)
    eng = stand_in.instances[-1]
    # the initial weights are snapshot once (the reference's shadow variables exist from graph construction,
    # GANMF.py:123-128); improvements at epochs 2 and 4 (two more); worse at 6, 8 (tolerated), 10 -> stop + restore
    assert (eng.snapshots, eng.restores) == (3, 1)
    assert last == 10 and len(eng.epochs) == 10                # GANMF.py:244: the epoch it stopped at
    # RecSysExp.py:274-276 turns that into the epoch count of the best model
    assert last - 2 * 2 == 6


@pytest.mark.parametrize("algo", ["GANMF", "DisGANMF"])
def test_item_mode_orientation_around_evaluations(stand_in, algo):
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.GANRec.GANMF import GANMF
    m = urm()
    cls = GANMF if algo == "GANMF" else DisGANMF
    rec = cls(m, mode="item", is_experiment=True)
    assert rec.URM_train.shape == (53, 37) and (rec.num_users, rec.num_items) == (53, 37)   # GANMF.py:31-36
    ev = SeqEvaluator([0.1, 0.2, 0.3])
    kw = dict(emb_dim=4) if algo == "GANMF" else dict(d_layers=1, d_nodes=4)
    rec.fit(num_factors=2, epochs=3, batch_size=8, allow_worse=5, freq=1, validation_evaluator=ev, **kw)
    eng = stand_in.instances[-1]
    assert (eng.n_rows, eng.width, eng.item_mode) == (53, 37, True)     # the model's rows are the items
    assert all(len(p[0]) == 53 for p in eng.epochs)
    assert ev.shapes == [(37, 53)] * 3                                   # users x items while evaluating (:215-228)
    assert rec.URM_train.shape == (37, 53)                               # flipped back at the end (:241-242)


def test_save_and_load_model_files(stand_in, tmp_path):
    from ganmf_b200.GANRec.GANMF import GANMF
    rec = GANMF(urm(), mode="item", is_experiment=True)
    rec.fit(num_factors=2, emb_dim=4, epochs=1, batch_size=8)
    eng = stand_in.instances[-1]
    eng.p["generator/user_embeddings"][:] = 3.5
    rec.saveModel(str(tmp_path))
    assert pickle.load(open(os.path.join(str(tmp_path), "build_params.pkl"), "rb")) == {"num_factors": 2, "emb_dim": 4}
    assert os.path.exists(os.path.join(str(tmp_path), "GANMF_item.npz"))          # GANMF.py:313: RECOMMENDER_NAME_mode
    rec2 = GANMF(urm(), mode="item", is_experiment=True)
    rec2.loadModel(str(tmp_path))
    eng2 = stand_in.instances[-1]
    assert eng2 is not eng and np.all(eng2.p["generator/user_embeddings"] == 3.5)
    assert rec2.URM_train.shape == (53, 37)                              # left transposed after loadModel (A.4)
    rec2.save_model(str(tmp_path), "alias")                              # north_star's snake-case aliases
    rec3 = GANMF(urm(), mode="item", is_experiment=True)
    rec3.load_model_from(str(tmp_path), "alias")
    assert np.all(stand_in.instances[-1].p["generator/user_embeddings"] == 3.5)
    with pytest.raises(IOError):
        rec3.loadModel(str(tmp_path), "missing")


def test_constructor_contract(stand_in):
    from ganmf_b200.GANRec.DisGANMF import DisGANMF
    from ganmf_b200.GANRec.GANMF import GANMF
    with pytest.raises(ValueError):
        GANMF(urm(), mode="both")                                        # GANMF.py:28-29
    g = GANMF(urm(), "user", False, 99, True)                            # (URM, mode, verbose, seed, is_experiment)
    d = DisGANMF(urm(), "user", 99, False, True)                         # (URM, mode, seed, verbose, is_experiment)
    assert (g.seed, g.verbose, d.seed, d.verbose) == (99, False, 99, False)
    assert (g.RECOMMENDER_NAME, d.RECOMMENDER_NAME) == ("GANMF", "DisGANMF")
    with pytest.raises(TypeError):
        d.saveModel("/tmp/x")                                            # DisGANMF.py:264: file_name is required


# ------------------------------------------------------------------------------------------------ evaluator host side
class SumsEngine(object):
    """Stand-in for the device evaluation stage: hands back the per-cutoff metric SUMS and item histograms the
    oracle accumulates (in the device's column order), so the test isolates what the host does with them."""

    def __init__(self, ores, cutoffs, n_items):
        from ganmf_b200 import _lib as L
        self.sums = np.zeros((len(cutoffs), L.MC_NCOL))
        self.counts = np.zeros((len(cutoffs), n_items), dtype=np.int64)
        for ci, c in enumerate(cutoffs):
            s = dict(ores[c]["_sums"])
            s["COVERED"] = s.pop("covered_users")
            for mi, name in enumerate(L.MC_NAMES):
                self.sums[ci, mi] = float(s[name])
            self.counts[ci] = ores[c]["_counts"]
        self.calls = []

    def set_test(self, test, train):
        self.calls.append("set_test")

    def evaluate(self, users, cutoffs, remove_seen=True):
        self.calls.append(("evaluate", len(users), list(cutoffs), remove_seen))
        return self.sums, self.counts


@pytest.mark.parametrize("name", ["eval_small_implicit", "eval_small_ratings", "eval_small_shortlists",
                                  "eval_small_ignore_users"])
def test_evaluator_host_side_reproduces_the_reference_results_dict(name):
    """EvaluatorHoldout.evaluateRecommender: users to evaluate, averaging, F1 of the averaged P and R,
    COVERAGE_USER, the histogram metrics, key order and the 7-decimal result string (Evaluator.py:95-110,
    119-179,386-414) -- against the golden run of the unmodified reference evaluator."""
    from oracle import eval_oracle as eo
    from tests.helpers import load_eval_fixture
    from ganmf_b200.Base.Evaluation.Evaluator import EvaluatorHoldout
    fx = load_eval_fixture(name)
    # "legacy" = the reference's pinned numpy 1.16 promotion rules, which the device and the host implement;
    # the golden run was made under numpy >= 2, where a few float32 scalars stay float32 (oracle/eval_oracle.py)
    ores, n_eval = eo.evaluate(lambda u: fx["scores"][u], fx["train"], fx["test"], fx["cutoffs"], promotion="legacy",
                               ignore_users=fx["ignore_users"])

    class Rec(object):
        _engine = SumsEngine(ores, fx["cutoffs"], fx["train"].shape[1])

        def get_URM_train(self):
            return fx["train"].copy()

    ev = EvaluatorHoldout(fx["test"], cutoff_list=fx["cutoffs"], exclude_seen=True, ignore_users=fx["ignore_users"])
    assert np.array_equal(np.asarray(ev.usersToEvaluate), fx["users"])               # Evaluator.py:151-176
    res, txt = ev.evaluateRecommender(Rec())
    assert Rec._engine.calls == ["set_test", ("evaluate", len(fx["users"]), fx["cutoffs"], True)]
    for ci, c in enumerate(fx["cutoffs"]):
        assert sorted(res[c].keys()) == fx["metric_names"]
        for mi, m in enumerate(fx["metric_names"]):
            want, got = fx["results"][ci, mi], float(res[c][m])
            if np.isnan(want):
                assert np.isnan(got)
                continue
            assert got == float(ores[c][m]), (c, m, got, float(ores[c][m]))         # bit for bit in legacy arithmetic
            assert got == pytest.approx(want, rel=2e-6, abs=1e-9), (c, m, got, want)  # golden: last bits only
    assert txt == eo.get_result_string({c: {k: ores[c][k] for k in res[c]} for c in fx["cutoffs"]})
    assert txt.startswith("CUTOFF: %d - ROC_AUC: " % fx["cutoffs"][0]) and txt.count("\n") == len(fx["cutoffs"])
