"""Training parity on the B200: identical initial weights and minibatch index stream, per-step
D/G losses and parameters after N steps against the fp32 oracle (oracle/train_oracle.py).
Tolerance (north_star): relative <= 1e-3 on every loss and on ||theta - theta_ref|| / ||theta_ref||
per tensor (TF32 tensor-core inputs, fp32 accumulation)."""
import numpy as np
import pytest
import scipy.sparse as sps

from oracle import train_oracle as to

pytestmark = pytest.mark.gpu
REL = 1e-3


def make_urm(n_rows, width, density, seed):
    rs = np.random.RandomState(seed)
    m = sps.random(n_rows, width, density, format="csr", dtype=np.float32, random_state=rs)
    m.data[:] = 1.0
    return m


def rel_err(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) /
                 max(np.linalg.norm(b.astype(np.float64)), 1e-30))


def run_ganmf(n_rows, width, k, E, B, epochs, hp, gemm_path, density=0.05, seed=0, explicit=False):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    urm = make_urm(n_rows, width, density, seed)
    if explicit:                                        # ratings 1..5 instead of implicit ones
        urm.data[:] = np.random.RandomState(seed + 3).randint(1, 6, size=urm.nnz).astype(np.float32)
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=seed + 1)
    rs = np.random.RandomState(seed + 2)
    p0["autoencoder/encoding/bias"] = (rs.standard_normal(E) * 0.01).astype(np.float32)
    p0["autoencoder/decoding/bias"] = (rs.standard_normal(width) * 0.01).astype(np.float32)
    eng = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B, gemm_path=gemm_path)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    dl_all, gl_all, odl, ogl = [], [], [], []
    for _, batches in to.epoch_index_stream(n_rows, B, epochs, seed=1337):
        perm = np.concatenate(batches)
        dl, gl = eng.train_epoch(perm, B, 1, 1, hp["d_lr"], hp["g_lr"], hp["d_reg"], hp["g_reg"], hp["m"],
                                 hp["alpha"])
        dl_all += list(dl)
        gl_all += list(gl)
        for b in batches:
            odl.append(orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"], m=hp["m"]))
        for b in batches:
            ogl.append(orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"]))
    got = eng.get_params()
    eng.close()
    return np.array(dl_all), np.array(gl_all), np.array(odl), np.array(ogl), got, orc.p


HP = dict(d_lr=1e-4, g_lr=2e-4, d_reg=1e-4, g_reg=0.0, m=10.0, alpha=0.01)


@pytest.mark.parametrize("path_name", ["simt", "tc", "tc-dense"])
@pytest.mark.parametrize("hp", [HP, dict(HP, m=0.05, g_reg=1e-3, alpha=0.3)])     # gate open / closed
def test_ganmf_100_steps_parity(monkeypatch, path_name, hp):
    from ganmf_b200 import _lib as L
    if path_name == "tc-dense":            # the reference graph's own GEMM list (no low-rank generator route)
        monkeypatch.setenv("GANMF_LOWRANK", "0")
        path_name = "tc"
    path = {"simt": L.GEMM_SIMT, "tc": L.GEMM_TC}[path_name]
    # 300 rows, B=64 -> 5 batches/epoch (last one short: 44 rows); 10 epochs = 50 D + 50 G steps
    dl, gl, odl, ogl, got, want = run_ganmf(300, 517, 24, 40, 64, 10, hp, path)
    assert len(dl) == 50 and len(gl) == 50
    np.testing.assert_allclose(dl, odl, rtol=REL)
    np.testing.assert_allclose(gl, ogl, rtol=REL)
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))


@pytest.mark.parametrize("hp", [HP, dict(HP, m=0.05, alpha=0.3)])                  # gate open / closed
@pytest.mark.parametrize("lowrank", ["0", "1"])       # fake profiles through the dense GEMMs / through V^T.We (capi.cu)
def test_ganmf_steps_parity_on_cta_pairs(monkeypatch, hp, lowrank):
    """GANMF_PAIR=2 forces every legal tcgen05 GEMM onto CTA pairs (cta_group::2), which the benchmark shapes
    use for all many-tile GEMMs: residual epilogue with prefetched addend + bias + energy sums (2B=320 rows: the
    second CTA of the last pair tile is partly out of range; 517 columns: ragged last tile), the fused-Adam
    epilogues of dWd / dWe, the plain generator GEMMs.  Same tolerance as the single-CTA paths."""
    from ganmf_b200 import _lib as L
    monkeypatch.setenv("GANMF_PAIR", "2")
    monkeypatch.setenv("GANMF_LOWRANK", lowrank)
    dl, gl, odl, ogl, got, want = run_ganmf(600, 517, 24, 300, 160, 8, hp, L.GEMM_TC)
    assert len(dl) == 32 and len(gl) == 32
    np.testing.assert_allclose(dl, odl, rtol=REL)
    np.testing.assert_allclose(gl, ogl, rtol=REL)
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))


@pytest.mark.parametrize("pair", ["1", "2"])
@pytest.mark.parametrize("hp,explicit", [(HP, False), (dict(HP, m=0.05, alpha=0.3), True)])   # gate open / closed
def test_ganmf_steps_parity_sparse_real_route(monkeypatch, pair, hp, explicit):
    """GANMF_SPARSE_REAL=1 forces the codes of the real rows through the CSR gather-sum
    (csr_encode_rows_kernel, exact fp32) and only the fake rows through the tensor cores -- the route the engine
    takes by itself up to 0.55 % density (cfg4, cfg5).  Same contract as the dense route; ratings other than 1 exercise
    the value column of the CSR; 44-row last batch = ragged fake-half GEMM."""
    from ganmf_b200 import _lib as L
    monkeypatch.setenv("GANMF_SPARSE_REAL", "1")
    monkeypatch.setenv("GANMF_PAIR", pair)
    dl, gl, odl, ogl, got, want = run_ganmf(600, 517, 24, 300, 160, 8, hp, L.GEMM_TC, explicit=explicit)
    assert len(dl) == 32 and len(gl) == 32
    np.testing.assert_allclose(dl, odl, rtol=REL)
    np.testing.assert_allclose(gl, ogl, rtol=REL)
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))


def test_ganmf_sparse_route_is_chosen_by_density_and_matches_dense(monkeypatch):
    """At 0.2 % density the engine takes the sparse route by itself: fewer tensor-core flops are launched than with
    GANMF_SPARSE_REAL=0, and both runs stay within the oracle tolerance of each other."""
    from ganmf_b200 import _lib as L
    res = {}
    for mode in ("auto", "0"):
        if mode == "auto":
            monkeypatch.delenv("GANMF_SPARSE_REAL", raising=False)
        else:
            monkeypatch.setenv("GANMF_SPARSE_REAL", mode)
        res[mode] = run_ganmf(512, 4096, 24, 64, 128, 2, HP, L.GEMM_TC, density=0.002)
    for mode in res:
        dl, gl, odl, ogl, got, want = res[mode]
        np.testing.assert_allclose(dl, odl, rtol=REL)
        np.testing.assert_allclose(gl, ogl, rtol=REL)
        for n in want:
            assert rel_err(got[n], want[n]) <= REL, (mode, n, rel_err(got[n], want[n]))
    # the two routes differ in rounding (exact fp32 vs TF32 on the real codes), so they are not bit-identical
    assert any(not np.array_equal(res["auto"][4][n], res["0"][4][n]) for n in res["0"][4])


def test_ganmf_steps_parity_resident_generator_gemm(monkeypatch):
    """GANMF_GEN_RESIDENT=1: the fake profiles through the resident-A generator kernel (gen_gemm.cuh), which the engine
    takes by itself from 256 minibatch rows on; 160-row batches = one partly filled pair tile, 44-row last batch."""
    from ganmf_b200 import _lib as L
    monkeypatch.setenv("GANMF_GEN_RESIDENT", "1")
    dl, gl, odl, ogl, got, want = run_ganmf(600, 517, 24, 300, 160, 4, HP, L.GEMM_TC)
    np.testing.assert_allclose(dl, odl, rtol=REL)
    np.testing.assert_allclose(gl, ogl, rtol=REL)
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))


def test_ganmf_decoder_bias_grad_from_the_colsum_pass(monkeypatch):
    """GANMF_COLPART=0: dbd from the pass over the residual instead of the residual GEMM's per-32-row column sums
    (the default whenever the real / fake boundary falls on a 32-row group, i.e. in every other test here)."""
    from ganmf_b200 import _lib as L
    monkeypatch.setenv("GANMF_COLPART", "0")
    dl, gl, odl, ogl, got, want = run_ganmf(600, 517, 24, 300, 160, 4, HP, L.GEMM_TC)
    np.testing.assert_allclose(dl, odl, rtol=REL)
    np.testing.assert_allclose(gl, ogl, rtol=REL)
    for n in want:
        assert rel_err(got[n], want[n]) <= REL, (n, rel_err(got[n], want[n]))


def test_ganmf_ml1m_shape_steps_parity():
    """cfg1 shape (GANMF-u ML-1M: I=3706, k=250, E=992, B=64) on the committed split, best params."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    from tests.helpers import load_split
    urm = load_split("Movielens1M")["train"]
    n_rows, width = urm.shape
    k, E, B = 250, 992, 64
    hp = dict(d_lr=1e-4, g_lr=1.653241474168571e-4, d_reg=1e-4, g_reg=0.0, m=10.0, alpha=0.01)
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=5)
    eng = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    _, batches = next(iter(to.epoch_index_stream(n_rows, B, 1, seed=1337)))
    batches = batches[:12] + [batches[-1]]              # 13 batches incl. the short last one (24 rows)
    perm = np.concatenate(batches)
    eng.upload_ids(perm)
    off, slot = 0, 0
    for b in batches:
        eng.d_step(off, len(b), hp["d_lr"], hp["d_reg"], hp["m"], loss_slot=slot)
        off += len(b)
        slot += 1
    off = 0
    for b in batches:
        eng.g_step(off, len(b), hp["g_lr"], hp["g_reg"], hp["alpha"], loss_slot=slot)
        off += len(b)
        slot += 1
    losses = eng.read_losses(slot)
    want = [orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"], m=hp["m"]) for b in batches]
    want += [orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"])
             for b in batches]
    np.testing.assert_allclose(losses, want, rtol=REL)
    got = eng.get_params()
    for n in orc.p:
        assert rel_err(got[n], orc.p[n]) <= REL, (n, rel_err(got[n], orc.p[n]))
    eng.close()


@pytest.mark.parametrize("act", ["tanh", "relu"])
def test_disganmf_cta_pairs_equal_single_ctas(monkeypatch, act):
    """Wide DisGANMF layers (300 nodes, 2B = 192 rows) with every legal GEMM forced onto CTA pairs (activation
    epilogue, rank-1 id term, weight gradients under cta_group::2) against the same run on single-CTA tiles:
    the two tilings accumulate each output element over the same K blocks in the same order, so losses and
    weights must agree to fp32 round-off (this wide net at d_lr = 1e-3 drifts ~1 % from the fp32 oracle under
    TF32 on EITHER tiling, so the oracle comparison lives in the narrower configurations below)."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_rows, width, k, B, layers, nodes = 400, 310, 12, 96, 2, 300
    urm = make_urm(n_rows, width, 0.06, 3)
    p0 = to.init_disganmf_params(n_rows, width, k, layers, nodes, seed=9)
    hp = dict(d_lr=1e-3, g_lr=2e-4, d_reg=1e-5, g_reg=1e-4, alpha=0.3)
    outs = []
    for pair in ("2", "0"):
        monkeypatch.setenv("GANMF_PAIR", pair)
        eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=layers, d_nodes=nodes, d_act=act, max_batch=B)
        eng.set_csr(L.CSR_TRAIN, urm)
        eng.set_params(p0)
        eng.reset_optimizers()
        losses = []
        for _, batches in to.epoch_index_stream(n_rows, B, 3, seed=1337):
            dl, gl = eng.train_epoch(np.concatenate(batches), B, 1, 1, hp["d_lr"], hp["g_lr"], hp["d_reg"], hp["g_reg"],
                                     1.0, hp["alpha"])
            losses += list(dl) + list(gl)
        outs.append((np.array(losses), eng.get_params()))
        eng.close()
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=2e-5)
    for n in outs[0][1]:
        assert rel_err(outs[0][1][n], outs[1][1][n]) < 2e-5, n


# (4 and 5 layers: 10 / 12 discriminator tensors, more than one fused-Adam launch holds -- regression for an overflow of
#  the optimiser's segment table that left the last tensors of a 4-layer net without updates)
@pytest.mark.parametrize("act,layers,nodes", [("linear", 1, 4), ("tanh", 2, 48), ("relu", 3, 33), ("sigmoid", 2, 130),
                                              ("tanh", 4, 24), ("linear", 5, 16)])
def test_disganmf_steps_parity(act, layers, nodes):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_rows, width, k, B = 200, 310, 12, 32
    urm = make_urm(n_rows, width, 0.06, 3)
    p0 = to.init_disganmf_params(n_rows, width, k, layers, nodes, seed=9)
    hp = dict(d_lr=1e-3, g_lr=2e-4, d_reg=1e-5, g_reg=1e-4, alpha=0.3)
    eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=layers, d_nodes=nodes, d_act=act, max_batch=B)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    orc = to.DisGanmfOracle(p0, layers, act, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    dl_all, gl_all, odl, ogl = [], [], [], []
    for _, batches in to.epoch_index_stream(n_rows, B, 4, seed=1337):
        perm = np.concatenate(batches)
        dl, gl = eng.train_epoch(perm, B, 1, 1, hp["d_lr"], hp["g_lr"], hp["d_reg"], hp["g_reg"], 1.0, hp["alpha"])
        dl_all += list(dl)
        gl_all += list(gl)
        for b in batches:
            odl.append(orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"]))
        for b in batches:
            ogl.append(orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"]))
    # ids up to 200 are an input FEATURE of the discriminator (DisGANMF.py:110): TF32 keeps them exact
    np.testing.assert_allclose(dl_all, odl, rtol=REL)
    np.testing.assert_allclose(gl_all, ogl, rtol=REL)
    got = eng.get_params()
    for n in orc.p:
        assert rel_err(got[n], orc.p[n]) <= REL, (n, rel_err(got[n], orc.p[n]))
    eng.close()


def test_phased_and_ranged_updates_equal_the_fused_step():
    """The data-parallel building blocks (phased backward, Adam on slab ranges) on one GPU reproduce the
    single-call step: same losses, same weights (they are different kernels/fusions of the same math)."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_rows, width, k, E, B = 256, 390, 16, 40, 64
    urm = make_urm(n_rows, width, 0.05, 11)
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=2)
    hp = dict(d_lr=1e-3, g_lr=1e-3, d_reg=1e-4, g_reg=1e-4, m=10.0, alpha=0.2)
    ids = np.random.RandomState(0).permutation(n_rows)[:3 * B].astype(np.int32)
    outs = []
    for variant in ("fused", "phased", "ranges"):
        eng = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B)
        eng.set_csr(L.CSR_TRAIN, urm)
        eng.set_params(p0)
        eng.reset_optimizers()
        eng.upload_ids(ids)
        n_d = sum(r * ((c + 31) // 32 * 32) for _, r, c, g in eng.param_infos() if not g)
        for s in range(3):
            off = s * B
            if variant == "fused":
                eng.d_step(off, B, hp["d_lr"], hp["d_reg"], hp["m"], loss_slot=2 * s)
            else:
                eng.d_forward_phase(off, B, 1)
                eng.d_forward_phase(off, B, 2)
                eng.d_backward_phase(B, B, hp["m"], 1)
                eng.d_backward_phase(B, B, hp["m"], 3)
                eng.d_backward_phase(B, B, hp["m"], 4)
                if variant == "phased":
                    eng.d_apply(hp["d_lr"], hp["d_reg"], 2 * s)
                else:
                    half = (n_d // 2) // 4 * 4
                    eng.d_apply_ranges(hp["d_lr"], hp["d_reg"], [0], [half], new_step=True)
                    eng.d_apply_ranges(hp["d_lr"], hp["d_reg"], [half], [n_d - half], new_step=False)
                    eng.finalize_loss(hp["d_reg"], 2 * s)
            if variant == "fused":
                eng.g_step(off, B, hp["g_lr"], hp["g_reg"], hp["alpha"], loss_slot=2 * s + 1)
            else:                                       # the two-part G backward of the data-parallel trainer
                eng.g_forward_backward_part(off, B, B, hp["alpha"], 1)
                eng.g_forward_backward_part(off, B, B, hp["alpha"], 2)
                eng.g_apply(B, B, hp["g_lr"], hp["g_reg"], hp["alpha"], 2 * s + 1)
        outs.append((eng.read_losses(6), eng.get_params()))
        eng.close()
    for losses, params in outs[1:]:
        np.testing.assert_allclose(losses, outs[0][0], rtol=2e-5)
        for n in params:
            assert rel_err(params[n], outs[0][1][n]) < 2e-5, n


def _run_lazy_ab(monkeypatch, no_lazy, log_cap=None, kind="ganmf"):
    """20 epochs of D+G steps with scoring, snapshot and restore in between; returns everything observable."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    monkeypatch.setenv("GANMF_NO_LAZY_ADAM", "1" if no_lazy else "0")
    if log_cap is not None:
        monkeypatch.setenv("GANMF_LAZY_LOG_CAP", str(log_cap))
    else:
        monkeypatch.delenv("GANMF_LAZY_LOG_CAP", raising=False)
    n_rows, width, k, E, B = 300, 517, 24, 40, 64
    urm = make_urm(n_rows, width, 0.05, 3)
    if kind == "ganmf":
        p0 = to.init_ganmf_params(n_rows, width, k, E, seed=4)
        eng = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B, gemm_path=L.GEMM_TC)
    else:
        p0 = to.init_disganmf_params(n_rows, width, k, 2, 32, seed=4)
        eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=2, d_nodes=32, d_act="tanh", max_batch=B)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    out = []
    users = np.arange(0, n_rows, 7, dtype=np.int32)
    for ep, batches in to.epoch_index_stream(n_rows, B, 20, seed=1337):
        perm = np.concatenate(batches)
        if ep % 3 == 2:
            perm = perm[:200]                       # some rows sit out whole epochs
        dl, gl = eng.train_epoch(perm, B, 1, 1, HP["d_lr"], HP["g_lr"], HP["d_reg"], 0.0, HP["m"], HP["alpha"])
        out += [np.asarray(dl), np.asarray(gl)]
        if ep == 4:
            out.append(eng.score(users))            # reads the user factors mid-training
        if ep == 7:
            eng.snapshot()
        if ep == 10:
            eng.reset_optimizers()                  # moments zeroed: steps still deferred must land first
        if ep == 12:
            eng.restore()                           # theta replaced, moments kept
        if ep == 15:                                # a step with g_reg != 0 takes the dense path
            dl, gl = eng.train_epoch(perm[:64], B, 1, 1, HP["d_lr"], HP["g_lr"], HP["d_reg"], 1e-3, HP["m"],
                                     HP["alpha"])
            out += [np.asarray(dl), np.asarray(gl)]
    got = eng.get_params()
    out += [got[n] for n in sorted(got)]
    eng.close()
    return out


@pytest.mark.parametrize("kind", ["ganmf", "disganmf"])
@pytest.mark.parametrize("log_cap", [None, 3])
def test_lazy_user_factor_adam_is_bit_identical_to_dense(monkeypatch, kind, log_cap):
    """Deferring the zero-gradient Adam steps of unsampled user-factor rows (kernels.cuh K6b) must not change
    a single bit of any loss, score or parameter relative to TF's dense sweep; log_cap=3 forces the
    step-size log to wrap (flush) every third G step."""
    dense = _run_lazy_ab(monkeypatch, True, None, kind)
    lazy = _run_lazy_ab(monkeypatch, False, log_cap, kind)
    assert len(dense) == len(lazy)
    for a, b in zip(dense, lazy):
        assert a.shape == b.shape
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
