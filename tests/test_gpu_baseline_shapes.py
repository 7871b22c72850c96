"""Step parity at the shapes BASELINE.json names: the committed splits with the best_params of
test_results/*, and five steps at the cfg4 synthetic shape.  Same contract as tests/test_gpu_train_parity.py: identical
initial weights and minibatch id stream, per-step losses and updated parameters within rel 1e-3 of the fp32 oracle.

DisGANMF runs its tensor-core GEMMs on split-TF32 operands (GANMF_GEMM_TC3, three MMAs per product): the BCE gradients
of the real and the fake half cancel in the weight-gradient sums, which makes plain TF32 rounding visible at ~1e-2 on
wide nets; the test of the wide configuration records that too."""
import numpy as np
import pytest

from oracle import train_oracle as to
from tests.helpers import load_quality_targets, load_split
from tests.test_gpu_train_parity import REL, rel_err

pytestmark = pytest.mark.gpu


def pick_batches(n_rows, B, n_first):
    _, batches = next(iter(to.epoch_index_stream(n_rows, B, 1, seed=1337)))
    out = batches[:n_first]
    if len(batches) > n_first:
        out = out + [batches[-1]]                  # the (usually short) last batch of the epoch
    return out


def run_engine(eng, batches, d_args, g_args):
    perm = np.concatenate(batches).astype(np.int32)
    eng.upload_ids(perm)
    off, slot = 0, 0
    for b in batches:
        eng.d_step(off, len(b), *d_args, loss_slot=slot)
        off += len(b)
        slot += 1
    off = 0
    for b in batches:
        eng.g_step(off, len(b), *g_args, loss_slot=slot)
        off += len(b)
        slot += 1
    return eng.read_losses(slot)


def check(losses, want, got, ref, rel=REL):
    np.testing.assert_allclose(losses, want, rtol=rel)
    worst = max((rel_err(got[n], ref[n]), n) for n in ref)
    assert worst[0] <= rel, worst


def test_cfg2_ganmf_item_lastfm_steps_parity():
    """GANMF --item on the committed LastFM split: rows are the 17 632 items, profiles are 1 884 users wide."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    bp = load_quality_targets()["GANMF_item_LastFM"]["best_params"]
    urm = load_split("LastFM")["train"].T.tocsr()                    # GANMF.py:32-33
    n_rows, width = urm.shape
    k, E, B = int(bp["num_factors"]), int(bp["emb_dim"]), int(bp["batch_size"])
    assert (n_rows, width, k, E, B) == (17632, 1884, 146, 680, 512)
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=7)
    eng = Engine(L.KIND_GANMF, n_rows, width, k, emb_dim=E, max_batch=B, item_mode=True)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    batches = pick_batches(n_rows, B, 6)
    losses = run_engine(eng, batches, (bp["d_lr"], bp["d_reg"], bp["m"]), (bp["g_lr"], 0.0, bp["recon_coefficient"]))
    orc = to.GanmfOracle(p0, bp["d_lr"], bp["g_lr"], dtype=np.float32)
    want = [orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=bp["d_reg"], m=bp["m"]) for b in batches]
    want += [orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=0.0, recon_coefficient=bp["recon_coefficient"])
             for b in batches]
    check(losses, want, eng.get_params(), orc.p)
    eng.close()


def disganmf_case(name, ds, item_mode, n_first, gemm_path):
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    bp = load_quality_targets()[name]["best_params"]
    urm = load_split(ds)["train"]
    urm = urm.T.tocsr() if item_mode else urm.tocsr()
    n_rows, width = urm.shape
    k, B, layers, nodes = int(bp["num_factors"]), int(bp["batch_size"]), int(bp["d_layers"]), int(bp["d_nodes"])
    act = bp["d_hidden_act"]
    p0 = to.init_disganmf_params(n_rows, width, k, layers, nodes, seed=3)
    eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=layers, d_nodes=nodes, d_act=act, max_batch=B,
                 item_mode=item_mode, gemm_path=getattr(L, gemm_path))
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    batches = pick_batches(n_rows, B, n_first)
    losses = run_engine(eng, batches, (bp["d_lr"], bp["d_reg"], 1.0), (bp["g_lr"], 0.0, bp["recon_coefficient"]))
    orc = to.DisGanmfOracle(p0, layers, act, bp["d_lr"], bp["g_lr"], dtype=np.float32)
    want = [orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=bp["d_reg"]) for b in batches]
    want += [orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=0.0, recon_coefficient=bp["recon_coefficient"])
             for b in batches]
    got = eng.get_params()
    eng.close()
    return (n_rows, width, k, B, layers, nodes), losses, np.array(want), got, orc.p


def test_cfg3_disganmf_user_hetrec2011_steps_parity():
    """DisGANMF --user, hetrec2011 split: 2113 rows, profiles 10 109 wide (+ the id column), one 4-unit layer."""
    shape, losses, want, got, ref = disganmf_case("DisGANMF_user_hetrec2011", "Movielenshetrec2011", False, 12, "GEMM_AUTO")
    assert shape == (2113, 10109, 243, 64, 1, 4)
    check(losses, want, got, ref)


def test_cfg3_disganmf_item_hetrec2011_steps_parity():
    """DisGANMF --item, hetrec2011 split: 10 109 rows, 4 hidden layers of 1024 units, B = 256 -- the wide case, and the
    only committed configuration with more than 8 discriminator tensors (this test found an overflow of the fused-Adam
    segment table that left layer_3/bias and the output layer without updates).  The raw row id (up to 10 108) is a
    discriminator input (DisGANMF.py:110): logits are O(100..1000), the D loss O(100).

    (1) free run on the shipping path (split-TF32 tensor-core GEMMs): losses and weights within 1e-3 of the fp32 oracle;
    (2) the same on plain TF32 for the record: the BCE gradients of the real and the fake half cancel in the
        weight-gradient sums, TF32 rounding drifts to ~1e-1 within ten steps -- which is why DisGANMF defaults to
        GANMF_GEMM_TC3;
    (3) per update (teacher forcing from the float64 oracle's weights, zero Adam moments): the device is as close to the
        float64 oracle as the float32 oracle is, on the exact-FMA path and on the tensor-core path."""
    shape, losses, want, got, ref = disganmf_case("DisGANMF_item_hetrec2011", "Movielenshetrec2011", True, 5, "GEMM_AUTO")
    assert shape == (10109, 2113, 25, 256, 4, 1024)
    check(losses, want, got, ref)
    _, l32, _, g32, _ = disganmf_case("DisGANMF_item_hetrec2011", "Movielenshetrec2011", True, 5, "GEMM_TC")
    print("plain TF32 on the 4x1024 net: loss rel err per step %s, worst tensor rel err %.1e" %
          (np.array2string(np.abs(l32 - want) / np.abs(want), precision=1), max(rel_err(g32[n], ref[n]) for n in ref)))
    bp = load_quality_targets()["DisGANMF_item_hetrec2011"]["best_params"]
    urm = load_split("Movielenshetrec2011")["train"].T.tocsr()
    k, B, layers, nodes = int(bp["num_factors"]), int(bp["batch_size"]), int(bp["d_layers"]), int(bp["d_nodes"])
    for path in ("GEMM_SIMT", "GEMM_AUTO"):
        _cfg3_item_teacher_forced(path, 1e-3, bp, urm, k, B, layers, nodes, bp["d_hidden_act"], to.disganmf_d_names(layers))


def _cfg3_item_teacher_forced(path, tol, bp, urm, k, B, layers, nodes, act, d_names):
    """One D update and one G update per batch, each from the float64 oracle's current weights with zero Adam moments:
    loss within 1e-3, update (theta_new - theta_old) as close to the float64 oracle's as the float32 oracle's is
    (factor 3, floor `tol`)."""
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    n_rows, width = urm.shape
    p = to.init_disganmf_params(n_rows, width, k, layers, nodes, seed=3)
    eng = Engine(L.KIND_DISGANMF, n_rows, width, k, d_layers=layers, d_nodes=nodes, d_act=act, max_batch=B, item_mode=True,
                 gemm_path=getattr(L, path))
    eng.set_csr(L.CSR_TRAIN, urm)

    def upd_err(got, ref64, names, p_old):
        num = sum(float(np.sum((got[n].astype(np.float64) - ref64[n]) ** 2)) for n in names)
        den = sum(float(np.sum((ref64[n] - p_old[n].astype(np.float64)) ** 2)) for n in names)
        return np.sqrt(num / max(den, 1e-300))

    report = []
    for b in pick_batches(n_rows, B, 4):
        R = to.csr_rows_to_dense(urm, b)
        for which in ("d", "g"):
            eng.set_params(p)
            eng.reset_optimizers()
            o32 = to.DisGanmfOracle(p, layers, act, bp["d_lr"], bp["g_lr"], dtype=np.float32)
            o64 = to.DisGanmfOracle(p, layers, act, bp["d_lr"], bp["g_lr"], dtype=np.float64)
            eng.upload_ids(b.astype(np.int32))
            if which == "d":
                eng.d_step(0, len(b), bp["d_lr"], bp["d_reg"], 1.0, loss_slot=0)
                l32, l64 = (o.d_step(b, R, d_reg=bp["d_reg"]) for o in (o32, o64))
                names = d_names
            else:
                eng.g_step(0, len(b), bp["g_lr"], 0.0, bp["recon_coefficient"], loss_slot=0)
                l32, l64 = (o.g_step(b, R, g_reg=0.0, recon_coefficient=bp["recon_coefficient"]) for o in (o32, o64))
                names = to.GANMF_G
            got_l = float(eng.read_losses(1)[0])
            assert abs(got_l - l64) <= REL * abs(l64), (which, got_l, l32, l64)
            got = eng.get_params()
            e_dev, e_o32 = upd_err(got, o64.p, names, p), upd_err(o32.p, o64.p, names, p)
            report.append((which, abs(got_l - l64) / abs(l64), e_dev, e_o32))
            # the update of the device is as close to the float64 update as the float32 oracle's is
            assert e_dev <= max(3.0 * e_o32, tol), (path, which, e_dev, e_o32)
        p = {n: v.astype(np.float32) for n, v in o64.p.items()}   # next batch starts from the float64 oracle's weights
    for r in report:
        print("cfg3-item %s, %s step: loss rel err %.1e, update error vs fp64: device %.2e, fp32 oracle %.2e" % ((path,) + r))
    eng.close()


def test_cfg4_synthetic_shape_five_steps_parity():
    """BASELINE.json configs[3]: I = 27 000, k = 250, E = 1024, B = 1024 (split-K over K = 27 000, CTA pairs, fused-Adam
    epilogues) -- five D and five G steps against the oracle (~1 s per step on the host)."""
    import bench
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    c = bench.workload("cfg4")
    n_rows, I, k, E, B = 8192, c["items"], c["k"], c["E"], c["B"]            # user table cut: the shapes of a step do not depend on it
    urm = bench.synthetic_urm(n_rows, I, c["density"], 11)
    p0 = to.init_ganmf_params(n_rows, I, k, E, seed=5)
    hp = bench.HP
    eng = Engine(L.KIND_GANMF, n_rows, I, k, emb_dim=E, max_batch=B)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    batches = pick_batches(n_rows, B, 5)[:5]
    losses = run_engine(eng, batches, (hp["d_lr"], hp["d_reg"], hp["m"]), (hp["g_lr"], hp["g_reg"], hp["alpha"]))
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    want = [orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"], m=hp["m"]) for b in batches]
    want += [orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"])
             for b in batches]
    check(losses, want, eng.get_params(), orc.p)
    eng.close()


def test_cfg5_item_width_two_steps_parity():
    """BASELINE.json configs[4] width: I = 200 000, k = 250, E = 1024 (K = 200 000 split-K accumulations, 782 pair tiles
    per row block); two D and two G steps at B = 128 against the oracle (the full B = 1024 step takes a minute per step
    on the host; the shapes that depend on I are the same)."""
    import bench
    from ganmf_b200 import _lib as L
    from ganmf_b200.engine import Engine
    c = bench.workload("cfg5")
    n_rows, I, k, E, B = 1024, c["items"], c["k"], c["E"], 128
    urm = bench.synthetic_urm(n_rows, I, c["density"], 13)
    p0 = to.init_ganmf_params(n_rows, I, k, E, seed=6)
    hp = bench.HP
    eng = Engine(L.KIND_GANMF, n_rows, I, k, emb_dim=E, max_batch=B)
    eng.set_csr(L.CSR_TRAIN, urm)
    eng.set_params(p0)
    eng.reset_optimizers()
    batches = pick_batches(n_rows, B, 2)[:2]
    losses = run_engine(eng, batches, (hp["d_lr"], hp["d_reg"], hp["m"]), (hp["g_lr"], hp["g_reg"], hp["alpha"]))
    orc = to.GanmfOracle(p0, hp["d_lr"], hp["g_lr"], dtype=np.float32)
    want = [orc.d_step(b, to.csr_rows_to_dense(urm, b), d_reg=hp["d_reg"], m=hp["m"]) for b in batches]
    want += [orc.g_step(b, to.csr_rows_to_dense(urm, b), g_reg=hp["g_reg"], recon_coefficient=hp["alpha"])
             for b in batches]
    check(losses, want, eng.get_params(), orc.p)
    eng.close()
