"""Cross-checks the closed-form gradients of oracle/train_oracle.py against torch.autograd (fp64)
on the TF graph of GANRec/GANMF.py:112-139 and GANRec/DisGANMF.py:103-140 (the per-step training
arithmetic is pinned by no reference test, see oracle/__init__.py)."""
import numpy as np
import pytest
import torch

from oracle import train_oracle as to

B1, B2, EPS = 0.9, 0.999, 1e-8


def tf_adam_torch(p, g, m, v, lr, t):
    alpha = lr * np.sqrt(1 - B2 ** t) / (1 - B1 ** t)
    m += (g - m) * (1 - B1)
    v += (g * g - v) * (1 - B2)
    p -= (m * alpha) / (v.sqrt() + EPS)


def mse(a, b):
    return ((a - b) ** 2).mean()


def _setup(n_rows=40, width=30, k=5, seed=0):
    rs = np.random.RandomState(seed)
    R = (rs.rand(n_rows, width) < 0.2).astype(np.float64)
    return rs, R


@pytest.mark.parametrize("m_hinge", [0.1, 10.0])      # both hinge branches (gate 0 / 1)
def test_ganmf_steps_match_autograd(m_hinge):
    rs, R_all = _setup()
    n_rows, width, k, E = 40, 30, 5, 7
    p0 = to.init_ganmf_params(n_rows, width, k, E, seed=3, dtype=np.float64)
    p0["autoencoder/encoding/bias"] = rs.randn(E) * 0.1
    p0["autoencoder/decoding/bias"] = rs.randn(width) * 0.1
    d_lr, g_lr, d_reg, g_reg, alpha = 1e-3, 2e-3, 1e-2, 1e-3, 0.3
    orc = to.GanmfOracle(p0, d_lr, g_lr, dtype=np.float64)
    tp = {n: torch.tensor(v, dtype=torch.float64, requires_grad=True) for n, v in p0.items()}
    st = {n: (torch.zeros_like(tp[n]), torch.zeros_like(tp[n])) for n in tp}

    def ae(x):
        h = x @ tp[to.GANMF_D[0]] + tp[to.GANMF_D[1]]
        d = h @ tp[to.GANMF_D[2]] + tp[to.GANMF_D[3]]
        return h, mse(x, d)

    def gen(uids):
        return tp[to.GANMF_G[0]][uids] @ tp[to.GANMF_G[1]].T

    td = tg = 0
    for step in range(6):
        uids = rs.permutation(n_rows)[:16 if step % 2 == 0 else 7]
        R = torch.tensor(R_all[uids])
        # D step
        fake = gen(torch.tensor(uids))
        _, lr_ = ae(R)
        _, lf_ = ae(fake)
        dloss = lr_ + torch.relu(m_hinge * lr_ - lf_) + d_reg * sum((tp[n] ** 2).sum() / 2 for n in to.GANMF_D)
        grads = torch.autograd.grad(dloss, [tp[n] for n in to.GANMF_D])
        td += 1
        with torch.no_grad():
            for n, g in zip(to.GANMF_D, grads):
                tf_adam_torch(tp[n], g, st[n][0], st[n][1], d_lr, td)
        got = orc.d_step(uids, R_all[uids], d_reg=d_reg, m=m_hinge)
        assert got == pytest.approx(float(dloss), rel=1e-10)
        # G step
        fake = gen(torch.tensor(uids))
        hr, _ = ae(R)
        hf, lf_ = ae(fake)
        gloss = (1 - alpha) * lf_ + alpha * mse(hr, hf) + g_reg * sum((tp[n] ** 2).sum() / 2 for n in to.GANMF_G)
        grads = torch.autograd.grad(gloss, [tp[n] for n in to.GANMF_G])
        tg += 1
        with torch.no_grad():
            for n, g in zip(to.GANMF_G, grads):
                tf_adam_torch(tp[n], g, st[n][0], st[n][1], g_lr, tg)
        got = orc.g_step(uids, R_all[uids], g_reg=g_reg, recon_coefficient=alpha)
        assert got == pytest.approx(float(gloss), rel=1e-10)
    for n in tp:
        np.testing.assert_allclose(orc.p[n], tp[n].detach().numpy(), rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("act,layers", [("linear", 1), ("tanh", 2), ("relu", 3), ("sigmoid", 2)])
def test_disganmf_steps_match_autograd(act, layers):
    rs, R_all = _setup(seed=1)
    n_rows, width, k, H = 40, 30, 5, 6
    p0 = to.init_disganmf_params(n_rows, width, k, layers, H, seed=5, dtype=np.float64)
    d_lr, g_lr, d_reg, g_reg, alpha = 1e-3, 2e-3, 1e-2, 1e-3, 0.3
    orc = to.DisGanmfOracle(p0, layers, act, d_lr, g_lr, dtype=np.float64)
    tp = {n: torch.tensor(v, dtype=torch.float64, requires_grad=True) for n, v in p0.items()}
    st = {n: (torch.zeros_like(tp[n]), torch.zeros_like(tp[n])) for n in tp}
    dn = to.disganmf_d_names(layers)
    actf = {"linear": lambda z: z, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid}[act]

    def disc(uids, prof):
        x = torch.cat([torch.tensor(uids, dtype=torch.float64).reshape(-1, 1), prof], dim=1)
        for l in range(layers):
            x = actf(x @ tp["discriminator/layer_%d/kernel" % l] + tp["discriminator/layer_%d/bias" % l])
        return x, x @ tp["discriminator/D_output/kernel"] + tp["discriminator/D_output/bias"]

    def gen(uids):
        return tp[to.GANMF_G[0]][uids] @ tp[to.GANMF_G[1]].T

    bce = torch.nn.functional.binary_cross_entropy_with_logits
    td = tg = 0
    for step in range(5):
        uids = rs.permutation(n_rows)[:12]
        R = torch.tensor(R_all[uids])
        fake = gen(torch.tensor(uids)).detach()
        _, o_r = disc(uids, R)
        _, o_f = disc(uids, fake)
        dloss = bce(o_r, torch.ones_like(o_r)) + bce(o_f, torch.zeros_like(o_f)) + \
            d_reg * sum((tp[n] ** 2).sum() / 2 for n in dn)
        grads = torch.autograd.grad(dloss, [tp[n] for n in dn])
        td += 1
        with torch.no_grad():
            for n, g in zip(dn, grads):
                tf_adam_torch(tp[n], g, st[n][0], st[n][1], d_lr, td)
        assert orc.d_step(uids, R_all[uids], d_reg=d_reg) == pytest.approx(float(dloss), rel=1e-10)
        fake = gen(torch.tensor(uids))
        f_r, _ = disc(uids, R)
        f_f, o_f = disc(uids, fake)
        gloss = bce(o_f, torch.zeros_like(o_f)) + alpha * mse(f_r, f_f) + \
            g_reg * sum((tp[n] ** 2).sum() / 2 for n in to.GANMF_G)
        grads = torch.autograd.grad(gloss, [tp[n] for n in to.GANMF_G])
        tg += 1
        with torch.no_grad():
            for n, g in zip(to.GANMF_G, grads):
                tf_adam_torch(tp[n], g, st[n][0], st[n][1], g_lr, tg)
        assert orc.g_step(uids, R_all[uids], g_reg=g_reg, recon_coefficient=alpha) == \
            pytest.approx(float(gloss), rel=1e-10)
    for n in tp:
        np.testing.assert_allclose(orc.p[n], tp[n].detach().numpy(), rtol=1e-8, atol=1e-11)


def test_index_stream_is_cumulative_shuffle():
    """GANMF.py:156,175: np.random.shuffle(all_users) in place, once per epoch, never reset."""
    np.random.seed(1337)
    ref = np.arange(23)
    want = []
    for _ in range(3):
        np.random.shuffle(ref)
        want.append(ref.copy())
    got = [np.concatenate(b) for _, b in to.epoch_index_stream(23, 5, 3, seed=1337)]
    for w, g in zip(want, got):
        assert w.tolist() == g.tolist()
    assert [len(b) for b in list(to.epoch_index_stream(23, 5, 1))[0][1]] == [5, 5, 5, 5, 3]
