"""NumPy restatement of the GANMF / DisGANMF training arithmetic (TEST INFRASTRUCTURE).

Follows the TF-1.12 graph built by the reference:
  GANRec/GANMF.py:53-139      (autoencoder discriminator, MF generator, losses, minimize)
  GANRec/GANMF.py:172-203     (epoch schedule: full D pass, then full G pass, same batches)
  GANRec/DisGANMF.py:51-140   (MLP discriminator on concat([float(id), profile]))
TF semantics restated (TensorFlow is not vendored in /root/reference and cannot be installed
here; TF 1.12.0 is pinned in conda_requirements.txt:15):
  tf.losses.mean_squared_error  = sum((a-b)^2) / num_elements, gradient through BOTH arguments
  tf.layers.dense               = x @ kernel + bias (bias zeros-initialised)
  tf.nn.l2_loss                 = sum(x^2) / 2
  tf.maximum(0.0, x)            = gradient 1 only where x > 0 (tie goes to the constant)
  tf.train.AdamOptimizer        = ApplyAdam: m += (g-m)(1-b1); v += (g^2-v)(1-b2);
                                  var -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps);
                                  the IndexedSlices gradient of user_embeddings is summed
                                  with the dense 0*l2 term, so the update is DENSE over all rows
  sigmoid_cross_entropy_with_logits(z, x) = max(x,0) - x*z + log1p(exp(-|x|))
Parity status: UNPINNED at step level (no reference test exists; see oracle/__init__.py);
cross-checked against torch.autograd (fp64) in tests/test_oracle_autograd.py.
"""
import numpy as np

BETA1, BETA2, EPS = 0.9, 0.999, 1e-8

GANMF_D = ["autoencoder/encoding/kernel", "autoencoder/encoding/bias",
           "autoencoder/decoding/kernel", "autoencoder/decoding/bias"]
GANMF_G = ["generator/user_embeddings", "generator/item_embeddings"]


def glorot_uniform(rng, shape, dtype=np.float32):
    """tf.glorot_uniform_initializer: U(-l, l), l = sqrt(6 / (fan_in + fan_out)) (GANMF.py:57)."""
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def epoch_index_stream(num_rows, batch_size, epochs, seed=1337):
    """Minibatch id stream of the reference fit loop (GANMF.py:156,175-184): ONE cumulative
    in-place np.random.shuffle per epoch; the same batches feed the D pass and the G pass.
    Yields (epoch, [uids_batch0, uids_batch1, ...])."""
    rs = np.random.RandomState(seed)      # legacy RandomState == np.random.seed(seed) stream
    all_users = np.arange(num_rows)
    for epoch in range(1, epochs + 1):
        rs.shuffle(all_users)
        yield epoch, [all_users[s:s + batch_size].copy() for s in range(0, num_rows, batch_size)]


class TFAdam:
    """tf.train.AdamOptimizer defaults (GANMF.py:104-105), dense ApplyAdam kernel form."""

    def __init__(self, params, names, lr, dtype):
        self.dtype = dtype
        self.lr = dtype(lr)
        self.m = {n: np.zeros_like(params[n]) for n in names}
        self.v = {n: np.zeros_like(params[n]) for n in names}
        self.b1p = dtype(BETA1)
        self.b2p = dtype(BETA2)

    def apply(self, params, grads):
        dt = self.dtype
        alpha = dt(self.lr * np.sqrt(dt(1) - self.b2p) / (dt(1) - self.b1p))
        for n, g in grads.items():
            m, v = self.m[n], self.v[n]
            m += (g - m) * dt(1 - BETA1)
            v += (g * g - v) * dt(1 - BETA2)
            params[n] -= (m * alpha) / (np.sqrt(v) + dt(EPS))
        self.b1p = dt(self.b1p * dt(BETA1))
        self.b2p = dt(self.b2p * dt(BETA2))


def init_ganmf_params(n_rows, width, num_factors, emb_dim, seed=1234, dtype=np.float32):
    rng = np.random.RandomState(seed)
    return {
        "autoencoder/encoding/kernel": glorot_uniform(rng, (width, emb_dim), dtype),
        "autoencoder/encoding/bias": np.zeros(emb_dim, dtype),
        "autoencoder/decoding/kernel": glorot_uniform(rng, (emb_dim, width), dtype),
        "autoencoder/decoding/bias": np.zeros(width, dtype),
        "generator/user_embeddings": glorot_uniform(rng, (n_rows, num_factors), dtype),
        "generator/item_embeddings": glorot_uniform(rng, (width, num_factors), dtype),
    }


class GanmfOracle:
    """One object = the TF session of GANMF.fit (GANMF.py:88-244) after variable init."""

    def __init__(self, params, d_lr, g_lr, dtype=np.float32):
        self.dtype = dtype
        self.p = {k: np.array(v, dtype=dtype) for k, v in params.items()}
        self.opt_d = TFAdam(self.p, GANMF_D, d_lr, dtype)
        self.opt_g = TFAdam(self.p, GANMF_G, g_lr, dtype)

    # -- forward pieces (GANMF.py:62-84)
    def _generator(self, uids):
        Pb = self.p["generator/user_embeddings"][uids]
        return Pb, Pb @ self.p["generator/item_embeddings"].T

    def _autoencoder(self, X):
        We, be = self.p[GANMF_D[0]], self.p[GANMF_D[1]]
        Wd, bd = self.p[GANMF_D[2]], self.p[GANMF_D[3]]
        H = X @ We + be
        D = H @ Wd + bd
        res = D - X
        loss = self.dtype((res * res).sum(dtype=np.float64) / res.size)
        return H, res, loss

    def d_step(self, uids, R, d_reg=0.0, m=1.0):
        """sess.run([dtrain, dloss]) (GANMF.py:131-132,138,186-187). Returns dloss before the update."""
        dt = self.dtype
        R = np.asarray(R, dtype=dt)
        _, F = self._generator(uids)                 # constant w.r.t. the D variables
        Hr, res_r, Lr = self._autoencoder(R)
        Hf, res_f, Lf = self._autoencoder(F)
        We, be, Wd, bd = (self.p[n] for n in GANMF_D)
        l2 = sum(float((self.p[n].astype(np.float64) ** 2).sum()) for n in GANMF_D) / 2.0
        hinge = dt(m) * Lr - Lf
        dloss = float(Lr) + max(0.0, float(hinge)) + d_reg * l2
        gate = 1.0 if hinge > 0 else 0.0             # strict: tie -> gradient to the constant 0.0
        N = dt(R.size)
        Gr = res_r * dt((1.0 + gate * m) * 2.0 / N)
        Gf = res_f * dt(-gate * 2.0 / N)
        dWd = Hr.T @ Gr + Hf.T @ Gf + dt(d_reg) * Wd
        dbd = Gr.sum(0) + Gf.sum(0) + dt(d_reg) * bd
        dHr = Gr @ Wd.T
        dHf = Gf @ Wd.T
        dWe = R.T @ dHr + F.T @ dHf + dt(d_reg) * We
        dbe = dHr.sum(0) + dHf.sum(0) + dt(d_reg) * be
        self.opt_d.apply(self.p, dict(zip(GANMF_D, (dWe, dbe, dWd, dbd))))
        return dloss

    def g_step(self, uids, R, g_reg=0.0, recon_coefficient=1e-2):
        """sess.run([gtrain, gloss]) (GANMF.py:133-135,139,200-201). Returns gloss before the update."""
        dt = self.dtype
        a = recon_coefficient
        R = np.asarray(R, dtype=dt)
        P, V = self.p[GANMF_G[0]], self.p[GANMF_G[1]]
        We, be, Wd, bd = (self.p[n] for n in GANMF_D)
        Pb, F = self._generator(uids)
        Hr = R @ We + be
        Hf, res_f, Lf = self._autoencoder(F)
        dH = Hr - Hf
        fm = dt((dH * dH).sum(dtype=np.float64) / dH.size)
        l2 = sum(float((self.p[n].astype(np.float64) ** 2).sum()) for n in GANMF_G) / 2.0
        gloss = (1.0 - a) * float(Lf) + a * float(fm) + g_reg * l2
        N, M = dt(F.size), dt(Hf.size)
        Gf = res_f * dt((1.0 - a) * 2.0 / N)
        dHf = Gf @ Wd.T + (Hf - Hr) * dt(a * 2.0 / M)
        dF = dHf @ We.T - Gf                           # F is also the MSE label: no stop_gradient in TF
        dV = dF.T @ Pb + dt(g_reg) * V
        dPb = dF @ V
        dP = dt(g_reg) * P                             # dense term; rows outside the batch get 0 (+reg)
        np.add.at(dP, uids, dPb)
        self.opt_g.apply(self.p, {GANMF_G[0]: dP, GANMF_G[1]: dV})
        return gloss

    def scores(self, ids):
        """_compute_item_score in the training orientation (GANMF.py:291-292)."""
        return self._generator(np.asarray(ids))[1]


# --------------------------------------------------------------------------- DisGANMF
def _act(name, z):
    if name in (None, "linear"):
        return z
    if name == "tanh":
        return np.tanh(z)
    if name == "relu":
        return np.maximum(z, 0)
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-z))
    raise ValueError(name)


def _act_grad_from_output(name, h):
    if name in (None, "linear"):
        return np.ones_like(h)
    if name == "tanh":
        return 1 - h * h
    if name == "relu":
        return (h > 0).astype(h.dtype)
    if name == "sigmoid":
        return h * (1 - h)
    raise ValueError(name)


def disganmf_d_names(d_layers):
    names = []
    for l in range(d_layers):
        names += ["discriminator/layer_%d/kernel" % l, "discriminator/layer_%d/bias" % l]
    return names + ["discriminator/D_output/kernel", "discriminator/D_output/bias"]


def init_disganmf_params(n_rows, width, num_factors, d_layers, d_nodes, seed=1234, dtype=np.float32):
    rng = np.random.RandomState(seed)
    p = {}
    fan_in = width + 1
    for l in range(d_layers):
        p["discriminator/layer_%d/kernel" % l] = glorot_uniform(rng, (fan_in, d_nodes), dtype)
        p["discriminator/layer_%d/bias" % l] = np.zeros(d_nodes, dtype)
        fan_in = d_nodes
    p["discriminator/D_output/kernel"] = glorot_uniform(rng, (fan_in, 1), dtype)
    p["discriminator/D_output/bias"] = np.zeros(1, dtype)
    p["generator/user_embeddings"] = glorot_uniform(rng, (n_rows, num_factors), dtype)
    p["generator/item_embeddings"] = glorot_uniform(rng, (width, num_factors), dtype)
    return p


def _softplus(x):
    # sigmoid_cross_entropy_with_logits(labels=0, logits=x) = max(x,0) + log1p(exp(-|x|))
    return np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class DisGanmfOracle:
    """TF session of DisGANMF.fit (DisGANMF.py:83-140)."""

    def __init__(self, params, d_layers, d_hidden_act, d_lr, g_lr, dtype=np.float32):
        self.dtype = dtype
        self.L = d_layers
        self.act = d_hidden_act
        self.p = {k: np.array(v, dtype=dtype) for k, v in params.items()}
        self.d_names = disganmf_d_names(d_layers)
        self.opt_d = TFAdam(self.p, self.d_names, d_lr, dtype)
        self.opt_g = TFAdam(self.p, GANMF_G, g_lr, dtype)

    def _generator(self, uids):
        Pb = self.p["generator/user_embeddings"][uids]
        return Pb, Pb @ self.p["generator/item_embeddings"].T

    def _disc_forward(self, uids, prof):
        """discriminator(cast(user_id, f32), profile) (DisGANMF.py:57-65,110-111)."""
        dt = self.dtype
        x = np.concatenate([np.asarray(uids, dtype=dt).reshape(-1, 1), prof], axis=1)
        hs = [x]
        for l in range(self.L):
            z = hs[-1] @ self.p["discriminator/layer_%d/kernel" % l] + self.p["discriminator/layer_%d/bias" % l]
            hs.append(_act(self.act, z).astype(dt))
        out = hs[-1] @ self.p["discriminator/D_output/kernel"] + self.p["discriminator/D_output/bias"]
        return hs, out

    def _disc_backward(self, hs, d_out, d_feat_extra=None):
        """Back-propagate d_out (B x 1) [+ extra gradient on the last hidden layer].
        Returns (param grads dict, gradient w.r.t. the concatenated input)."""
        g = {}
        wo = self.p["discriminator/D_output/kernel"]
        g["discriminator/D_output/kernel"] = hs[-1].T @ d_out
        g["discriminator/D_output/bias"] = d_out.sum(0)
        dh = d_out @ wo.T
        if d_feat_extra is not None:
            dh = dh + d_feat_extra
        for l in reversed(range(self.L)):
            dz = dh * _act_grad_from_output(self.act, hs[l + 1])
            g["discriminator/layer_%d/kernel" % l] = hs[l].T @ dz
            g["discriminator/layer_%d/bias" % l] = dz.sum(0)
            dh = dz @ self.p["discriminator/layer_%d/kernel" % l].T
        return g, dh

    def d_step(self, uids, R, d_reg=0.0):
        dt = self.dtype
        R = np.asarray(R, dtype=dt)
        B = dt(len(uids))
        _, F = self._generator(uids)
        hs_r, out_r = self._disc_forward(uids, R)
        hs_f, out_f = self._disc_forward(uids, F)
        loss_real = float(_softplus(-out_r).mean(dtype=np.float64))   # labels = 1
        loss_fake = float(_softplus(out_f).mean(dtype=np.float64))    # labels = 0
        l2 = sum(float((self.p[n].astype(np.float64) ** 2).sum()) for n in self.d_names) / 2.0
        dloss = loss_real + loss_fake + d_reg * l2
        g_r, _ = self._disc_backward(hs_r, (-_sigmoid(-out_r) / B).astype(dt))
        g_f, _ = self._disc_backward(hs_f, (_sigmoid(out_f) / B).astype(dt))
        grads = {n: g_r[n] + g_f[n] + dt(d_reg) * self.p[n] for n in self.d_names}
        self.opt_d.apply(self.p, grads)
        return dloss

    def g_step(self, uids, R, g_reg=0.0, recon_coefficient=1e-2):
        dt = self.dtype
        a = recon_coefficient
        R = np.asarray(R, dtype=dt)
        B = dt(len(uids))
        P, V = self.p[GANMF_G[0]], self.p[GANMF_G[1]]
        Pb, F = self._generator(uids)
        hs_r, _ = self._disc_forward(uids, R)
        hs_f, out_f = self._disc_forward(uids, F)
        feat_r, feat_f = hs_r[-1], hs_f[-1]
        loss_fake = float(_softplus(out_f).mean(dtype=np.float64))    # G minimises this as written
        dfe = feat_r - feat_f
        fm = float((dfe * dfe).sum(dtype=np.float64) / dfe.size)
        l2 = sum(float((self.p[n].astype(np.float64) ** 2).sum()) for n in GANMF_G) / 2.0
        gloss = loss_fake + a * fm + g_reg * l2
        extra = (feat_f - feat_r) * dt(a * 2.0 / feat_f.size)
        _, dx = self._disc_backward(hs_f, (_sigmoid(out_f) / B).astype(dt), extra)
        dF = dx[:, 1:]                                  # column 0 is the (constant) id feature
        dV = dF.T @ Pb + dt(g_reg) * V
        dPb = dF @ V
        dP = dt(g_reg) * P
        np.add.at(dP, uids, dPb)
        self.opt_g.apply(self.p, {GANMF_G[0]: dP, GANMF_G[1]: dV})
        return gloss

    def scores(self, ids):
        return self._generator(np.asarray(ids))[1]


def csr_rows_to_dense(urm, uids, dtype=np.float32):
    """URM_train[uids].toarray() (GANMF.py:184)."""
    return np.asarray(urm[uids].toarray(), dtype=dtype)
