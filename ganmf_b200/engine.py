"""Host-side handle of one device-resident GANMF / DisGANMF model (thin layer over the C ABI).

numpy / scipy objects in, numpy out; all arithmetic happens in libganmf_b200.so on the GPU."""
import ctypes as C

import numpy as np
import scipy.sparse as sps

from . import _lib as L


class DeviceArray(object):
    """Zero-copy view of a device buffer owned by the engine (for torch.as_tensor / NCCL)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class Engine(object):
    def __init__(self, kind, n_rows, width, num_factors, emb_dim=0, d_layers=0, d_nodes=0, d_act="linear",
                 max_batch=32, item_mode=False, row_id_offset=0, device=0, gemm_path=L.GEMM_AUTO,
                 global_width=0, item_offset=0, tp_rank=0, tp_world=0):
        """global_width / item_offset / tp_rank / tp_world: item-sharded training (parallel.ItemShardedTrainer);
        this engine then holds columns [item_offset, item_offset + width) of the training matrix."""
        self.lib = L.load()
        cfg = L.Config(kind=kind, n_rows=int(n_rows), width=int(width), num_factors=int(num_factors),
                       emb_dim=int(emb_dim), d_layers=int(d_layers), d_nodes=int(d_nodes), d_act=L.ACT[d_act],
                       max_batch=int(max_batch), item_mode=int(bool(item_mode)), row_id_offset=int(row_id_offset),
                       device=int(device), gemm_path=int(gemm_path), global_width=int(global_width),
                       item_offset=int(item_offset), tp_rank=int(tp_rank), tp_world=int(tp_world))
        self.cfg = cfg
        self.ctx = L._ctx()
        L.check(self.lib.ganmf_create(C.byref(cfg), C.byref(self.ctx)))
        self.n_rows, self.width = int(n_rows), int(width)
        self.n_users = self.width if item_mode else self.n_rows
        self.n_items = self.n_rows if item_mode else self.width
        self._test_key = None

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.ganmf_destroy(self.ctx)
            self.ctx = L._ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ plumbing
    def set_stream(self, cuda_stream_ptr):
        L.check(self.lib.ganmf_set_stream(self.ctx, C.c_void_p(int(cuda_stream_ptr))))

    def set_gemm_sms(self, n_sms):
        L.check(self.lib.ganmf_set_gemm_sms(self.ctx, int(n_sms)))

    def synchronize(self):
        L.check(self.lib.ganmf_synchronize(self.ctx))

    def step_routes(self):
        """{'sparse_real': codes of the real rows by CSR gather-sum, 'bias_grad_from_gemm': dbd from G3's column sums,
        'lowrank_fake': products over the generated profiles through V^T . We (ganmf_step_routes)}"""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(self.lib.ganmf_step_routes(self.ctx, C.byref(a), C.byref(b), C.byref(c)))
        return {"sparse_real": bool(a.value), "bias_grad_from_gemm": bool(b.value), "lowrank_fake": bool(c.value)}

    def launch_count(self):
        return int(self.lib.ganmf_launch_count(self.ctx))

    def profile(self, enable):
        L.check(self.lib.ganmf_profile(self.ctx, int(enable)))

    def profile_records(self, cap=4096):
        ms = np.zeros(cap, dtype=np.float64)
        shp = np.zeros((cap, 4), dtype=np.int32)
        n = C.c_int32()
        L.check(self.lib.ganmf_profile_records(self.ctx, ms.ctypes.data_as(L._f64p), shp.ctypes.data_as(L._i32p),
                                               cap, C.byref(n)))
        return ms[:n.value], shp[:n.value]

    def profile_read(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_int64()
        L.check(self.lib.ganmf_profile_read(self.ctx, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value

    def device_buffer(self, name):
        ptr, n = C.c_void_p(), C.c_int64()
        L.check(self.lib.ganmf_device_buffer(self.ctx, name.encode(), C.byref(ptr), C.byref(n)))
        return DeviceArray(ptr.value, n.value, "<f8" if name == "step_scalars" else "<f4")

    def device_buffer_ld(self, name):
        ld = C.c_int32()
        L.check(self.lib.ganmf_device_buffer_ld(self.ctx, name.encode(), C.byref(ld)))
        return ld.value

    # ------------------------------------------------------------------ data
    def set_csr(self, which, m, with_data=True):
        m = sps.csr_matrix(m)
        if not m.has_canonical_format:             # URM[uids].toarray() SUMS duplicate (row, col) entries
            m = m.copy()
            m.sum_duplicates()                     # (also sorts the indices of every row)
        _, ip = L.i32(m.indptr)
        idx, ix = L.i32(m.indices)
        if with_data:
            dat, dp = L.f32(m.data)
        else:
            dp = None
        L.check(self.lib.ganmf_set_csr(self.ctx, which, m.shape[0], m.shape[1], ip, ix, dp))
        return m

    def set_csr_transposed(self, which, m, with_data=True):
        """CSR `which` := m.T, transposed on the device (item mode: GANMF.py:32-33 does `URM_train.T.tocsr()` on the
        host).  Returns the canonical host matrix that was uploaded (NOT transposed)."""
        m = sps.csr_matrix(m)
        if not m.has_canonical_format:
            m = m.copy()
            m.sum_duplicates()
        _, ip = L.i32(m.indptr)
        idx, ix = L.i32(m.indices)
        if with_data:
            dat, dp = L.f32(m.data)
        else:
            dp = None
        L.check(self.lib.ganmf_set_csr_transposed(self.ctx, which, m.shape[0], m.shape[1], ip, ix, dp))
        return m

    def get_csr(self, which, with_data=True):
        """The resident CSR `which` as a scipy matrix (values = ones when it was set without data)."""
        nr, nc, nnz = C.c_int32(), C.c_int32(), C.c_int64()
        L.check(self.lib.ganmf_get_csr(self.ctx, which, C.byref(nr), C.byref(nc), C.byref(nnz), None, None, None))
        ip = np.empty(nr.value + 1, dtype=np.int32)
        ix = np.empty(nnz.value, dtype=np.int32)
        dat = np.ones(nnz.value, dtype=np.float32)
        L.check(self.lib.ganmf_get_csr(self.ctx, which, None, None, None, ip.ctypes.data_as(L._i32p),
                                       ix.ctypes.data_as(L._i32p), dat.ctypes.data_as(L._f32p) if with_data else None))
        return sps.csr_matrix((dat, ix, ip), shape=(nr.value, nc.value))

    def set_csr_device(self, which, n_rows, n_cols, indptr, indices, data=None):
        """CSR already on the device (torch int32 / float32 tensors or anything with data_ptr()); indices of a row
        sorted and unique.  Copied device-to-device; the caller may free its arrays afterwards."""
        nnz = int(indices.numel())
        L.check(self.lib.ganmf_set_csr_device(self.ctx, which, int(n_rows), int(n_cols), C.c_void_p(indptr.data_ptr()),
                                              C.c_void_p(indices.data_ptr() if nnz else 0),
                                              C.c_void_p(data.data_ptr()) if data is not None else None, nnz))

    # ------------------------------------------------------------------ parameters
    def param_infos(self):
        out = []
        for i in range(self.lib.ganmf_param_count(self.ctx)):
            name = C.create_string_buffer(128)
            r, c, g = C.c_int32(), C.c_int32(), C.c_int32()
            L.check(self.lib.ganmf_param_info(self.ctx, i, name, 128, C.byref(r), C.byref(c), C.byref(g)))
            out.append((name.value.decode(), r.value, c.value, bool(g.value)))
        return out

    def _shape(self, name):
        for n, r, c, _ in self.param_infos():
            if n == name:
                return r, c
        raise KeyError(name)

    def set_param(self, name, value):
        r, c = self._shape(name)
        a, p = L.f32(np.asarray(value, dtype=np.float32).reshape(r, c))
        L.check(self.lib.ganmf_set_param(self.ctx, name.encode(), p, a.size))

    def get_param(self, name):
        r, c = self._shape(name)
        out = np.empty((r, c), dtype=np.float32)
        L.check(self.lib.ganmf_get_param(self.ctx, name.encode(), out.ctypes.data_as(L._f32p), out.size))
        return out[0] if name.endswith("bias") else out

    def get_params(self):
        return {n: self.get_param(n) for n, _, _, _ in self.param_infos()}

    def set_params(self, params):
        for n, v in params.items():
            self.set_param(n, v)

    def init_params(self, seed):
        L.check(self.lib.ganmf_init_params(self.ctx, C.c_uint64(int(seed) & (2 ** 64 - 1))))

    def reset_optimizers(self):
        L.check(self.lib.ganmf_reset_optimizers(self.ctx))

    def snapshot(self):
        L.check(self.lib.ganmf_snapshot(self.ctx))

    def restore(self):
        L.check(self.lib.ganmf_restore(self.ctx))

    # ------------------------------------------------------------------ training
    def upload_ids(self, ids):
        a, p = L.i32(ids)
        L.check(self.lib.ganmf_upload_ids(self.ctx, p, a.size))

    def d_step(self, ids_offset, B, lr, reg, m_hinge=1.0, loss_slot=0, n_rows_global=None):
        L.check(self.lib.ganmf_d_step(self.ctx, ids_offset, B, n_rows_global or B, lr, reg, m_hinge, loss_slot))

    def g_step(self, ids_offset, B, lr, reg, recon_coefficient, loss_slot=0, n_rows_global=None):
        L.check(self.lib.ganmf_g_step(self.ctx, ids_offset, B, n_rows_global or B, lr, reg, recon_coefficient,
                                      loss_slot))

    def d_forward(self, ids_offset, B):
        L.check(self.lib.ganmf_d_forward(self.ctx, ids_offset, B))

    def d_backward(self, B, n_rows_global, m_hinge):
        L.check(self.lib.ganmf_d_backward(self.ctx, B, n_rows_global, m_hinge))

    def d_backward_phase(self, B, n_rows_global, m_hinge, phase):
        L.check(self.lib.ganmf_d_backward_phase(self.ctx, B, n_rows_global, m_hinge, phase))

    def d_apply(self, lr, reg, loss_slot):
        L.check(self.lib.ganmf_d_apply(self.ctx, lr, reg, loss_slot))

    def g_forward_backward_part(self, ids_offset, B, n_rows_global, recon_coefficient, part):
        L.check(self.lib.ganmf_g_forward_backward_part(self.ctx, ids_offset, B, n_rows_global, recon_coefficient,
                                                       part))

    def g_forward_backward(self, ids_offset, B, n_rows_global, recon_coefficient):
        L.check(self.lib.ganmf_g_forward_backward(self.ctx, ids_offset, B, n_rows_global, recon_coefficient))

    def g_apply(self, B, n_rows_global, lr, reg, recon_coefficient, loss_slot):
        L.check(self.lib.ganmf_g_apply(self.ctx, B, n_rows_global, lr, reg, recon_coefficient, loss_slot))

    def d_apply_ranges(self, lr, reg, offsets, counts, new_step=True):
        o = np.ascontiguousarray(offsets, dtype=np.int64)
        n = np.ascontiguousarray(counts, dtype=np.int64)
        L.check(self.lib.ganmf_d_apply_ranges(self.ctx, lr, reg, o.ctypes.data_as(L._i64p), n.ctypes.data_as(L._i64p),
                                              o.size, int(new_step)))

    def d_forward_phase(self, ids_offset, B, phase):
        L.check(self.lib.ganmf_d_forward_phase(self.ctx, ids_offset, B, phase))

    def tp_d_phase(self, phase, ids_offset, B, lr, reg, m_hinge, loss_slot):
        L.check(self.lib.ganmf_tp_d_phase(self.ctx, phase, ids_offset, B, lr, reg, m_hinge, loss_slot))

    def tp_g_phase(self, phase, ids_offset, B, lr, reg, recon_coefficient, loss_slot):
        L.check(self.lib.ganmf_tp_g_phase(self.ctx, phase, ids_offset, B, lr, reg, recon_coefficient, loss_slot))

    def finalize_loss(self, reg, loss_slot):
        L.check(self.lib.ganmf_finalize_loss(self.ctx, reg, loss_slot))

    def read_losses(self, n):
        out = np.empty(n, dtype=np.float32)
        L.check(self.lib.ganmf_read_losses(self.ctx, out.ctypes.data_as(L._f32p), n))
        return out

    def train_epoch(self, perm, batch_size, d_steps, g_steps, d_lr, g_lr, d_reg, g_reg, m_hinge, recon_coefficient):
        a, p = L.i32(perm)
        nb = (a.size + batch_size - 1) // batch_size
        dl = np.empty(nb * d_steps, dtype=np.float32)
        gl = np.empty(nb * g_steps, dtype=np.float32)
        L.check(self.lib.ganmf_train_epoch(self.ctx, p, a.size, batch_size, d_steps, g_steps, d_lr, g_lr, d_reg, g_reg,
                                           m_hinge, recon_coefficient, dl.ctypes.data_as(L._f32p),
                                           gl.ctypes.data_as(L._f32p)))
        return dl, gl

    # ------------------------------------------------------------------ scoring / evaluation
    def score(self, users):
        a, p = L.i32(users)
        out = np.empty((a.size, self.n_items), dtype=np.float32)
        L.check(self.lib.ganmf_score(self.ctx, p, a.size, out.ctypes.data_as(L._f32p)))
        return out

    def encode(self, rows):
        a, p = L.i32(rows)
        out = np.empty((a.size, self.cfg.emb_dim), dtype=np.float32)
        L.check(self.lib.ganmf_encode(self.ctx, p, a.size, out.ctypes.data_as(L._f32p)))
        return out

    def mask_topk(self, scores, K, users=None, remove_seen=False, write_back=False):
        s = np.ascontiguousarray(scores, dtype=np.float32)
        if write_back and s is not scores:
            raise ValueError("write_back needs a C-contiguous float32 array")
        n, n_items = s.shape
        idx = np.empty((n, K), dtype=np.int32)
        val = np.empty((n, K), dtype=np.float32)
        up = L.i32(users)[1] if users is not None else None
        ua = L.i32(users)[0] if users is not None else None     # keep alive
        L.check(self.lib.ganmf_mask_topk(self.ctx, s.ctypes.data_as(L._f32p), n, n_items,
                                         ua.ctypes.data_as(L._i32p) if ua is not None else None,
                                         int(remove_seen), K, idx.ctypes.data_as(L._i32p),
                                         val.ctypes.data_as(L._f32p), int(write_back)))
        return idx, val

    def recommend(self, users, K, remove_seen=True, return_scores=False):
        a, p = L.i32(users)
        idx = np.empty((a.size, K), dtype=np.int32)
        val = np.empty((a.size, K), dtype=np.float32)
        sc = np.empty((a.size, self.n_items), dtype=np.float32) if return_scores else None
        L.check(self.lib.ganmf_recommend(self.ctx, p, a.size, int(remove_seen), K, idx.ctypes.data_as(L._i32p),
                                         val.ctypes.data_as(L._f32p),
                                         sc.ctypes.data_as(L._f32p) if sc is not None else None))
        return idx, val, sc

    def set_test(self, urm_test, urm_train_for_popularity=None, item_popularity=None):
        """Uploads the held-out matrix and the numpy-made lookup tables of the metric kernels
        (gains 2^r-1, ln(j+2), per-item novelty / normalised popularity, metrics.py:298-392,693-722).
        The popularity comes from the users x items TRAINING matrix, or (row-sharded evaluation, where no rank
        holds all rows) from a ready per-item interaction count."""
        if (urm_train_for_popularity is None) == (item_popularity is None):
            raise ValueError("pass exactly one of urm_train_for_popularity / item_popularity")
        if urm_test.shape[1] != self.n_items:
            raise ValueError("URM_test has %d columns, the model ranks %d items" % (urm_test.shape[1], self.n_items))
        if item_popularity is not None:
            item_popularity = np.asarray(item_popularity)
            if item_popularity.shape != (urm_test.shape[1],):
                raise ValueError("item_popularity must have one entry per column of URM_test")
            key = (id(urm_test), urm_test.nnz, "pop", int(item_popularity.sum()))
        else:
            if urm_train_for_popularity.shape != urm_test.shape:
                # e.g. an item-mode model whose URM_train was left transposed (GANMF.py:32-33): the per-item
                # tables would be indexed out of bounds
                raise ValueError("URM_train is %r but URM_test is %r: both must be users x items" %
                                 (urm_train_for_popularity.shape, urm_test.shape))
            key = (id(urm_test), urm_test.nnz, urm_train_for_popularity.shape, urm_train_for_popularity.nnz)
        if self._test_key == key:
            return
        te = self.set_csr(L.CSR_TEST, urm_test)
        gain = (np.power(2, te.data.astype(np.float32)) - 1).astype(np.float32)
        # ideal DCG sorts the RELEVANCES descending (metrics.py:707); 2^r-1 is monotone in r
        rows = np.repeat(np.arange(te.shape[0]), np.diff(te.indptr))
        rel_sorted_gain = gain[np.lexsort((-gain, rows))] if gain.size else gain
        logtab = np.log(np.arange(L.TOPK_MAX, dtype=np.float32) + 2)
        if item_popularity is not None:
            pop = item_popularity.astype(np.int64)
        else:
            tr = sps.csc_matrix(urm_train_for_popularity)
            tr.eliminate_zeros()
            pop = np.ediff1d(tr.indptr)
        n_inter = pop.sum()
        n_items = len(pop)
        with np.errstate(divide="ignore"):
            nov = np.where(pop != 0, -np.log2(pop / n_inter) / n_items, 0.0)
        popn = pop / pop.max()
        haspop = (pop != 0).astype(np.uint8)
        g, gp = L.f32(gain)
        gd, gdp = L.f32(rel_sorted_gain)
        lt, ltp = L.f32(logtab)
        nv, nvp = L.f64(nov)
        pn, pnp = L.f64(popn)
        hp = np.ascontiguousarray(haspop)
        L.check(self.lib.ganmf_set_eval_tables(self.ctx, gp, gdp, ltp, lt.size, nvp, hp.ctypes.data_as(L._u8p), pnp,
                                               int(n_items)))
        self._test_key = key
        self._test_keepalive = urm_test

    def evaluate(self, users, cutoffs, remove_seen=True, block_size=0, want_counts=True):
        ua, up = L.i32(users)
        ca, cp = L.i32(cutoffs)
        sums = np.zeros((ca.size, L.MC_NCOL), dtype=np.float64)
        counts = np.zeros((ca.size, self.n_items), dtype=np.int64) if want_counts else None
        L.check(self.lib.ganmf_evaluate(self.ctx, up, ua.size, cp, ca.size, int(remove_seen), int(block_size),
                                        sums.ctypes.data_as(L._f64p),
                                        counts.ctypes.data_as(L._i64p) if counts is not None else None))
        return sums, counts

    def evaluate_values(self, users, cutoffs, remove_seen=True, block_size=0):
        """First half of evaluate(): per-user metric values of `users` (kept on the device)."""
        ua, up = L.i32(users)
        ca, cp = L.i32(cutoffs)
        self._ev_ncut = ca.size
        L.check(self.lib.ganmf_evaluate_values(self.ctx, up, ua.size, cp, ca.size, int(remove_seen), int(block_size)))

    def evaluate_sums(self, carry_in=None, want_counts=True):
        """Second half: running sums over those users IN ORDER, continuing from carry_in (the sums of the users
        that precede them on other GPUs)."""
        sums = np.zeros((self._ev_ncut, L.MC_NCOL), dtype=np.float64)
        counts = np.zeros((self._ev_ncut, self.n_items), dtype=np.int64) if want_counts else None
        cin = None
        if carry_in is not None:
            cin = np.ascontiguousarray(carry_in, dtype=np.float64)
            assert cin.shape == sums.shape
        L.check(self.lib.ganmf_evaluate_sums(self.ctx, cin.ctypes.data_as(L._f64p) if cin is not None else None,
                                             sums.ctypes.data_as(L._f64p),
                                             counts.ctypes.data_as(L._i64p) if counts is not None else None))
        return sums, counts

    def eval_stats(self):
        """(rows ranked by the fused scorer, rows that needed the exact fallback) since the engine was created."""
        a, b = C.c_int64(), C.c_int64()
        L.check(self.lib.ganmf_eval_stats(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def evaluate_scores(self, score_fn, users, cutoffs, remove_seen=True, block_size=1000):
        """Streaming evaluation of an arbitrary scorer: score_fn(user_ids) -> float32 [n, n_items] on the host;
        mask, top-k and metric arithmetic run on the device."""
        ua = np.ascontiguousarray(users, dtype=np.int32)
        ca, cp = L.i32(cutoffs)
        L.check(self.lib.ganmf_eval_begin(self.ctx, ua.size, cp, ca.size))
        for s in range(0, ua.size, block_size):
            blk = ua[s:s + block_size]
            sc = np.ascontiguousarray(score_fn(blk), dtype=np.float32)
            if sc.shape != (blk.size, self.n_items):
                raise ValueError("scores have shape %r, expected %r" % (sc.shape, (blk.size, self.n_items)))
            L.check(self.lib.ganmf_eval_scores_block(self.ctx, sc.ctypes.data_as(L._f32p),
                                                     blk.ctypes.data_as(L._i32p), blk.size, int(remove_seen), 0))
        sums = np.zeros((ca.size, L.MC_NCOL), dtype=np.float64)
        counts = np.zeros((ca.size, self.n_items), dtype=np.int64)
        L.check(self.lib.ganmf_eval_end(self.ctx, sums.ctypes.data_as(L._f64p), counts.ctypes.data_as(L._i64p)))
        return sums, counts

    def metrics_from_topk(self, topk_idx, users, cutoffs, want_per_user=False):
        t = np.ascontiguousarray(topk_idx, dtype=np.int32)
        ua, up = L.i32(users)
        ca, cp = L.i32(cutoffs)
        n, K = t.shape
        n_items = self.n_items
        sums = np.zeros((ca.size, L.MC_NCOL), dtype=np.float64)
        counts = np.zeros((ca.size, n_items), dtype=np.int64)
        per = np.zeros((n, ca.size, L.MC_NCOL), dtype=np.float64) if want_per_user else None
        L.check(self.lib.ganmf_metrics_from_topk(self.ctx, t.ctypes.data_as(L._i32p), K, up, n, cp, ca.size,
                                                 per.ctypes.data_as(L._f64p) if per is not None else None,
                                                 sums.ctypes.data_as(L._f64p), counts.ctypes.data_as(L._i64p)))
        return sums, counts, per
