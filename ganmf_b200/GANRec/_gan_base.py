"""Shared host logic of GANMF and DisGANMF: the reference's fit loop, snapshot hooks, scoring,
save / load -- everything of GANRec/GANMF.py and GANRec/DisGANMF.py that is not the TF graph."""
import os
import pickle
from datetime import datetime

import numpy as np

from .. import _lib as L
from ..Base.BaseRecommender import BaseRecommender
from ..Base.Incremental_Training_Early_Stopping import Incremental_Training_Early_Stopping
from ..Utils_ import EarlyStoppingScheduler
from ..engine import Engine


def _dist_state():
    """(torch.distributed, world size, rank) when a process group is up (torchrun), else (None, 1, 0)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist, dist.get_world_size(), dist.get_rank()
    except ImportError:
        pass
    return None, 1, 0


class GanRecommenderBase(BaseRecommender, Incremental_Training_Early_Stopping):
    RECOMMENDER_NAME = "GAN_Base"
    KIND = None

    def _init_common(self, URM_train, mode, seed, verbose, is_experiment):
        if mode not in ['user', 'item']:                               # GANMF.py:28-29
            raise ValueError('Accepted training modes are `user` and `item`. Given was {}.', mode)
        self.mode = mode
        URM_train = URM_train.tocsr()
        self._URM_users_items = URM_train                              # users x items, never flipped
        self.URM_train = URM_train.T.tocsr() if mode == 'item' else URM_train   # GANMF.py:31-35
        self.num_users, self.num_items = self.URM_train.shape          # rows / width of the training matrix
        self.config = None
        self.seed = seed
        self.verbose = verbose
        self.logsdir = os.path.join('plots', self.RECOMMENDER_NAME, datetime.now().strftime("%Y%m%d-%H%M%S"))
        self.is_experiment = is_experiment
        if not self.is_experiment:                                     # GANMF.py:44-51 (kept harmless)
            os.makedirs(os.path.join(self.logsdir, 'code'), exist_ok=True)
        self.items_to_ignore_flag = False
        self.items_to_ignore_ID = np.array([], dtype=int)
        self.filterTopPop_ItemsID = np.array([], dtype=int)
        self._engine = None
        self._trainer = None            # ItemShardedTrainer when fit() runs under torchrun (GANMF, world size > 1)
        self._scorer = None             # factor-only context holding the gathered factors of an item-sharded model
        self._scorer_dirty = True
        self._stop_training = False
        self.train_d_loss, self.train_g_loss = [], []
        self.params = self.best_params = None

    # ------------------------------------------------------------------ engine
    def _engine_kwargs(self):
        raise NotImplementedError()

    def _build_engine(self, batch_size, device=None, gemm_path=L.GEMM_AUTO):
        """One process (the reference's situation, GANMF.py:142-150): one context on one GPU.  Under torchrun
        (torch.distributed initialised, world size N > 1) a GANMF model is ITEM-SHARDED over the N GPUs
        (parallel.ItemShardedTrainer): every rank holds 1/N of the columns of the training matrix and of
        We / Wd / bd / V, steps on the same minibatch stream as a single GPU would, and all-reduces activations.
        DisGANMF has no sharded step: every rank then trains an identical replica."""
        if self._engine is not None:
            self._engine.close()
        if self._scorer is not None:
            self._scorer.close()
        self._trainer, self._scorer, self._scorer_dirty = None, None, True
        dist, world, rank = _dist_state()
        if os.environ.get("GANMF_GEMM_PATH"):             # A/B switch: 1 = exact fp32 FMA, 2 = TF32, 3 = split-TF32
            gemm_path = int(os.environ["GANMF_GEMM_PATH"])
        if device is None:
            device = 0
            if dist is not None:
                import torch
                device = torch.cuda.current_device()
        # item mode trains on the transposed matrix (GANMF.py:32-33); it is transposed on the device (set_csr_transposed)
        urm = self._URM_users_items
        n_rows, width = urm.shape[::-1] if self.mode == 'item' else urm.shape
        if world > 1 and self.KIND == L.KIND_GANMF:
            from ..parallel import ItemShardedTrainer, item_slices
            lo, hi = item_slices(width, world)[rank]
            eng = Engine(self.KIND, n_rows, hi - lo, max_batch=int(batch_size), item_mode=(self.mode == 'item'),
                         device=device, gemm_path=gemm_path, global_width=width, item_offset=lo, tp_rank=rank,
                         tp_world=world, **self._engine_kwargs())
            if self.mode == 'item':                 # columns [lo, hi) of URM^T = (rows [lo, hi) of URM)^T
                eng.set_csr_transposed(L.CSR_TRAIN, urm[lo:hi])
            else:
                eng.set_csr(L.CSR_TRAIN, urm[:, lo:hi].tocsr())
            eng.init_params(self.seed)           # every rank draws ITS slice of the same whole tensors
            self._engine = eng
            self._trainer = ItemShardedTrainer(eng)
            self._tp_slice = (lo, hi, width)
        else:
            eng = Engine(self.KIND, n_rows, width, max_batch=int(batch_size), item_mode=(self.mode == 'item'),
                         device=device, gemm_path=gemm_path, **self._engine_kwargs())
            if self.mode == 'item':
                eng.set_csr_transposed(L.CSR_TRAIN, urm)
            else:
                eng.set_csr(L.CSR_TRAIN, urm)
            eng.set_csr(L.CSR_SEEN, urm, with_data=False)
            eng.init_params(self.seed)
            self._engine = eng
        names = [(n, g) for n, _, _, g in eng.param_infos()]
        # same grouping as the reference's self.params / self.best_params (GANMF.py:119-128)
        self.params = {'D': [n for n, g in names if not g], 'G': [n for n, g in names if g]}
        self.best_params = {k: list(v) for k, v in self.params.items()}
        return eng

    # ------------------------------------------------------------------ fit loop (GANMF.py:142-244)
    def _fit_loop(self, epochs, batch_size, d_lr, g_lr, d_steps, g_steps, d_reg, g_reg, m, recon_coefficient,
                  allow_worse, freq, after, metrics, sample_every, validation_evaluator, validation_set,
                  earlystopping_kwargs):
        eng = self._build_engine(batch_size)
        self._has_best = False                     # a fresh engine holds no snapshot of an earlier fit()
        self._hp = dict(batch_size=int(batch_size), d_steps=int(d_steps), g_steps=int(g_steps), d_lr=float(d_lr),
                        g_lr=float(g_lr), d_reg=float(d_reg), g_reg=float(g_reg), m_hinge=float(m),
                        recon_coefficient=float(recon_coefficient))
        self._stop_training = False
        self._all_users = np.arange(self.num_users)
        self.train_d_loss, self.train_g_loss = [], []
        if earlystopping_kwargs:
            # PoliMi flavour (Base/Incremental_Training_Early_Stopping.py), as IALSRecommender.fit does
            self._train_with_early_stopping(epochs, algorithm_name=self.RECOMMENDER_NAME, **earlystopping_kwargs)
            if eng is self._engine and eng.lib is not None and self._has_best:
                eng.restore()
            self._finish_fit()
            return self.epochs_best
        early_stop = None
        if validation_evaluator is not None:
            # the reference's shadow variables exist (randomly initialised) from graph construction
            # (GANMF.py:123-128), so EarlyStoppingScheduler may call load_model() before any save_current_model():
            # the freshly initialised weights are the first snapshot
            eng.snapshot()
            early_stop = EarlyStoppingScheduler(self, evaluator=validation_evaluator, allow_worse=allow_worse,
                                                freq=freq, metrics=metrics, after=after)
        epoch = 1
        while not self._stop_training and epoch < epochs + 1:
            self._run_epoch(epoch)
            if validation_set is not None and sample_every is not None and epoch % sample_every == 0:
                self._flip_for_eval()
                _, results_run_string = validation_evaluator.evaluateRecommender(self)
                self._flip_for_eval()
                print(results_run_string)
            if early_stop is not None:
                self._flip_for_eval()
                early_stop(epoch)
                self._flip_for_eval()
                if self._stop_training:
                    print('Training stopped, epoch:', epoch)
            epoch += 1
        self._finish_fit()
        return epoch - 1 if self._stop_training else epoch            # GANMF.py:244

    def _flip_for_eval(self):
        if self.mode == 'item':                                        # GANMF.py:215-219,224-228
            self.URM_train = self.URM_train.T.tocsr()

    def _finish_fit(self):
        if self.mode == 'item':                                        # GANMF.py:241-242
            self.URM_train = self._URM_users_items

    # -- hooks of Incremental_Training_Early_Stopping
    _has_best = False

    def _run_epoch(self, num_epoch):
        np.random.shuffle(self._all_users)                             # GANMF.py:175 (global numpy RNG, cumulative)
        h = self._hp
        if self._trainer is not None:
            # every rank must step on the SAME minibatches: rank 0's shuffle is the one that counts
            import torch
            dist = self._trainer.dist
            t = torch.from_numpy(np.ascontiguousarray(self._all_users, dtype=np.int64))
            dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
            t = t.to(dev)
            dist.broadcast(t, src=0)
            self._all_users[:] = t.cpu().numpy()
            dl, gl = self._trainer.train_epoch(self._all_users.astype(np.int32), h['batch_size'], h['d_steps'],
                                               h['g_steps'], dict(d_lr=h['d_lr'], g_lr=h['g_lr'], d_reg=h['d_reg'],
                                                                  g_reg=h['g_reg'], m=h['m_hinge'],
                                                                  alpha=h['recon_coefficient']))
            self._scorer_dirty = True
        else:
            dl, gl = self._engine.train_epoch(self._all_users, h['batch_size'], h['d_steps'], h['g_steps'], h['d_lr'],
                                              h['g_lr'], h['d_reg'], h['g_reg'], h['m_hinge'], h['recon_coefficient'])
        self.last_d_losses, self.last_g_losses = dl, gl
        self.train_d_loss.append(float(np.mean(dl)) if dl.size else float('nan'))    # GANMF.py:205-209
        self.train_g_loss.append(float(np.mean(gl)) if gl.size else float('nan'))

    def _prepare_model_for_validation(self):
        if self.mode == 'item':
            self.URM_train = self._URM_users_items

    def _update_best_model(self):
        self.save_current_model()
        self._has_best = True

    # ------------------------------------------------------------------ reference hooks
    def stop_fit(self):                                                # GANMF.py:246-247
        self._stop_training = True

    def save_current_model(self):                                      # GANMF.py:249-251
        self._engine.snapshot()

    def load_model(self):                                              # GANMF.py:253-255 (restore best snapshot)
        self._engine.restore()
        self._scorer_dirty = True

    def get_URM_train(self):
        return self.URM_train.copy()

    # ------------------------------------------------------------------ scoring side
    def _device_engine(self):
        """The context that scores and ranks.  One GPU: the training context itself.  Item-sharded model: a
        factor-only context (GANMF_KIND_MF) on every rank holding ALL item factors (the slices are gathered over
        NVLink, device to device) and the replicated user factors; refreshed lazily after training steps."""
        if self._trainer is None:
            return self._engine
        import torch
        dist = self._trainer.dist
        eng = self._engine
        if self._scorer is None:
            lo, hi, width = self._tp_slice
            self._scorer = Engine(L.KIND_MF, eng.n_rows, width, eng.cfg.num_factors, max_batch=1,
                                  item_mode=(self.mode == 'item'), device=eng.cfg.device)
            self._scorer.set_csr(L.CSR_SEEN, self._URM_users_items, with_data=False)
            self._scorer_dirty = True
        if self._scorer_dirty:
            from ..parallel import item_slices
            lo, hi, width = self._tp_slice
            dev = torch.device("cuda", eng.cfg.device)
            ld = eng.device_buffer_ld("user_factors")
            wrap = lambda e, n: torch.as_tensor(e.device_buffer(n), device=dev)
            wrap(self._scorer, "user_factors").copy_(wrap(eng, "user_factors"))       # replicated (deferred steps applied)
            v_all, v_own = wrap(self._scorer, "item_factors"), wrap(eng, "item_factors")
            for r, (a, b) in enumerate(item_slices(width, dist.get_world_size())):
                if r == dist.get_rank():
                    v_all[a * ld:b * ld].copy_(v_own)
                dist.broadcast(v_all[a * ld:b * ld], src=r)
            torch.cuda.synchronize()
            self._scorer_dirty = False
        return self._scorer

    def _full_params(self):
        """All tensors under their TF names; an item-sharded model gathers its slices (host side, API-sized models)."""
        own = self._engine.get_params()
        if self._trainer is None:
            return own
        dist = self._trainer.dist
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, own)
        out = {}
        for n in own:
            if n in ("autoencoder/encoding/kernel", "generator/item_embeddings"):
                out[n] = np.concatenate([q[n] for q in parts], axis=0)
            elif n == "autoencoder/decoding/kernel":
                out[n] = np.concatenate([q[n] for q in parts], axis=1)
            elif n == "autoencoder/decoding/bias":
                out[n] = np.concatenate([q[n] for q in parts], axis=0)
            else:
                out[n] = own[n]
        return out

    def _install_params(self, params):
        if self._trainer is None:
            self._engine.set_params(params)
            return
        lo, hi, _ = self._tp_slice
        sl = {}
        for n, v in params.items():
            v = np.asarray(v)
            if n in ("autoencoder/encoding/kernel", "generator/item_embeddings", "autoencoder/decoding/bias"):
                sl[n] = v[lo:hi]
            elif n == "autoencoder/decoding/kernel":
                sl[n] = v[:, lo:hi]
            else:
                sl[n] = v
        self._engine.set_params(sl)
        self._scorer_dirty = True

    def _compute_item_score(self, user_id_array, items_to_compute=None):   # GANMF.py:285-292
        return self._device_engine().score(np.asarray(user_id_array).reshape(-1))

    def user_factors(self):                                            # GANMF.py:294-297
        return self._device_engine().get_param("generator/user_embeddings")

    def item_factors(self):                                            # GANMF.py:299-302
        return self._device_engine().get_param("generator/item_embeddings")

    # ------------------------------------------------------------------ disk
    def _build_params(self):
        raise NotImplementedError()

    def saveModel(self, folder_path, file_name=None):
        """build_params.pkl (as GANMF.py:309-312) + <name>.npz holding the tensors under their TF
        variable names (the reference writes a TF bundle of the same variables, :313-314)."""
        os.makedirs(folder_path, exist_ok=True)
        with open(os.path.join(folder_path, 'build_params.pkl'), 'wb') as f:
            pickle.dump(self._build_params(), f, pickle.HIGHEST_PROTOCOL)
        name = self.RECOMMENDER_NAME + '_' + self.mode if file_name is None else file_name
        params = self._full_params()
        if _dist_state()[2] == 0:                                      # one writer under torchrun
            np.savez(os.path.join(folder_path, name + '.npz'), **params)

    def loadModel(self, folder_path, file_name=None):
        with open(os.path.join(folder_path, 'build_params.pkl'), 'rb') as f:
            build_params = pickle.load(f)
        self._apply_build_params(build_params)
        eng = self._build_engine(batch_size=getattr(self, '_load_batch', 32))
        name = self.RECOMMENDER_NAME + '_' + self.mode if file_name is None else file_name
        npz = os.path.join(folder_path, name + '.npz')
        tf_data = os.path.join(folder_path, name + '.data-00000-of-00001')
        if os.path.exists(npz):
            z = np.load(npz)
            self._install_params({k: z[k] for k in z.files})
        elif os.path.exists(tf_data):                                  # a model saved by the reference itself
            from ..tf_bundle import read_tf_bundle
            shapes = {n: ((c,) if n.endswith('bias') else (r, c)) for n, r, c, _ in eng.param_infos()}
            if self._trainer is not None:
                lo, hi, width = self._tp_slice
                full = {"autoencoder/encoding/kernel": 0, "generator/item_embeddings": 0, "autoencoder/decoding/bias": 0,
                        "autoencoder/decoding/kernel": 1}
                shapes = {n: tuple(width if (n in full and ax == full[n]) else d for ax, d in enumerate(sh))
                          for n, sh in shapes.items()}
            self._install_params(read_tf_bundle(tf_data, shapes))
        else:
            raise IOError("no saved model %s(.npz|.data-00000-of-00001) in %s" % (name, folder_path))
        if self.mode == 'item':
            self.URM_train = self._URM_users_items.T.tocsr()           # reference leaves it transposed (A.4)

    # north_star's snake-case disk aliases (load_model() without arguments is "restore best", see above)
    def save_model(self, folder_path, file_name=None):
        return self.saveModel(folder_path, file_name)

    def load_model_from(self, folder_path, file_name=None):
        return self.loadModel(folder_path, file_name)

    def set_weights(self, params):
        """Install exported initial weights (parity harness)."""
        self._install_params(params)
        self._engine.reset_optimizers()

    def get_weights(self):
        return self._full_params()
