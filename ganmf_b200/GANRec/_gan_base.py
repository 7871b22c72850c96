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
        self._stop_training = False
        self.train_d_loss, self.train_g_loss = [], []
        self.params = self.best_params = None

    # ------------------------------------------------------------------ engine
    def _engine_kwargs(self):
        raise NotImplementedError()

    def _build_engine(self, batch_size, device=0, gemm_path=L.GEMM_AUTO):
        if self._engine is not None:
            self._engine.close()
        train_rows = self._URM_users_items.T.tocsr() if self.mode == 'item' else self._URM_users_items
        eng = Engine(self.KIND, train_rows.shape[0], train_rows.shape[1], max_batch=int(batch_size),
                     item_mode=(self.mode == 'item'), device=device, gemm_path=gemm_path, **self._engine_kwargs())
        eng.set_csr(L.CSR_TRAIN, train_rows)
        eng.set_csr(L.CSR_SEEN, self._URM_users_items, with_data=False)
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

    def get_URM_train(self):
        return self.URM_train.copy()

    def _compute_item_score(self, user_id_array, items_to_compute=None):   # GANMF.py:285-292
        return self._engine.score(np.asarray(user_id_array).reshape(-1))

    def user_factors(self):                                            # GANMF.py:294-297
        return self._engine.get_param("generator/user_embeddings")

    def item_factors(self):                                            # GANMF.py:299-302
        return self._engine.get_param("generator/item_embeddings")

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
        np.savez(os.path.join(folder_path, name + '.npz'), **self._engine.get_params())

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
            eng.set_params({k: z[k] for k in z.files})
        elif os.path.exists(tf_data):                                  # a model saved by the reference itself
            from ..tf_bundle import read_tf_bundle
            shapes = {n: ((c,) if n.endswith('bias') else (r, c)) for n, r, c, _ in eng.param_infos()}
            eng.set_params(read_tf_bundle(tf_data, shapes))
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
        self._engine.set_params(params)
        self._engine.reset_optimizers()

    def get_weights(self):
        return self._engine.get_params()
