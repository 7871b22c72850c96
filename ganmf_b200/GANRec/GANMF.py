"""GANMF recommender -- drop-in for the reference's GANRec/GANMF.py on a B200.

Same constructor, fit(**hyperparams) signature and defaults (GANMF.py:26,88-90), the same hooks
(stop_fit, save_current_model, load_model, _compute_item_score, user_factors, item_factors,
autoencoder_codes, saveModel/loadModel) and return value of fit.  The TF graph (MF generator +
linear auto-encoder discriminator, hinge/energy D loss, reconstruction + feature-matching G loss,
two Adam optimisers, GANMF.py:53-139) runs as hand-written sm_100a kernels."""
import numpy as np

from .. import _lib as L
from ._gan_base import GanRecommenderBase


class GANMF(GanRecommenderBase):
    RECOMMENDER_NAME = 'GANMF'
    KIND = L.KIND_GANMF

    def __init__(self, URM_train, mode='user', verbose=False, seed=1234, is_experiment=False):
        self._init_common(URM_train, mode, seed, verbose, is_experiment)

    def build(self, num_factors=10, emb_dim=32):                       # GANMF.py:53-55
        self.num_factors = num_factors
        self.emb_dim = emb_dim

    def _engine_kwargs(self):
        return dict(num_factors=self.num_factors, emb_dim=self.emb_dim)

    def _build_params(self):
        return {'num_factors': self.num_factors, 'emb_dim': self.emb_dim}

    def _apply_build_params(self, bp):
        self.build(**bp)

    def fit(self, num_factors=10, emb_dim=32, epochs=300, batch_size=32, d_lr=1e-4, g_lr=1e-4, d_steps=1, g_steps=1,
            d_reg=0, g_reg=0, m=1, recon_coefficient=1e-2, allow_worse=None, freq=None, after=0, metrics=['MAP'],
            sample_every=None, validation_evaluator=None, validation_set=None, **earlystopping_kwargs):
        self.config = dict(locals())                                   # GANMF.py:93-94
        del self.config['self']
        self.build(num_factors, emb_dim)
        return self._fit_loop(epochs, batch_size, d_lr, g_lr, d_steps, g_steps, d_reg, g_reg, m, recon_coefficient,
                              allow_worse, freq, after, metrics, sample_every, validation_evaluator, validation_set,
                              earlystopping_kwargs)

    def autoencoder_codes(self):                                       # GANMF.py:304-307: R . We + be for all rows
        if self._trainer is not None:
            return self._trainer.encode(np.arange(self.num_users))
        return self._engine.encode(np.arange(self.num_users))
