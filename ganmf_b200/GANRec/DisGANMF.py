"""DisGANMF recommender -- drop-in for the reference's GANRec/DisGANMF.py on a B200.

Ablation of GANMF with a binary-classifier MLP discriminator on concat([float(id), profile])
(DisGANMF.py:57-65), BCE losses + feature matching (DisGANMF.py:114-136).  Note the reference's
positional order (URM_train, mode, seed, verbose, is_experiment) differs from GANMF's."""
from .. import _lib as L
from ._gan_base import GanRecommenderBase


class DisGANMF(GanRecommenderBase):
    RECOMMENDER_NAME = 'DisGANMF'
    KIND = L.KIND_DISGANMF

    def __init__(self, URM_train, mode='user', seed=1234, verbose=False, is_experiment=False):
        self._init_common(URM_train, mode, seed, verbose, is_experiment)

    def build(self, d_layers=1, d_nodes=32, d_hidden_act='linear', num_factors=10):   # DisGANMF.py:51
        if d_hidden_act not in L.ACT:
            raise ValueError("unsupported d_hidden_act %r" % (d_hidden_act,))
        self.d_layers, self.d_nodes, self.d_hidden_act, self.num_factors = d_layers, d_nodes, d_hidden_act, num_factors

    def _engine_kwargs(self):
        return dict(num_factors=self.num_factors, d_layers=self.d_layers, d_nodes=self.d_nodes,
                    d_act=self.d_hidden_act)

    def _build_params(self):
        return {'d_layers': self.d_layers, 'd_nodes': self.d_nodes, 'd_hidden_act': self.d_hidden_act,
                'num_factors': self.num_factors}

    def _apply_build_params(self, bp):
        self.build(**bp)

    def fit(self, num_factors=10, d_layers=1, d_nodes=32, d_hidden_act='linear', epochs=300, batch_size=32, d_lr=1e-4,
            g_lr=1e-4, d_steps=1, g_steps=1, d_reg=0, g_reg=0, recon_coefficient=1e-2, allow_worse=None, freq=None,
            after=0, metrics=['MAP'], sample_every=None, validation_evaluator=None, validation_set=None,
            **earlystopping_kwargs):
        self.config = dict(locals())                                   # DisGANMF.py:88-89
        del self.config['self']
        self.build(num_factors=num_factors, d_layers=d_layers, d_nodes=d_nodes, d_hidden_act=d_hidden_act)
        return self._fit_loop(epochs, batch_size, d_lr, g_lr, d_steps, g_steps, d_reg, g_reg, 1.0, recon_coefficient,
                              allow_worse, freq, after, metrics, sample_every, validation_evaluator, validation_set,
                              earlystopping_kwargs)

    def saveModel(self, folder_path, file_name):                       # DisGANMF.py:264 (file_name required)
        return super(DisGANMF, self).saveModel(folder_path, file_name)
