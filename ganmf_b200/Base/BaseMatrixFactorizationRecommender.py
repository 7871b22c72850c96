"""BaseMatrixFactorizationRecommender of the reference (Base/BaseMatrixFactorizationRecommender.py:76-143) with the
scorer on the device: any recommender that ends up with USER_factors [n_users, k] and ITEM_factors [n_items, k]
(PureSVD, IALS, MF-BPR, ... of the reference's baselines; out of scope here, SURVEY.md section 8f-4 names only their
common scoring base) ranks and evaluates through the same kernels as GANMF -- score GEMM / fused score-select,
seen mask, top-k, metric sums -- instead of shipping host score rows over PCIe.

Same attributes and semantics: `use_bias` adds ITEM_bias + GLOBAL_bias + USER_bias (:119-122), users without
training interactions ("cold", :124-139) get -inf for ALL items (the item-KNN estimate for cold users,
set_URM_train(..., estimate_model_for_cold_users=True), is not part of the hot path and raises).  With biases or
cold users in the request the score rows are edited on the host exactly as the reference does and re-enter the device
mask -> top-k stage; otherwise nothing but ids and lists crosses PCIe."""
import numpy as np
import scipy.sparse as sps

from .. import _lib as L
from ..engine import Engine
from .BaseRecommender import BaseRecommender


class BaseMatrixFactorizationRecommender(BaseRecommender):
    RECOMMENDER_NAME = "BaseMatrixFactorizationRecommender"

    def __init__(self, URM_train):
        URM_train = sps.csr_matrix(URM_train, dtype=np.float32)
        URM_train.eliminate_zeros()
        super(BaseMatrixFactorizationRecommender, self).__init__(URM_train)
        self.n_users, self.n_items = self.URM_train.shape
        self.use_bias = False
        self._cold_user_mask = np.ediff1d(self.URM_train.indptr) == 0          # BaseRecommender.py:36
        self._cold_user_KNN_model_available = False
        self._engine = None
        self._engine_key = None

    def _get_cold_user_mask(self):
        return self._cold_user_mask

    # ------------------------------------------------------------------ device scorer
    def _device_engine(self):
        """Factor-only device context; rebuilt when the factor arrays are replaced (e.g. by another fit())."""
        U, V = getattr(self, "USER_factors", None), getattr(self, "ITEM_factors", None)
        if U is None or V is None:
            return None
        assert U.shape[1] == V.shape[1], \
            "{}: User and Item factors have inconsistent shape".format(self.RECOMMENDER_NAME)       # :105-106
        key = (id(U), id(V), U.shape, V.shape)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            eng = Engine(L.KIND_MF, U.shape[0], V.shape[0], U.shape[1], max_batch=1)
            eng.set_csr(L.CSR_SEEN, self.URM_train, with_data=False)
            eng.set_param("generator/user_embeddings", np.asarray(U, dtype=np.float32))
            eng.set_param("generator/item_embeddings", np.asarray(V, dtype=np.float32))
            self._engine, self._engine_key = eng, key
        return self._engine

    def _scores_need_host_edit(self, user_id_array=None):
        """True when the ranking cannot be left to the raw factor product: biases, or cold users among the request."""
        if self.use_bias:
            return True
        cold = self._cold_user_mask if user_id_array is None else self._cold_user_mask[np.asarray(user_id_array)]
        return bool(np.any(cold))

    def _compute_item_score(self, user_id_array, items_to_compute=None):          # :94-143
        user_id_array = np.asarray(user_id_array).reshape(-1)
        eng = self._device_engine()
        if eng is None:
            raise RuntimeError("%s: USER_factors / ITEM_factors are not set (call fit() first)" % self.RECOMMENDER_NAME)
        assert self.USER_factors.shape[0] > user_id_array.max(), \
            "{}: Cold users not allowed. Users in trained model are {}, requested prediction for users up to {}".format(
                self.RECOMMENDER_NAME, self.USER_factors.shape[0], user_id_array.max())            # :108-110
        item_scores = eng.score(user_id_array)
        if items_to_compute is not None:                                          # :112-114
            keep = np.zeros(item_scores.shape[1], dtype=bool)
            keep[np.asarray(items_to_compute)] = True
            item_scores[:, ~keep] = -np.inf
        if self.use_bias:                                                         # :119-122
            item_scores += self.ITEM_bias + self.GLOBAL_bias
            item_scores = (item_scores.T + self.USER_bias[user_id_array]).T
        cold = self._cold_user_mask[user_id_array]
        if cold.any():                                                            # :124-139
            if self._cold_user_KNN_model_available:
                raise NotImplementedError("item-KNN estimate for cold users is outside the GANMF hot path")
            item_scores[cold, :] = -np.inf
        return item_scores

    def set_URM_train(self, URM_train_new, estimate_model_for_cold_users=False, **kwargs):       # :148-170
        assert self.URM_train.shape == URM_train_new.shape, \
            "{}: set_URM_train old and new URM train have different shapes".format(self.RECOMMENDER_NAME)
        if estimate_model_for_cold_users:
            raise NotImplementedError("item-KNN estimate for cold users is outside the GANMF hot path")
        self.URM_train = sps.csr_matrix(URM_train_new.copy(), dtype=np.float32)
        self.URM_train.eliminate_zeros()
        self._cold_user_mask = np.ediff1d(self.URM_train.indptr) == 0
        if self._engine is not None:
            self._engine.set_csr(L.CSR_SEEN, self.URM_train, with_data=False)

    def saveModel(self, folder_path, file_name=None):                              # :200-230 (npz instead of DataIO zip)
        import os
        name = self.RECOMMENDER_NAME if file_name is None else file_name
        os.makedirs(folder_path, exist_ok=True)
        data = {"USER_factors": self.USER_factors, "ITEM_factors": self.ITEM_factors, "use_bias": self.use_bias,
                "_cold_user_mask": self._cold_user_mask}
        if self.use_bias:
            data.update(ITEM_bias=self.ITEM_bias, USER_bias=self.USER_bias, GLOBAL_bias=self.GLOBAL_bias)
        np.savez(os.path.join(folder_path, name + ".npz"), **data)

    def loadModel(self, folder_path, file_name=None):
        import os
        name = self.RECOMMENDER_NAME if file_name is None else file_name
        z = np.load(os.path.join(folder_path, name + ".npz"))
        for k in z.files:
            v = z[k]
            setattr(self, k, bool(v) if k == "use_bias" else (float(v) if v.ndim == 0 else v))
