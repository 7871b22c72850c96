"""Shared loaders for the golden fixtures (tests/golden/*.npz)."""
import json
import os

import numpy as np
import scipy.sparse as sps

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_eval_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    n_users, n_items = z["shape"]
    train = sps.csr_matrix((np.ones(len(z["train_indices"]), np.float32), z["train_indices"], z["train_indptr"]),
                           shape=(n_users, n_items))
    test = sps.csr_matrix((z["test_data"].astype(np.float32), z["test_indices"], z["test_indptr"]),
                          shape=(n_users, n_items))
    return dict(scores=z["scores"], train=train, test=test, cutoffs=[int(c) for c in z["cutoffs"]],
                users=z["users"], lists=z["lists"], metric_names=[str(m) for m in z["metric_names"]],
                results=z["results"],
                ignore_items=(z["ignore_items"].astype(np.int64) if "ignore_items" in z.files and
                              len(z["ignore_items"]) else None),
                ignore_users=(z["ignore_users"].astype(np.int64) if "ignore_users" in z.files and
                              len(z["ignore_users"]) else None))


def load_lastfm_kat():
    z = np.load(os.path.join(GOLDEN, "lastfm_kat.npz"))
    n_users, n_items = z["shape"]
    train = sps.csr_matrix((np.ones(len(z["train_indices"]), np.float32), z["train_indices"], z["train_indptr"]),
                           shape=(n_users, n_items))
    test = sps.csr_matrix((np.ones(len(z["test_indices"]), np.float32), z["test_indices"], z["test_indptr"]),
                          shape=(n_users, n_items))
    return dict(user_embeddings=z["user_embeddings"], item_embeddings=z["item_embeddings"], train=train,
                test=test, cutoffs=[int(c) for c in z["cutoffs"]],
                metric_names=[str(m) for m in z["metric_names"]], results=z["results"],
                num_factors=int(z["num_factors"]), emb_dim=int(z["emb_dim"]))


def load_split(ds):
    z = np.load(os.path.join(GOLDEN, "splits_%s.npz" % ds))
    shape = tuple(int(x) for x in z["shape"])
    out = {}
    for part in ("train", "test"):
        idx = z[part + "_indices"].astype(np.int32)
        out[part] = sps.csr_matrix((np.ones(len(idx), np.float32), idx, z[part + "_indptr"]), shape=shape)
    return out


def load_quality_targets():
    return json.load(open(os.path.join(GOLDEN, "quality_targets.json")))
