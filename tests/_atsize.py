"""Helpers for parity tests at sizes where the oracle cannot index the whole training set:
oracle.SparseIndex on fit attributes thinned to the leaves the sampled rows reach (entries of all
other leaves are masked to -1 = "out of bag", which never matches)."""
import numpy as np

from oracle import qrf_oracle as O


def thin_index(est, X_leaves):
    leaves, weights = est.y_train_leaves_, est.y_weights_
    thin = np.full_like(leaves, -1)
    for t in range(leaves.shape[0]):
        keep = np.isin(leaves[t], np.unique(X_leaves[:, t]))
        thin[t, keep] = leaves[t, keep]
    return O.SparseIndex(est.y_train_, thin, weights, est._flat.sorter)


def make_data(N, n, F, seed, ties=False):
    rng = np.random.RandomState(seed)
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    if ties:
        y = np.round(y, 1)
    return X[:N], y[:N], X[N:]
