"""Drop-in quantile forests: same estimator API as scikit-garden's
``RandomForestQuantileRegressor`` / ``ExtraTreesQuantileRegressor``
(skgarden/quantile/ensemble.py:146-484), with ``predict`` / ``apply`` running on a B200.

* ``fit(X, y)`` is scikit-learn's CPU forest fit (the reference's fit is exactly that,
  skgarden/forest.py:637-761) followed by ONE flattening pass that replaces the
  reference's O(T*L*N) bookkeeping loop (ensemble.py:87-101).
* ``predict(X, quantile=None)`` keeps the reference signature and outputs
  (ensemble.py:104-143): ``None`` -> mean (float64), scalar q -> (n,) float64 holding the
  reference's float32 results.  Extension: a sequence of q -> (n, Q) in one pass.
* There is no CPU fallback: predict/apply need the CUDA library and an sm_100 GPU.
* Keyword-only extras: ``device`` (CUDA ordinal), ``devices`` (list of ordinals: the forest is
  replicated and the rows of every ``predict`` / ``apply`` / ``predict_stream`` call are split
  over the GPUs, one host thread each -- results do not depend on the split), ``mode``
  (``"exact"``: bit-identical to the reference; ``"fast"``: within 1e-4), ``builder``.
* Limits of the GPU path (the reference has none of them): dense X only (sparse X raises
  ``NotImplementedError``); the packed node layout holds ``31 - ceil(log2(n_features))`` bits of
  child offset (level-order distance), i.e. trees wider than 2^(31 - feature_bits) nodes on one
  level are refused at ``fit`` with ``QFError`` (QF_ERR_LIMIT); at most 2^31 training rows.
"""
from __future__ import annotations

import numbers
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
from sklearn.exceptions import NotFittedError
from sklearn.utils import check_array, check_X_y

from . import _native
from .flatten import FlatForest, bootstrap_counts, flatten_forest, flatten_forest_device

__all__ = ["RandomForestQuantileRegressor", "ExtraTreesQuantileRegressor"]

# scikit-learn >= 1.1 spellings of the reference's defaults (ensemble.py:282,287)
_CRITERION = {"mse": "squared_error", "mae": "absolute_error"}


def _check_quantiles(quantile):
    """utils.py:42-44 for a scalar or a sequence; returns (list of floats, scalar?)."""
    scalar = isinstance(quantile, numbers.Real) or np.ndim(quantile) == 0
    qs = [float(quantile)] if scalar else [float(q) for q in quantile]
    if not qs:
        raise ValueError("quantile sequence is empty")
    for q in qs:
        if not (0 <= q <= 100):
            raise ValueError("q should be in-between 0 and 100, got %s" % (
                "%d" % q if q == q else "nan"))
    return qs, scalar


class _BaseForestQuantileRegressor(RegressorMixin, BaseEstimator):
    _sk_class = None

    def __init__(self, n_estimators=10, criterion="mse", max_depth=None, min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_features="auto",
                 max_leaf_nodes=None, bootstrap=True, oob_score=False, n_jobs=1,
                 random_state=None, verbose=0, warm_start=False, *, device=0, devices=None,
                 mode="exact", builder="host"):
        self.n_estimators = n_estimators
        self.criterion = criterion
        self.max_depth = max_depth
        self.min_samples_split = min_samples_split
        self.min_samples_leaf = min_samples_leaf
        self.min_weight_fraction_leaf = min_weight_fraction_leaf
        self.max_features = max_features
        self.max_leaf_nodes = max_leaf_nodes
        self.bootstrap = bootstrap
        self.oob_score = oob_score
        self.n_jobs = n_jobs
        self.random_state = random_state
        self.verbose = verbose
        self.warm_start = warm_start
        self.device = device
        self.devices = devices        # list of CUDA ordinals: forest replicated, rows of every call sharded over them
        self.mode = mode
        self.builder = builder        # "host": qf_pack_tree; "device": qf_build_csr (member lists on the GPU)

    # -- fit (CPU, scikit-learn builders) -------------------------------------------------
    def _make_forest(self):
        return self._sk_class(
            n_estimators=self.n_estimators,
            criterion=_CRITERION.get(self.criterion, self.criterion),
            max_depth=self.max_depth, min_samples_split=self.min_samples_split,
            min_samples_leaf=self.min_samples_leaf,
            min_weight_fraction_leaf=self.min_weight_fraction_leaf,
            max_features=1.0 if self.max_features == "auto" else self.max_features,
            max_leaf_nodes=self.max_leaf_nodes, bootstrap=self.bootstrap,
            oob_score=self.oob_score, n_jobs=self.n_jobs, random_state=self.random_state,
            verbose=self.verbose, warm_start=self.warm_start)

    def fit(self, X, y):
        """Build the forest on the CPU, then flatten it for the GPU (ensemble.py:42-102)."""
        X, y = check_X_y(X, y, accept_sparse=False, dtype=np.float32, multi_output=False)   # :79-80
        forest = getattr(self, "_forest", None)
        if not (self.warm_start and forest is not None):
            forest = self._make_forest()
        else:                                     # warm start: every constructor parameter follows set_params
            forest.set_params(**self._make_forest().get_params(deep=False))
        forest.fit(X, y)                                                                    # :81
        self._forest = forest
        self.estimators_ = forest.estimators_
        self.n_features_ = self.n_features_in_ = X.shape[1]
        self.n_outputs_ = 1
        self.y_train_ = y                                                                   # :83
        self._set_flat(self._flatten(X, y))
        return self

    def _flatten(self, X, y) -> FlatForest:
        N = len(y)
        trees = [e.tree_ for e in self.estimators_]
        workers = self._workers()
        if self.builder not in ("host", "device"):
            raise ValueError("builder must be 'host' or 'device'")
        if self.builder == "device":
            counts = [bootstrap_counts(e.random_state, N) for e in self.estimators_] if self.bootstrap else None
            return flatten_forest_device(trees, X, counts, y, X.shape[1], sorter=np.argsort(y),
                                         n_jobs=workers, device=self.device)

        def leaves_of(t):                         # tree.py:101 (all N samples, incl. out-of-bag)
            return trees[t].apply(X)

        if workers > 1 and len(trees) > 1:
            with ThreadPoolExecutor(workers) as ex:        # Tree.apply releases the GIL
                train_leaves = list(ex.map(leaves_of, range(len(trees))))
        else:
            train_leaves = [leaves_of(t) for t in range(len(trees))]
        if self.bootstrap:
            counts = [bootstrap_counts(e.random_state, N) for e in self.estimators_]        # :88-94
        else:
            counts = None
        return flatten_forest(trees, train_leaves, counts, y, X.shape[1],
                              sorter=np.argsort(y), n_jobs=workers)                         # :133

    def _workers(self):
        nj = self.n_jobs
        if nj is None:
            return 1
        if nj < 0:
            return max(1, (os.cpu_count() or 1) + 1 + nj)
        return max(1, nj)

    def _set_flat(self, flat: FlatForest):
        self._flat = flat
        self._device_forest = None
        self._dense_attrs = None

    # -- fitted attributes of the reference, materialised on demand --------------------
    def _dense_fit_attributes(self, trees=None):
        """(T,N) ``y_train_leaves_`` int32 / ``y_weights_`` float32 (ensemble.py:84-101),
        rebuilt from the member lists (out-of-bag: -1 / 0).  ``trees``: only these rows
        (not cached)."""
        if trees is not None or getattr(self, "_dense_attrs", None) is None:
            fl = self._check_fitted()
            which = range(fl.n_trees) if trees is None else list(trees)
            leaves = -np.ones((len(which), fl.n_train), dtype=np.int32)
            weights = np.zeros((len(which), fl.n_train), dtype=np.float32)
            for row, t in enumerate(which):
                n0, n1 = fl.tree_offsets[t], fl.tree_offsets[t + 1]
                m0, m1 = fl.member_offsets[t], fl.member_offsets[t + 1]
                nd = fl.nodes[n0:n1]
                hi = (nd >> np.uint64(32)).astype(np.uint32)
                is_leaf = np.flatnonzero((hi & np.uint32(0x80000000)) != 0)
                # the member lists follow the leaves from left to right = by first member index
                is_leaf = is_leaf[np.argsort((nd[is_leaf] & np.uint64(0xFFFFFFFF)), kind="stable")]
                cnt = (hi[is_leaf] & np.uint32(0x7FFFFFFF)).astype(np.int64)
                is_leaf = is_leaf[cnt > 0]
                cnt = cnt[cnt > 0]
                ids = (np.arange(n1 - n0, dtype=np.int32) if fl.node_ids is None
                       else fl.node_ids[n0:n1])[is_leaf]
                mem = fl.members[m0:m1]
                sample = fl.sorter[(mem & np.uint64(0xFFFFFFFF)).astype(np.int64)]
                leaves[row, sample] = np.repeat(ids, cnt)
                weights[row, sample] = (mem >> np.uint64(32)).astype(np.uint32).view(np.float32)
            if trees is not None:
                return leaves, weights
            self._dense_attrs = (leaves, weights)
        return self._dense_attrs

    # fitted attributes the reference classes inherit from scikit-learn's forest
    @property
    def oob_score_(self):
        return self._fitted_forest().oob_score_

    @property
    def oob_prediction_(self):
        return self._fitted_forest().oob_prediction_

    @property
    def feature_importances_(self):
        return self._fitted_forest().feature_importances_

    def _fitted_forest(self):
        forest = getattr(self, "_forest", None)
        if forest is None:
            raise NotFittedError("This %s instance is not fitted yet." % type(self).__name__)
        return forest

    @property
    def y_train_leaves_(self):
        return self._dense_fit_attributes()[0]

    @property
    def y_weights_(self):
        return self._dense_fit_attributes()[1]

    # -- device ------------------------------------------------------------------------------
    def _check_fitted(self) -> FlatForest:
        flat = getattr(self, "_flat", None)
        if flat is None:
            raise NotFittedError("This %s instance is not fitted yet." % type(self).__name__)
        return flat

    def _on_device(self):
        flat = self._check_fitted()
        if getattr(self, "_device_forest", None) is None:
            from .device_forest import DeviceForest, ForestReplicas
            devices = getattr(self, "devices", None)
            if devices is not None:
                self._device_forest = ForestReplicas(flat, devices, with_impurity=self._needs_impurity)
            else:
                self._device_forest = DeviceForest(flat, device=self.device, with_impurity=self._needs_impurity)
        return self._device_forest

    _needs_impurity = False            # the return_std forests (forest.py) upload tree_.impurity too

    def _mode(self):
        if self.mode not in ("exact", "fast"):
            raise ValueError("mode must be 'exact' or 'fast'")
        return _native.QF_MODE_EXACT if self.mode == "exact" else _native.QF_MODE_FAST

    def _validate_X(self, X):
        """ensemble.py:129 (`check_array(X, dtype=float32)`: casts, rejects NaN / inf and 0 rows) with the
        finite scan of large inputs split over host threads (it is as long as the GPU work otherwise)."""
        if hasattr(X, "tocsr"):
            raise NotImplementedError("sparse X is not supported on the GPU path")
        self._check_fitted()
        big = np.size(X) >= (1 << 22)
        try:
            X = check_array(X, dtype=np.float32, order="C", ensure_all_finite=not big)
        except TypeError:                                                # scikit-learn < 1.6 spelling
            X = check_array(X, dtype=np.float32, order="C", force_all_finite=not big)
        if X.shape[1] != self.n_features_:
            raise ValueError("X has %d features, but %s is expecting %d features as input."
                             % (X.shape[1], type(self).__name__, self.n_features_))
        if big:
            from .device_forest import assert_all_finite_parallel
            assert_all_finite_parallel(X, workers=min(16, os.cpu_count() or 1))
        return X

    # -- the hot path --------------------------------------------------------------------------
    def predict(self, X, quantile=None):
        """ensemble.py:104-143.  ``quantile`` in [0, 100] (int or float); None = mean."""
        X = self._validate_X(X)
        dev = self._on_device()
        if quantile is None:                                             # ensemble.py:130-131
            return dev.predict_mean_host(X)
        qs, scalar = _check_quantiles(quantile)
        out = dev.predict_quantiles_host(X, qs, mode=self._mode())
        return out[:, 0] if scalar else out

    def predict_stream(self, chunks, quantile):
        """Extension for test sets that do not fit one array (BASELINE.json config 5): an
        iterable of (m_i, F) chunks in, one ``predict(chunk, quantile)`` result per chunk out
        (generator), each chunk pipelined H2D / kernels / D2H on the device."""
        self._check_fitted()
        qs, scalar = _check_quantiles(quantile)
        dev = self._on_device()

        def prepare(X):                           # runs in the streaming worker threads
            if np.shape(X)[0] == 0:
                return np.empty((0, self.n_features_), dtype=np.float32)        # an empty chunk is fine in a stream
            return check_array(X, dtype=np.float32, order="C")

        def shaped():
            for X in chunks:
                if hasattr(X, "tocsr"):
                    raise NotImplementedError("sparse X is not supported on the GPU path")
                if np.ndim(X) != 2 or np.shape(X)[1] != self.n_features_:
                    self._validate_X(X)           # raises the reference's error for the shape
                yield X
        for out in dev.predict_stream(shaped(), qs, mode=self._mode(), prepare=prepare):
            yield out[:, 0] if scalar else out

    def apply(self, X):
        """forest.py:583-604: (n, T) int64 scikit-learn leaf ids."""
        X = self._validate_X(X)
        return self._on_device().apply_host(X)

    # -- persistence: the device handle never travels ----------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state["_device_forest"] = None
        state["_dense_attrs"] = None
        return state


class RandomForestQuantileRegressor(_BaseForestQuantileRegressor):
    """Random-forest quantile regressor (skgarden/quantile/ensemble.py:146-315)."""
    _sk_class = RandomForestRegressor


class ExtraTreesQuantileRegressor(_BaseForestQuantileRegressor):
    """Extra-trees quantile regressor (skgarden/quantile/ensemble.py:318-484).  Note the
    code default ``bootstrap=True`` (ensemble.py:458), kept here."""
    _sk_class = ExtraTreesRegressor
