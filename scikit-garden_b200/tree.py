"""Drop-in tree-level quantile regressors: same estimator API as scikit-garden's
``DecisionTreeQuantileRegressor`` / ``ExtraTreeQuantileRegressor``
(skgarden/quantile/tree.py:10-246), with ``predict`` / ``apply`` running on a B200.

The reference predicts a quantile per leaf with ``weighted_percentile(y_train_[leaf], q)``
(tree.py:41-47): unit weights over ALL training samples of the leaf (there is no bootstrap
at tree level), own ``argsort`` of the float32 values (utils.py:59-62; the order among
equal values cannot change the result because their weights are equal).  On the device this
is the forest path with T = 1 and every member weight 1.0 -- same kernels, same float32
arithmetic (utils.py:68-81).
"""
from __future__ import annotations

import numpy as np
from sklearn.base import BaseEstimator, RegressorMixin
from sklearn.exceptions import NotFittedError
from sklearn.tree import DecisionTreeRegressor, ExtraTreeRegressor
from sklearn.utils import check_array, check_X_y

from . import _native
from .ensemble import _CRITERION, _check_quantiles
from .flatten import FlatForest, flatten_forest

__all__ = ["DecisionTreeQuantileRegressor", "ExtraTreeQuantileRegressor"]


class _BaseTreeQuantileRegressor(RegressorMixin, BaseEstimator):
    _sk_class = None
    _auto_max_features = None          # what the reference's default max_features means today

    def __init__(self, criterion="mse", splitter="best", max_depth=None, min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_features=None,
                 random_state=None, max_leaf_nodes=None, *, device=0, mode="exact"):
        self.criterion = criterion
        self.splitter = splitter
        self.max_depth = max_depth
        self.min_samples_split = min_samples_split
        self.min_samples_leaf = min_samples_leaf
        self.min_weight_fraction_leaf = min_weight_fraction_leaf
        self.max_features = max_features
        self.random_state = random_state
        self.max_leaf_nodes = max_leaf_nodes
        self.device = device
        self.mode = mode

    def fit(self, X, y, sample_weight=None, check_input=True):
        """tree.py:50-100: scikit-learn's CPU tree fit, then one flattening pass."""
        y = np.asarray(y)
        if np.ndim(y) == 2 and y.shape[1] == 1:                         # tree.py:87-89
            y = np.ravel(y)
        X, y = check_X_y(X, y, accept_sparse=False, dtype=np.float32, multi_output=False)   # :92-93
        tree = self._sk_class(
            criterion=_CRITERION.get(self.criterion, self.criterion), splitter=self.splitter,
            max_depth=self.max_depth, min_samples_split=self.min_samples_split,
            min_samples_leaf=self.min_samples_leaf,
            min_weight_fraction_leaf=self.min_weight_fraction_leaf,
            max_features=self._auto_max_features if self.max_features == "auto" else self.max_features,
            max_leaf_nodes=self.max_leaf_nodes, random_state=self.random_state)
        tree.fit(X, y, sample_weight=sample_weight, check_input=check_input)                # :94-96
        self._tree = tree
        self.tree_ = tree.tree_
        self.n_features_ = self.n_features_in_ = X.shape[1]
        self.n_outputs_ = 1
        self.y_train_ = y                                                                   # :97
        self.y_train_leaves_ = self.tree_.apply(X)                                          # :100
        self._flat = flatten_forest([self.tree_], [self.y_train_leaves_], None, y, X.shape[1],
                                    sorter=np.argsort(y), n_jobs=1, unit_weights=True)
        self._device_forest = None
        return self

    # -- device ------------------------------------------------------------------------------
    def _check_fitted(self) -> FlatForest:
        flat = getattr(self, "_flat", None)
        if flat is None:
            raise NotFittedError("This %s instance is not fitted yet." % type(self).__name__)
        return flat

    def _on_device(self):
        flat = self._check_fitted()
        if getattr(self, "_device_forest", None) is None:
            from .device_forest import DeviceForest
            self._device_forest = DeviceForest(flat, device=self.device)
        return self._device_forest

    def _validate_X(self, X):
        if hasattr(X, "tocsr"):
            raise NotImplementedError("sparse X is not supported on the GPU path")
        self._check_fitted()
        X = check_array(X, dtype=np.float32, order="C")                  # tree.py:37
        if X.shape[1] != self.n_features_:
            raise ValueError("X has %d features, but %s is expecting %d features as input."
                             % (X.shape[1], type(self).__name__, self.n_features_))
        return X

    def predict(self, X, quantile=None, check_input=False):
        """tree.py:12-47.  ``quantile`` in [0, 100]; None = the leaf mean (tree_.value)."""
        import torch
        X = self._validate_X(X)
        dev = self._on_device()
        if quantile is None:                                             # tree.py:38-39
            with torch.cuda.device(dev.device):
                xd = torch.from_numpy(X).to("cuda:%d" % dev.device)
                return dev.predict_mean(xd).cpu().numpy()
        if self.mode not in ("exact", "fast"):
            raise ValueError("mode must be 'exact' or 'fast'")
        mode = _native.QF_MODE_EXACT if self.mode == "exact" else _native.QF_MODE_FAST
        qs, scalar = _check_quantiles(quantile)
        out = dev.predict_quantiles_host(X, qs, mode=mode)
        return out[:, 0] if scalar else out

    def apply(self, X):
        """(n,) int64 scikit-learn leaf ids (sklearn BaseDecisionTree.apply, tree.py:42)."""
        import torch
        X = self._validate_X(X)
        dev = self._on_device()
        with torch.cuda.device(dev.device):
            xd = torch.from_numpy(X).to("cuda:%d" % dev.device)
            return dev.apply(xd).cpu().numpy()[:, 0]

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_device_forest"] = None
        return state


class DecisionTreeQuantileRegressor(_BaseTreeQuantileRegressor):
    """Decision-tree quantile regressor (skgarden/quantile/tree.py:103-223)."""
    _sk_class = DecisionTreeRegressor
    _auto_max_features = None


class ExtraTreeQuantileRegressor(_BaseTreeQuantileRegressor):
    """Extra-tree quantile regressor (skgarden/quantile/tree.py:226-246); the reference's
    defaults are splitter='random', max_features='auto' (= all features for regression)."""
    _sk_class = ExtraTreeRegressor
    _auto_max_features = 1.0

    def __init__(self, criterion="mse", splitter="random", max_depth=None, min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_features="auto",
                 random_state=None, max_leaf_nodes=None, *, device=0, mode="exact"):
        super().__init__(criterion=criterion, splitter=splitter, max_depth=max_depth,
                         min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf,
                         min_weight_fraction_leaf=min_weight_fraction_leaf, max_features=max_features,
                         random_state=random_state, max_leaf_nodes=max_leaf_nodes, device=device,
                         mode=mode)
