"""Drop-in ``return_std`` forests: same estimator API as scikit-garden's
``RandomForestRegressor`` / ``ExtraTreesRegressor`` (skgarden/forest.py:181-548), with
``predict(X, return_std=False)`` running on a B200.

``std(y | x)`` is E[Var(y | tree)] + Var(E[y | tree]) over the trees (forest.py:133-178):
per tree the impurity (variance) and value (mean) of the leaf the row falls in, impurities
clamped from below by ``min_variance``.  The leaf lookup is the K1 kernel of the quantile
path; the reduction is ``mean_std_kernel`` (float64, trees in increasing order).
"""
from __future__ import annotations

import numpy as np
from sklearn.ensemble import ExtraTreesRegressor as _SkExtraTrees
from sklearn.ensemble import RandomForestRegressor as _SkRandomForest

from .ensemble import _CRITERION, _BaseForestQuantileRegressor

__all__ = ["RandomForestRegressor", "ExtraTreesRegressor"]


class _BaseStdForestRegressor(_BaseForestQuantileRegressor):
    _needs_impurity = True

    def __init__(self, n_estimators=10, criterion="mse", max_depth=None, min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_features="auto",
                 max_leaf_nodes=None, bootstrap=True, oob_score=False, n_jobs=1,
                 random_state=None, verbose=0, warm_start=False, min_variance=0.0, *, device=0,
                 devices=None, builder="host"):
        super().__init__(n_estimators=n_estimators, criterion=criterion, max_depth=max_depth,
                         min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf,
                         min_weight_fraction_leaf=min_weight_fraction_leaf, max_features=max_features,
                         max_leaf_nodes=max_leaf_nodes, bootstrap=bootstrap, oob_score=oob_score,
                         n_jobs=n_jobs, random_state=random_state, verbose=verbose,
                         warm_start=warm_start, device=device, devices=devices, builder=builder)
        self.min_variance = min_variance

    def predict(self, X, return_std=False):
        """forest.py:332-362.  Returns the forest mean, or ``(mean, std)`` with ``return_std``."""
        X = self._validate_X(X)
        dev = self._on_device()
        if not return_std:
            return dev.predict_mean_host(X)
        if _CRITERION.get(self.criterion, self.criterion) != "squared_error":    # forest.py:355-358
            raise ValueError("Expected impurity to be 'mse', got %s instead" % self.criterion)
        return dev.predict_mean_std_host(X, self.min_variance)


class RandomForestRegressor(_BaseStdForestRegressor):
    """Random forest with conditional std (skgarden/forest.py:181-362)."""
    _sk_class = _SkRandomForest


class ExtraTreesRegressor(_BaseStdForestRegressor):
    """Extra trees with conditional std (skgarden/forest.py:365-548).  The reference's default
    is ``bootstrap=False`` here (forest.py:500), unlike the quantile variant."""
    _sk_class = _SkExtraTrees

    def __init__(self, n_estimators=10, criterion="mse", max_depth=None, min_samples_split=2,
                 min_samples_leaf=1, min_weight_fraction_leaf=0.0, max_features="auto",
                 max_leaf_nodes=None, bootstrap=False, oob_score=False, n_jobs=1,
                 random_state=None, verbose=0, warm_start=False, min_variance=0.0, *, device=0,
                 devices=None, builder="host"):
        super().__init__(n_estimators=n_estimators, criterion=criterion, max_depth=max_depth,
                         min_samples_split=min_samples_split, min_samples_leaf=min_samples_leaf,
                         min_weight_fraction_leaf=min_weight_fraction_leaf, max_features=max_features,
                         max_leaf_nodes=max_leaf_nodes, bootstrap=bootstrap, oob_score=oob_score,
                         n_jobs=n_jobs, random_state=random_state, verbose=verbose,
                         warm_start=warm_start, min_variance=min_variance, device=device, devices=devices,
                         builder=builder)
