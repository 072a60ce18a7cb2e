"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED scikit-garden reference.

Only ``tests/``, ``tests/golden/make_golden.py`` and the ``cpu_baseline`` leg of
``bench.py`` may import this module.  Nothing on the product path does.

The reference (``/root/reference``, scikit-garden 0.1.x) was written against
scikit-learn 0.22 / numpy 1.x.  Its hot-path files

    skgarden/quantile/ensemble.py   (BaseForestQuantileRegressor.fit/predict)
    skgarden/quantile/utils.py      (weighted_percentile)
    skgarden/quantile/tree.py       (tree-level quantile regressors)
    skgarden/forest.py              (vendored BaseForest.apply / ForestRegressor.predict)

run *unmodified* under the in-process shims below.  No
reference file is copied or edited; the shims only patch names that newer
scikit-learn removed.  ``/root/reference`` exists only in the build container,
never on the GPU box, so every caller must go through :func:`available`.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("QF_REFERENCE_ROOT", "/root/reference")

# sklearn >= 1.1 spells the reference's defaults differently (Appendix B, shim 6):
# criterion='mse' -> 'squared_error', max_features='auto' -> 1.0 (all features).
MODERN_KW = dict(criterion="squared_error", max_features=1.0)

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "skgarden", "quantile", "ensemble.py"))


def load():
    """Import ``skgarden.quantile`` from the reference tree and return the module.

    Returns a namespace with RandomForestQuantileRegressor, ExtraTreesQuantileRegressor,
    DecisionTreeQuantileRegressor, ExtraTreeQuantileRegressor, weighted_percentile.
    """
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    import sklearn.tree._tree as _t
    import sklearn.utils.fixes as _f
    from sklearn.ensemble import BaseEnsemble
    from sklearn.tree import DecisionTreeRegressor
    from sklearn.utils import check_X_y

    # shim 1: forest.py:8 imports DTYPE / DOUBLE, removed from sklearn.tree._tree
    if not hasattr(_t, "DTYPE"):
        _t.DTYPE = np.float32
    if not hasattr(_t, "DOUBLE"):
        _t.DOUBLE = np.float64
    # shim 2: forest.py:19 imports _joblib_parallel_args (now plain joblib kwargs)
    if not hasattr(_f, "_joblib_parallel_args"):
        _f._joblib_parallel_args = lambda **kw: kw
    # shim 4: forest.py:569-572 passes base_estimator=, renamed to estimator=
    if not getattr(BaseEnsemble.__init__, "_qf_shim", False):
        _orig_init = BaseEnsemble.__init__

        def _init(self, estimator=None, *, base_estimator=None, n_estimators=10,
                  estimator_params=tuple()):
            _orig_init(self, estimator if estimator is not None else base_estimator,
                       n_estimators=n_estimators, estimator_params=estimator_params)

        _init._qf_shim = True
        BaseEnsemble.__init__ = _init
    # shim 3: skgarden/__init__.py imports the Mondrian Cython modules (unbuildable
    # here and off the hot path) -> register a stub package that only has a path.
    if "skgarden" not in sys.modules:
        pkg = types.ModuleType("skgarden")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "skgarden")]
        sys.modules["skgarden"] = pkg

    import skgarden.quantile as Q  # noqa: E402  (the reference package itself)
    import skgarden.quantile.tree as QT  # noqa: E402

    # shim 5: sklearn >= 1.0 DecisionTreeRegressor.fit bypasses the reference's
    # BaseTreeQuantileRegressor.fit (tree.py:49-102), so y_train_/y_train_leaves_
    # would never be set.  Re-attach the same bookkeeping (tree.py:88-101).
    def _tree_fit(self, X, y, sample_weight=None, check_input=True, X_idx_sorted=None):
        y = np.asarray(y)
        if np.ndim(y) == 2 and y.shape[1] == 1:
            y = np.ravel(y)
        X, y = check_X_y(X, y, accept_sparse="csc", dtype=np.float32, multi_output=False)
        DecisionTreeRegressor.fit(self, X, y, sample_weight=sample_weight,
                                  check_input=check_input)
        self.y_train_ = y
        self.y_train_leaves_ = self.tree_.apply(X)
        return self

    QT.DecisionTreeQuantileRegressor.fit = _tree_fit
    QT.ExtraTreeQuantileRegressor.fit = _tree_fit

    ns = types.SimpleNamespace(
        module=Q,
        RandomForestQuantileRegressor=Q.RandomForestQuantileRegressor,
        ExtraTreesQuantileRegressor=Q.ExtraTreesQuantileRegressor,
        DecisionTreeQuantileRegressor=Q.DecisionTreeQuantileRegressor,
        ExtraTreeQuantileRegressor=Q.ExtraTreeQuantileRegressor,
        weighted_percentile=Q.utils.weighted_percentile,
        MODERN_KW=MODERN_KW,
    )
    _loaded = ns
    return ns
