"""Host side of the drop-in estimators (no GPU): constructor/params parity with
skgarden/quantile/ensemble.py:280-315 / :449-484, fit bookkeeping parity with
ensemble.py:83-101 (ported from skgarden/quantile/tests/test_ensemble.py:24-48),
validation and persistence."""
import inspect
import pickle

import numpy as np
import pytest
from numpy.testing import assert_array_equal
from sklearn.base import clone
from sklearn.exceptions import NotFittedError

import _golden
import skgarden_b200 as sg
from skgarden_b200.ensemble import _check_quantiles

REFERENCE_SIGNATURE = ["n_estimators", "criterion", "max_depth", "min_samples_split",
                       "min_samples_leaf", "min_weight_fraction_leaf", "max_features",
                       "max_leaf_nodes", "bootstrap", "oob_score", "n_jobs", "random_state",
                       "verbose", "warm_start"]


@pytest.mark.parametrize("cls", [sg.RandomForestQuantileRegressor, sg.ExtraTreesQuantileRegressor])
def test_constructor_matches_reference(cls):
    sig = inspect.signature(cls.__init__)
    pos = [p.name for p in sig.parameters.values()
           if p.kind == p.POSITIONAL_OR_KEYWORD and p.name != "self"]
    assert pos == REFERENCE_SIGNATURE                       # ensemble.py:280-294
    est = cls()
    assert est.n_estimators == 10 and est.criterion == "mse" and est.max_features == "auto"
    assert est.bootstrap is True and est.n_jobs == 1        # ETQR default is True too (:458)
    est2 = clone(est.set_params(n_estimators=3, bootstrap=False))
    assert est2.get_params()["n_estimators"] == 3 and est2.bootstrap is False


def boston():
    g = _golden.load("c1_rfqr_boston")
    return g.X_train, g.y_train, g.X_test


@pytest.mark.parametrize("cls", [sg.RandomForestQuantileRegressor, sg.ExtraTreesQuantileRegressor])
def test_quantile_attributes(cls):                          # test_ensemble.py:24-48
    X_train, y_train, _ = boston()
    est = cls(random_state=0).fit(X_train, y_train)
    assert_array_equal(np.vstack(np.where(est.y_train_leaves_ == -1)),
                       np.vstack(np.where(est.y_weights_ == 0)))
    assert_array_equal(np.sum(est.y_weights_, axis=1),
                       [sum(tree.tree_.children_left == -1) for tree in est.estimators_])
    est.set_params(bootstrap=False)
    est.fit(X_train, y_train)
    assert_array_equal(np.sum(est.y_weights_, axis=1),
                       [sum(tree.tree_.children_left == -1) for tree in est.estimators_])
    assert np.all(est.y_train_leaves_ != -1)
    assert est.y_train_leaves_.dtype == np.int32 and est.y_weights_.dtype == np.float32


@pytest.mark.parametrize("name,cls", [("c1_rfqr_boston", sg.RandomForestQuantileRegressor),
                                      ("c1_etqr_boston", sg.ExtraTreesQuantileRegressor)])
def test_refit_reproduces_reference_fit(name, cls):
    # same data + seed -> the same trees and fit attributes as the reference produced
    g = _golden.load(name)
    est = cls(n_estimators=10, random_state=0).fit(g.X_train, g.y_train)
    assert_array_equal(np.concatenate([e.tree_.threshold for e in est.estimators_]), g.threshold)
    assert_array_equal(est.y_train_leaves_, g.y_train_leaves_)
    assert_array_equal(est.y_weights_, g.y_weights_)
    assert_array_equal(est.y_train_, g.y_train)
    assert est.n_features_ == 13 and est.n_outputs_ == 1


def test_base_estimators_and_refit_invalidation():          # test_ensemble.py:86-102 (classes differ by design)
    rng = np.random.RandomState(0)
    X, y = rng.randn(10, 1), np.linspace(0.0, 100.0, 10)
    est = sg.RandomForestQuantileRegressor(random_state=0, max_depth=1).fit(X, y)
    assert len(est.estimators_) == 10
    first = est._flat
    est.fit(X, y)
    assert est._flat is not first and est._device_forest is None


def test_validation_errors():
    X_train, y_train, X_test = boston()
    est = sg.RandomForestQuantileRegressor(n_estimators=3, random_state=0)
    with pytest.raises(NotFittedError):
        est.predict(X_test, quantile=50)
    est.fit(X_train, y_train)
    with pytest.raises(ValueError, match="q should be in-between 0 and 100, got 101"):
        _check_quantiles(101)                               # utils.py:42-44
    with pytest.raises(ValueError):
        _check_quantiles([50, -1])
    assert _check_quantiles(97.5) == ([97.5], True)
    assert _check_quantiles([5, 50]) == ([5.0, 50.0], False)
    with pytest.raises(ValueError):
        est._validate_X(X_test[:, :5])                      # wrong width
    bad = X_test.copy()
    bad[0, 0] = np.nan
    with pytest.raises(ValueError):
        est._validate_X(bad)                                # check_array rejects NaN (ensemble.py:129)
    with pytest.raises(ValueError):
        est._validate_X(X_test[:0])                         # 0 rows
    import scipy.sparse as sp
    with pytest.raises(NotImplementedError):
        est._validate_X(sp.csr_matrix(X_test))
    with pytest.raises(ValueError):
        est.fit(X_train, np.stack([y_train, y_train], axis=1))   # multi_output=False (:79-80)


def test_pickle_drops_device_state():
    X_train, y_train, _ = boston()
    est = sg.ExtraTreesQuantileRegressor(n_estimators=4, random_state=1).fit(X_train, y_train)
    est._device_forest = object()
    back = pickle.loads(pickle.dumps(est))
    assert back._device_forest is None
    assert_array_equal(back._flat.nodes, est._flat.nodes)
    est._device_forest = None
