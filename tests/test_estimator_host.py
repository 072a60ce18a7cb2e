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


def test_stream_predict_orders_results_and_rotates_replicas():
    # host logic of the streaming front end (device_forest.stream_predict) with stand-in replicas:
    # results come back in chunk order, chunk k goes to replica k mod G, an empty chunk yields an empty
    # result, the caller's validation runs in the workers and its error reaches the consumer
    import threading
    import time
    from skgarden_b200.device_forest import stream_predict

    class FakeReplica:
        n_features = 3

        def __init__(self, tag):
            self.tag, self.seen, self.lock = tag, [], threading.Lock()

        def predict_quantiles_host(self, X, q, out_host=None, mode=0):
            time.sleep(0.002 * (5 - int(X[0, 0]) % 5))          # later chunks finish earlier
            with self.lock:
                self.seen.append(int(X[0, 0]))
            return np.repeat(X[:, :1].astype(np.float64), len(q), axis=1) + self.tag

    reps = [FakeReplica(0.0), FakeReplica(0.5)]
    chunks = [np.full((k % 4 + 1, 3), k, dtype=np.float32) for k in range(9)]
    chunks.insert(4, np.empty((0, 3), dtype=np.float32))
    out = list(stream_predict(reps, iter(chunks), [10, 90], pinned=False,
                              prepare=lambda X: np.ascontiguousarray(X, dtype=np.float32)))
    assert [o.shape for o in out] == [(c.shape[0], 2) for c in chunks]
    nonempty = [c for c in chunks if c.shape[0]]
    got = [o for o in out if o.shape[0]]
    for k, (c, o) in enumerate(zip(nonempty, got)):
        assert np.all(o == c[0, 0] + (0.5 if k % 2 else 0.0))     # in order, replica k mod 2
    assert sorted(reps[0].seen) == [0, 2, 4, 6, 8] and sorted(reps[1].seen) == [1, 3, 5, 7]

    def bad(X):
        if X[0, 0] == 3:
            raise ValueError("Input X contains NaN.")
        return X
    with pytest.raises(ValueError, match="NaN"):
        list(stream_predict(reps, iter(chunks), [50], pinned=False, prepare=bad))
    with pytest.raises(ValueError, match="expects"):
        list(stream_predict(reps, [np.zeros((2, 4), dtype=np.float32)], [50], pinned=False))
