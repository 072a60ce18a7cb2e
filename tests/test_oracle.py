"""Pin the CPU oracle (oracle/qrf_oracle.py) to the reference:
  (a) the reference's own known-answer tests for weighted_percentile
      (skgarden/quantile/tests/test_utils.py:8-50),
  (b) committed outputs of the unmodified reference (tests/golden/*.npz),
  (c) when /root/reference is present: the live reference on random forests.
"""
import numpy as np
import pytest
from numpy.testing import assert_almost_equal, assert_array_almost_equal, assert_array_equal, assert_equal

import _golden
from oracle import qrf_oracle as O
from oracle import reference_shim


# ---- (a) reference known-answer tests, restated against the oracle --------------------
def test_percentile_equal_weights():            # test_utils.py:8-25
    rng = np.random.RandomState(0)
    x = rng.randn(10)
    weights = 0.1 * np.ones(10)
    sorted_x = np.sort(x)
    expected = 0.5 * (sorted_x[1:] + sorted_x[:-1])
    actual = [O.weighted_percentile(x, q, weights) for q in np.arange(10, 100, 10)]
    assert_array_almost_equal(expected, actual)
    actual = [O.weighted_percentile(x, q, weights) for q in np.arange(5, 105, 10)]
    assert_array_almost_equal(sorted_x, actual)


def test_percentile_toy_data():                 # test_utils.py:28-39 (golden vector)
    x, weights = [1, 2, 3], [1, 4, 5]
    assert_equal(O.weighted_percentile(x, 0, weights), 1)
    assert_equal(O.weighted_percentile(x, 100, weights), 3)
    assert_equal(O.weighted_percentile(x, 5, weights), 1)
    assert_equal(O.weighted_percentile(x, 30, weights), 2)
    assert_equal(O.weighted_percentile(x, 75, weights), 3)
    assert_almost_equal(O.weighted_percentile(x, 50, weights), 2.44, 2)
    assert O.weighted_percentile(x, 50, weights) == np.float32(2.4444444)


def test_zero_weights():                        # test_utils.py:42-50
    x, w = [1, 2, 3, 4, 5], [0, 0, 0, 0.1, 0.1]
    for q in np.arange(0, 110, 10):
        assert_equal(O.weighted_percentile(x, q, w), O.weighted_percentile([4, 5], q, [0.1, 0.1]))


def test_bad_quantile_raises():                 # utils.py:42-44
    with pytest.raises(ValueError):
        O.weighted_percentile([1, 2], 101, [1, 1])
    with pytest.raises(ValueError):
        O.weighted_percentile([1, 2], -1, [1, 1])


# ---- (b) golden fixtures ---------------------------------------------------------------
@pytest.mark.parametrize("name", _golden.CASES)
def test_oracle_matches_golden(name):
    g = _golden.load(name)
    assert_array_equal(O.forest_apply(g.estimators, g.X_test), g.apply)
    leaves, weights = O.fit_attributes(g.estimators, g.X_train, g.bootstrap)
    assert_array_equal(leaves, g.y_train_leaves_)
    assert_array_equal(weights, g.y_weights_)
    assert_array_equal(O.predict_mean(g.estimators, g.X_test), g.mean)
    index = O.SparseIndex(g.y_train, leaves, weights, g.sorter)
    sparse = O.predict_sparse(index, g.apply, g.q)
    assert_array_equal(sparse.T, g.pred)                      # bit-exact, all q at once
    rows = slice(0, min(25, len(g.X_test)))                   # dense form is O(T*N) per row
    for k in (0, 3, 6, 11, 13):
        q = float(g.q[k])
        dense = O.predict_dense(g.y_train, leaves, weights, g.apply[rows], q, g.sorter)
        assert_array_equal(dense, g.pred[k, rows])


# ---- (c) live reference ------------------------------------------------------------------
@pytest.mark.skipif(not reference_shim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("kind,bootstrap,msl,ties", [("rf", True, 1, True), ("et", True, 5, False),
                                                      ("rf", False, 3, True), ("et", False, 20, False)])
def test_oracle_matches_live_reference(kind, bootstrap, msl, ties):
    ref = reference_shim.load()
    cls = ref.RandomForestQuantileRegressor if kind == "rf" else ref.ExtraTreesQuantileRegressor
    rng = np.random.RandomState(11)
    N, n, F = 800, 120, 6
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    if ties:
        y = np.round(y, 1)
    est = cls(n_estimators=9, random_state=5, bootstrap=bootstrap, min_samples_leaf=msl,
              **ref.MODERN_KW).fit(X[:N], y[:N])
    leaves, weights = O.fit_attributes(est.estimators_, X[:N], bootstrap)
    assert_array_equal(leaves, est.y_train_leaves_)
    assert_array_equal(weights, est.y_weights_)
    app = O.forest_apply(est.estimators_, X[N:])
    assert_array_equal(app, est.apply(X[N:]))
    sorter = np.argsort(est.y_train_)
    index = O.SparseIndex(est.y_train_, leaves, weights, sorter)
    qs = [0, 5, 50, 97.5, 100]
    sparse = O.predict_sparse(index, app, qs)
    for k, q in enumerate(qs):
        assert_array_equal(sparse[:, k], est.predict(X[N:], quantile=q))
    assert_array_equal(O.predict_mean(est.estimators_, X[N:]), est.predict(X[N:]))
