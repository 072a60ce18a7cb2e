"""Tree-level quantile regressors (skgarden/quantile/tree.py): oracle pinned to the
reference's own outputs (CPU), device path against them (GPU, through the C ABI)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal, assert_array_equal

import _golden_tree
from oracle import qrf_oracle as O


# ---- CPU: the oracle against the unmodified reference's outputs --------------------------------
@pytest.mark.parametrize("name", _golden_tree.CASES)
def test_oracle_tree_predict_matches_reference(name):
    g = _golden_tree.load(name)
    leaves = O.tree_apply(g.tree, g.X_test)
    assert_array_equal(leaves, g.apply)
    assert_array_equal(O.tree_apply(g.tree, g.X_train), g.y_train_leaves_)
    for k, q in enumerate(g.q):
        q = int(q) if float(q).is_integer() else float(q)
        assert_array_equal(O.tree_predict_quantile(g.y_train, g.y_train_leaves_, leaves, q), g.pred[k])
    assert_array_equal(np.asarray(g.tree.value)[leaves, 0, 0], g.mean)


@pytest.mark.parametrize("name", _golden_tree.CASES)
def test_unit_weight_flattening_equals_tree_predict(name):
    # the forest arithmetic with T = 1 and weights 1.0 IS the tree-level arithmetic
    g = _golden_tree.load(name)
    N = len(g.y_train)
    index = O.SparseIndex(g.y_train, g.y_train_leaves_[None, :].astype(np.int32),
                          np.ones((1, N), dtype=np.float32))
    got = O.predict_sparse(index, g.apply[:, None], g.q)
    assert_array_equal(got.T, g.pred)
    flat = _golden_tree.flat_from_golden(g)
    w = (flat.members >> np.uint64(32)).astype(np.uint32).view(np.float32)
    assert flat.n_trees == 1 and flat.n_members == N and np.all(w == 1.0)


def test_estimator_signature_and_fit_attributes():
    import skgarden_b200 as sg
    rng = np.random.RandomState(0)
    X = rng.randn(300, 5).astype(np.float32)
    y = X[:, 0] + rng.randn(300)
    for cls, splitter, mf in ((sg.DecisionTreeQuantileRegressor, "best", None),
                              (sg.ExtraTreeQuantileRegressor, "random", "auto")):
        est = cls(random_state=0)
        p = est.get_params()
        assert p["splitter"] == splitter and p["max_features"] == mf and p["criterion"] == "mse"
        est.set_params(max_depth=3).fit(X, y)
        assert_array_equal(est.y_train_, y)
        assert_array_equal(est.y_train_leaves_, est.tree_.apply(X))          # tree.py:100
        assert est._flat.n_trees == 1 and est._flat.n_members == 300
    with pytest.raises(Exception):
        sg.DecisionTreeQuantileRegressor().predict(X, quantile=50)           # not fitted


# ---- GPU ----------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", _golden_tree.CASES)
def test_device_tree_predict_matches_reference(name):
    torch = pytest.importorskip("torch")
    import skgarden_b200 as sg
    g = _golden_tree.load(name)
    dev = sg.DeviceForest(_golden_tree.flat_from_golden(g), device=0)
    X = torch.from_numpy(np.ascontiguousarray(g.X_test)).cuda()
    assert_array_equal(dev.apply(X).cpu().numpy()[:, 0], g.apply)            # leaf ids bit-exact
    assert_array_equal(dev.predict_quantiles(X, g.q).cpu().numpy().T, g.pred)
    assert_array_equal(dev.predict_mean(X).cpu().numpy(), g.mean)
    # fast mode on unit weights: the total weight of a row is its leaf size (up to ~150 here),
    # the library scales its fixed-point weights by that bound
    from skgarden_b200 import _native
    dev.set_option("warp_elems", 0)
    scale = np.abs(g.y_train).max()
    for lo in range(0, len(g.q), 7):
        fast = dev.predict_quantiles(X, g.q[lo:lo + 7], mode=_native.QF_MODE_FAST).cpu().numpy().T
        assert_allclose(fast, g.pred[lo:lo + 7], rtol=1e-4, atol=1e-4 * scale)
    dev.close()


@pytest.mark.gpu
def test_reference_tree_tests_through_the_drop_in_api():
    # skgarden/quantile/tests/test_tree.py:22-84 on synthetic data
    import skgarden_b200 as sg
    rng = np.random.RandomState(0)
    X = rng.randn(400, 6).astype(np.float32)
    y = 3 * X[:, 0] + rng.randn(400)
    Xtr, Xte, ytr = X[:250], X[250:], y[:250]
    for cls in (sg.DecisionTreeQuantileRegressor, sg.ExtraTreeQuantileRegressor):
        est = cls(random_state=0, max_depth=1).fit(Xtr, ytr)                 # test_quantiles
        tree = est.tree_
        for q in [20, 40, 50, 60, 80, 90]:
            left = Xtr[:, tree.feature[0]] <= tree.threshold[0]
            lq, rq = O.weighted_percentile(ytr[left], q), O.weighted_percentile(ytr[~left], q)
            for cur in (Xtr, Xte):
                go_left = cur[:, tree.feature[0]] <= tree.threshold[0]
                assert_array_equal(est.predict(cur, quantile=q), np.where(go_left, lq, rq))
        est = cls(random_state=0, max_depth=None).fit(Xtr, ytr)              # test_max_depth_None
        for q in [20, 40, 50, 60, 80, 90]:
            for cur in (Xtr, Xte):
                assert_array_almost_equal(est.predict(cur), est.predict(cur, quantile=q), 1)
        with pytest.raises(ValueError):
            est.predict(Xte, quantile=101)
        exact = est.predict(Xte, quantile=[10, 50, 90])
        est.mode = "fast"               # unit weights: a leaf's weights sum to its size, not to one
        est._on_device().set_option("warp_elems", 0)            # selection kernel for every row
        assert_array_almost_equal(est.predict(Xte, quantile=[10, 50, 90]), exact, 5)
        est.mode = "exact"
    # test_tree_toy_data: two clusters, depth 1 -> np.percentile of each cluster (3 decimals)
    x1 = rng.randn(1, 10)
    x2 = 20.0 * rng.randn(1, 10)
    Xc = np.vstack((np.tile(x1, (10000, 1)), np.tile(x2, (10000, 1)))).astype(np.float32)
    y1, y2 = rng.randn(10000), 5.0 + rng.randn(10000)
    for cls in (sg.DecisionTreeQuantileRegressor, sg.ExtraTreeQuantileRegressor):
        est = cls(random_state=0, max_depth=1).fit(Xc, np.concatenate((y1, y2)))
        for q in [20, 30, 40, 50, 60, 70, 80]:
            assert_array_almost_equal(est.predict(x1.astype(np.float32), quantile=q), [np.percentile(y1, q)], 3)
            assert_array_almost_equal(est.predict(x2.astype(np.float32), quantile=q), [np.percentile(y2, q)], 3)
