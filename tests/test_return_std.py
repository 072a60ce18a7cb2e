"""``return_std`` forests (skgarden/forest.py:133-178, :332-362, :516-548): oracle pinned to the
reference's own outputs (CPU), device path against them (GPU, through the C ABI)."""
import glob
import os
import types

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from _golden import ArrayTree
from oracle import qrf_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_std")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = types.SimpleNamespace(**{k: z[k] for k in z.files})
    off = z["node_offsets"]
    g.trees = []
    for t in range(len(off) - 1):
        s = slice(off[t], off[t + 1])
        tree = ArrayTree(z["children_left"][s], z["children_right"][s], z["feature"][s], z["threshold"][s],
                         z["value"][s])
        tree.impurity = z["impurity"][s]
        g.trees.append(tree)
    return g


def _flat(g):
    import skgarden_b200 as sg
    leaves = [t.apply(g.X_train) for t in g.trees]
    return sg.flatten_forest(g.trees, leaves, None, g.y_train, g.X_train.shape[1])


@pytest.mark.parametrize("name", CASES)
def test_oracle_return_std_matches_reference(name):
    g = _load(name)
    leaves = O.forest_apply(g.trees, g.X_test)
    assert_array_equal(leaves, g.apply)
    mean = O.predict_mean(g.trees, g.X_test, leaves)
    for k, mv in enumerate(g.min_variance):
        # the reference's mean is sklearn's threaded accumulation: same values, bit-exact here
        assert_array_equal(mean, g.mean[k])
        assert_array_equal(O.forest_return_std(g.trees, leaves, mean, mv), g.std[k])


def test_impurities_travel_with_the_flattened_forest(tmp_path):
    import skgarden_b200 as sg
    g = _load("rf_msl10")
    flat = _flat(g)
    assert flat.leaf_impurity is not None and flat.leaf_impurity.shape == flat.leaf_values.shape
    for t, tree in enumerate(g.trees):
        n0, n1 = flat.tree_offsets[t], flat.tree_offsets[t + 1]
        ids = flat.node_ids[n0:n1]
        assert_array_equal(flat.leaf_impurity[n0:n1][ids >= 0], tree.impurity[ids[ids >= 0]])
    path = str(tmp_path / "f.npz")
    flat.save(path)
    assert_array_equal(sg.FlatForest.load(path).leaf_impurity, flat.leaf_impurity)
    est = sg.ExtraTreesRegressor()
    assert est.get_params()["bootstrap"] is False and est.get_params()["min_variance"] == 0.0
    assert sg.RandomForestRegressor().get_params()["bootstrap"] is True


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_return_std_matches_reference(name):
    torch = pytest.importorskip("torch")
    import skgarden_b200 as sg
    g = _load(name)
    dev = sg.DeviceForest(_flat(g), device=0, with_impurity=True)
    X = torch.from_numpy(np.ascontiguousarray(g.X_test)).cuda()
    for k, mv in enumerate(g.min_variance):
        mean, std = dev.predict_mean_std(X, float(mv))
        assert_array_equal(mean.cpu().numpy(), g.mean[k])
        assert_array_equal(std.cpu().numpy(), g.std[k])                       # float64, bit-exact
    dev.close()


@pytest.mark.gpu
def test_return_std_estimators():
    # skgarden/tests/test_forest.py:12-68 through the drop-in API
    import skgarden_b200 as sg
    rng = np.random.RandomState(1)
    X = rng.randn(500, 5).astype(np.float32)
    y = 2 * X[:, 0] + rng.randn(500)
    for cls in (sg.RandomForestRegressor, sg.ExtraTreesRegressor):
        est = cls(n_estimators=8, random_state=0, min_samples_leaf=5, n_jobs=1).fit(X[:400], y[:400])
        mean = est.predict(X[400:])
        mean2, std = est.predict(X[400:], return_std=True)
        assert_array_equal(mean, mean2)
        leaves = est.apply(X[400:])
        assert_array_equal(mean, O.predict_mean(est.estimators_, X[400:], leaves))
        assert_array_equal(std, O.forest_return_std(est.estimators_, leaves, mean, 0.0))
        assert np.all(std >= 0)
        est.set_params(min_variance=10.0)                       # clamps every leaf variance from below
        _, std_hi = est.predict(X[400:], return_std=True)
        assert np.all(std_hi >= np.sqrt(10.0) - 1e-9) and np.all(std_hi >= std)
        assert_array_equal(std_hi, O.forest_return_std(est.estimators_, leaves, mean, 10.0))
        est.set_params(criterion="mae")
        with pytest.raises(ValueError, match="mse"):
            est.predict(X[400:], return_std=True)
        # pure leaves (msl=1, no bootstrap): zero variance inside a leaf -> std is the spread of the trees
    est = sg.ExtraTreesRegressor(n_estimators=5, random_state=0).fit(X[:400], y[:400])
    m, s = est.predict(X[:400], return_std=True)
    assert_allclose(m, y[:400], rtol=0, atol=1e-9)
    assert_allclose(s, 0.0, atol=1e-6)
