"""K4, the device CSR builder (qf_build_csr): must reproduce the host flattener bit for bit
(which tests/test_flatten.py pins to the reference's fit attributes, ensemble.py:83-101)."""
import ctypes as C

import numpy as np
import pytest
from numpy.testing import assert_array_equal


def _forest(kind, **kw):
    import skgarden_b200 as sg
    rng = np.random.RandomState(4)
    N, F = 3000, 7
    X = rng.randn(N, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N)
    if kw.pop("ties", False):
        y = np.round(y, 1)
    cls = sg.RandomForestQuantileRegressor if kind == "rf" else sg.ExtraTreesQuantileRegressor
    return cls(**{**dict(n_estimators=9, random_state=1, n_jobs=-1), **kw}), X, y


def test_structure_only_packing_matches_full_packing():
    # qf_pack_tree with train_leaves = NULL: same internal nodes and node ids, empty leaves
    from skgarden_b200 import _native
    from skgarden_b200.flatten import _ptr, _tree_arrays
    est, X, y = _forest("rf", max_leaf_nodes=40)           # best-first builder: renumbered nodes
    est.fit(X, y)
    fl = est._flat
    L = _native.lib()
    for t, e in enumerate(est.estimators_):
        left, right, feat, thr, _ = _tree_arrays(e.tree_)
        n = len(left)
        nodes, ids, depth = np.empty(n + 1, np.uint64), np.empty(n + 1, np.int32), C.c_int32(0)
        base = int(fl.tree_offsets[t]) // 2                # first leaf ordinal of the tree
        rc = L.qf_pack_tree(n, _ptr(left), _ptr(right), _ptr(feat), _ptr(thr), fl.feature_bits, 0, None, None,
                            None, base, _ptr(nodes), _ptr(ids), None, None, None, C.byref(depth), None)
        assert rc == 0, _native.last_error()
        ref = fl.nodes[fl.tree_offsets[t]:fl.tree_offsets[t + 1]]
        leaf = (ref >> np.uint64(63)) == 1
        assert_array_equal(nodes[~leaf], ref[~leaf])
        # empty leaves carrying their left-to-right ordinal: same order as the member ranges
        assert np.all(nodes[leaf] >> np.uint64(32) == np.uint64(0x80000000))
        ordinal = (nodes[leaf] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert sorted(ordinal) == list(range(base, base + (n + 1) // 2))
        start = (ref[leaf] & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.all(np.diff(start[np.argsort(ordinal)]) >= 0)
        assert_array_equal(ids, fl.node_ids[fl.tree_offsets[t]:fl.tree_offsets[t + 1]])
        assert depth.value == e.tree_.max_depth


@pytest.mark.gpu
@pytest.mark.parametrize("kind,kw", [
    ("rf", dict()), ("et", dict(min_samples_leaf=5)), ("rf", dict(bootstrap=False, min_samples_leaf=3, ties=True)),
    ("rf", dict(max_leaf_nodes=60)), ("et", dict(max_depth=4)), ("et", dict(min_samples_leaf=40, bootstrap=False)),
    ("rf", dict(max_depth=3, min_samples_leaf=30, n_estimators=100))])        # every leaf is a long list (found by tools/fuzz_parity.py)
def test_device_builder_equals_host_flattener(kind, kw):
    pytest.importorskip("torch")
    est, X, y = _forest(kind, **kw)
    est.fit(X, y)
    host = est._flat
    est.set_params(builder="device")
    dev = est._flatten(X, y)
    assert_array_equal(dev.nodes, host.nodes)
    assert_array_equal(dev.members, host.members)             # ranks and float32 weight bits
    assert_array_equal(dev.member_offsets, host.member_offsets)
    assert_array_equal(dev.leaf_values, host.leaf_values)
    assert_array_equal(dev.node_ids, host.node_ids)
    assert dev.max_row_members == host.max_row_members and dev.max_depth == host.max_depth
    assert dev.expected_row_members == host.expected_row_members
    # and the estimator predicts the same through either builder
    want = est.__class__(**{**est.get_params(), "builder": "host"}).fit(X, y).predict(X[:200], quantile=[5, 50, 95])
    got = est.fit(X, y).predict(X[:200], quantile=[5, 50, 95])
    assert_array_equal(got, want)


@pytest.mark.gpu
def test_device_builder_refuses_very_long_leaf_lists():
    from skgarden_b200 import _native
    est, X, y = _forest("et", max_depth=1, bootstrap=False)
    est.set_params(builder="device")
    with pytest.raises(_native.QFError, match="4096"):
        est.fit(np.tile(X, (4, 1)), np.tile(y, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1_rfqr_boston", "c1_etqr_boston", "rf_noboot_msl3_ties", "et_msl5", "rf_maxleaf50",
                                  "et_depth2_large_sets", "rf_pure_leaves"])
def test_device_builder_reproduces_the_reference_fit_attributes(name):
    # K4 against the REFERENCE's own outputs (tests/golden, generated from the unmodified reference):
    # y_train_leaves_ / y_weights_ of ensemble.py:83-101 and the predictions, not just the host flattener
    pytest.importorskip("torch")
    import _golden
    import skgarden_b200 as sg
    from skgarden_b200.flatten import flatten_forest_device
    g = _golden.load(name)
    N = len(g.y_train)
    trees = [e.tree_ for e in g.estimators]
    counts = [sg.bootstrap_counts(e.random_state, N) for e in g.estimators] if g.bootstrap else None
    hi_leaf = max(int(np.bincount(t.apply(g.X_train)).max()) for t in trees)
    if hi_leaf > 4096:
        pytest.skip("a leaf of %d members: the device builder hands such forests to the host flattener" % hi_leaf)
    flat = flatten_forest_device(trees, g.X_train, counts, g.y_train, g.X_train.shape[1], sorter=g.sorter)
    shell = sg.RandomForestQuantileRegressor()
    shell._set_flat(flat)
    assert_array_equal(shell.y_train_leaves_, g.y_train_leaves_)
    assert_array_equal(shell.y_weights_, g.y_weights_)
    dev = sg.DeviceForest(flat, device=0)
    import torch
    X = torch.from_numpy(np.ascontiguousarray(g.X_test)).cuda()
    assert_array_equal(dev.apply(X).cpu().numpy(), g.apply)
    assert_array_equal(dev.predict_quantiles(X, g.q).cpu().numpy().T, g.pred)
    dev.close()
