"""Host logic: the flattener (csrc/pack_forest.cpp via the C ABI) reproduces the
reference's outputs when its layout is decoded on the CPU (tests/_layout.py)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

import _golden
import _layout
import skgarden_b200 as sg


@pytest.mark.parametrize("name", _golden.CASES)
def test_layout_reproduces_reference(name):
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    assert flat.n_trees == len(g.estimators)
    assert flat.n_nodes == g.node_offsets[-1] + flat.n_trees          # one padding slot per tree
    node_idx, start, count = _layout.decode_apply(flat, g.X_test)
    assert_array_equal(_layout.sklearn_ids(flat, node_idx), g.apply)        # leaf ids bit-exact
    assert count.min() >= 1                                                 # every leaf has members
    assert count.sum(axis=1).max() <= flat.max_row_members
    pred = _layout.decode_predict(flat, start, count, g.q)
    assert_array_equal(pred.T, g.pred)                                      # bit-exact
    # mean path data
    got = flat.leaf_values[node_idx].sum(axis=1)                            # order-insensitive check
    np.testing.assert_allclose(got / flat.n_trees, g.mean, rtol=1e-12)


def test_level_order_layout_with_sibling_pairs():
    """include/qforest_b200.h: slot 0 root, slot 1 padding, children side by side and later than
    their parent, every level contiguous; leaves' member ranges grow from left to right."""
    for name in ("rf_maxleaf50", "et_msl5"):           # best-first and depth-first sklearn numbering
        g = _golden.load(name)
        flat = _golden.flat_from_golden(g)
        assert flat.node_ids is not None
        lo = (flat.nodes & np.uint64(0xFFFFFFFF)).astype(np.int64)
        hi = (flat.nodes >> np.uint64(32)).astype(np.uint32)
        for t, est in enumerate(g.estimators):
            tr = est.tree_
            n0, n1 = int(flat.tree_offsets[t]), int(flat.tree_offsets[t + 1])
            assert n1 - n0 == tr.node_count + 1 and n0 % 2 == 0
            ids = flat.node_ids[n0:n1]
            assert ids[0] == 0 and ids[1] == -1 and flat.nodes[n0 + 1] == 0
            assert sorted(ids[ids >= 0]) == list(range(tr.node_count))
            pos = np.empty(tr.node_count, dtype=np.int64)
            pos[ids[ids >= 0]] = np.flatnonzero(ids >= 0)
            inner = np.flatnonzero(tr.children_left != -1)
            off = (hi[n0 + pos[inner]] >> np.uint32(flat.feature_bits)).astype(np.int64)
            assert_array_equal(pos[tr.children_left[inner]], pos[inner] + off)
            assert_array_equal(pos[tr.children_right[inner]], pos[inner] + off + 1)
            assert np.all(pos[tr.children_left[inner]] % 2 == 0)
            # level order: depth never decreases along the slots
            depth = np.zeros(tr.node_count, dtype=np.int64)
            for nd in np.argsort(pos):                   # parents come before their children
                if tr.children_left[nd] != -1:
                    depth[tr.children_left[nd]] = depth[tr.children_right[nd]] = depth[nd] + 1
            assert np.all(np.diff(depth[np.argsort(pos)]) >= 0)
            # member ranges follow the leaves from left to right (depth-first order)
            order, stack = [], [0]
            while stack:
                nd = stack.pop()
                if tr.children_left[nd] == -1:
                    order.append(nd)
                else:
                    stack += [tr.children_right[nd], tr.children_left[nd]]
            starts = lo[n0 + pos[order]]
            cnts = (hi[n0 + pos[order]] & np.uint32(0x7FFFFFFF)).astype(np.int64)
            assert starts[0] == flat.member_offsets[t]
            assert_array_equal(starts[1:], starts[:-1] + cnts[:-1])


def test_members_sorted_by_rank_within_leaf():
    g = _golden.load("et_msl5")
    flat = _golden.flat_from_golden(g)
    hi = (flat.nodes >> np.uint64(32)).astype(np.uint32)
    lo = (flat.nodes & np.uint64(0xFFFFFFFF)).astype(np.int64)
    leaf = (hi & np.uint32(0x80000000)) != 0
    rank = (flat.members & np.uint64(0xFFFFFFFF)).astype(np.int64)
    for s, c in zip(lo[leaf][:500], (hi[leaf] & np.uint32(0x7FFFFFFF))[:500]):
        seg = rank[s:s + int(c)]
        assert np.all(np.diff(seg) > 0)


def test_floor_threshold_is_exact_on_grid_valued_features():
    # thresholds are midpoints of float32 feature values; rounding them to the NEAREST
    # float32 flips `<=` decisions (SURVEY.md Appendix C probe 12); floor does not.
    from sklearn.ensemble import RandomForestRegressor
    rng = np.random.RandomState(0)
    X = (rng.randint(0, 4000, size=(3000, 5)) / np.float32(7.0)).astype(np.float32)
    y = X[:, 0] - X[:, 1] + rng.randn(3000)
    rf = RandomForestRegressor(n_estimators=5, random_state=0, bootstrap=False).fit(X, y)
    trees = [e.tree_ for e in rf.estimators_]
    leaves = [t.apply(X) for t in trees]
    flat = sg.flatten_forest(trees, leaves, None, y, 5)
    Xt = (rng.randint(0, 4000, size=(4000, 5)) / np.float32(7.0)).astype(np.float32)
    node_idx, _, _ = _layout.decode_apply(flat, Xt)
    assert_array_equal(_layout.sklearn_ids(flat, node_idx), rf.apply(Xt))


def test_flat_forest_roundtrip(tmp_path):
    g = _golden.load("c1_rfqr_boston")
    flat = _golden.flat_from_golden(g)
    p = str(tmp_path / "forest.npz")
    flat.save(p)
    back = sg.FlatForest.load(p)
    assert_array_equal(back.nodes, flat.nodes)
    assert_array_equal(back.members, flat.members)
    assert back.max_row_members == flat.max_row_members
    assert_array_equal(back.node_ids, flat.node_ids)


def test_pack_tree_rejects_bad_input():
    from skgarden_b200 import _native
    g = _golden.load("c1_rfqr_boston")
    trees = [e.tree_ for e in g.estimators]
    with pytest.raises(ValueError):
        sg.flatten_forest(trees, [np.zeros(3, dtype=np.int64)] * len(trees), None, g.y_train, 13)
    bad = [np.zeros(len(g.y_train), dtype=np.int64)] * len(trees)     # node 0 is not a leaf
    with pytest.raises(_native.QFError):
        sg.flatten_forest(trees, bad, None, g.y_train, 13)


@pytest.mark.parametrize("block_levels", ["0", "1", "2", "5"])
def test_packer_on_odd_tree_shapes(block_levels, monkeypatch):
    """Single-leaf trees, stumps, chains, best-first numbering, with plain level order and the
    blocked variants of the layout (QF_NODE_BLOCK_LEVELS): leaf ids and member lists decode to
    scikit-learn's apply and to the oracle's percentiles."""
    from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
    from oracle import qrf_oracle as O
    monkeypatch.setenv("QF_NODE_BLOCK_LEVELS", block_levels)
    rng = np.random.RandomState(int(block_levels))
    X = rng.randn(400, 4).astype(np.float32)
    Xt = rng.randn(64, 4).astype(np.float32)
    chain_y = np.cumsum(np.abs(rng.randn(400))) * (X[:, 0] > -10)       # deep one-sided splits
    cases = [
        (RandomForestRegressor(n_estimators=3, random_state=0), np.zeros(400)),               # root is a leaf
        (RandomForestRegressor(n_estimators=3, max_depth=1, random_state=0), X[:, 0] + 0.0),   # stumps
        (ExtraTreesRegressor(n_estimators=3, max_leaf_nodes=7, random_state=0), X[:, 1] * X[:, 0]),
        (RandomForestRegressor(n_estimators=2, min_samples_leaf=1, random_state=1), chain_y),
        (ExtraTreesRegressor(n_estimators=4, min_samples_leaf=3, bootstrap=True, random_state=2),
         np.round(X[:, 0] - X[:, 2], 1)),
    ]
    for est, y in cases:
        est.fit(X, y)
        trees = [e.tree_ for e in est.estimators_]
        boot = bool(est.bootstrap)
        counts = [sg.bootstrap_counts(e.random_state, 400) for e in est.estimators_] if boot else None
        flat = sg.flatten_forest(trees, [t.apply(X) for t in trees], counts, y, 4)
        assert flat.n_nodes == sum(t.node_count + 1 for t in trees)
        node_idx, start, count = _layout.decode_apply(flat, Xt)
        leaves = est.apply(Xt)
        assert_array_equal(_layout.sklearn_ids(flat, node_idx), leaves)
        index = O.SparseIndex(y, *O.fit_attributes(est.estimators_, X, boot), flat.sorter)
        want = O.predict_sparse(index, leaves, [10, 50, 90])
        assert_array_equal(_layout.decode_predict(flat, start, count, [10, 50, 90]), want)


def test_load_refuses_another_packed_layout(tmp_path):
    g = _golden.load("c1_rfqr_boston")
    flat = _golden.flat_from_golden(g)
    p = str(tmp_path / "forest.npz")
    flat.save(p)
    z = dict(np.load(p))
    z.pop("layout")                                   # a file written before the layout was versioned
    old = str(tmp_path / "old.npz")
    np.savez(old, **z)
    with pytest.raises(ValueError, match="layout"):
        sg.FlatForest.load(old)
