"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every call goes through
the C ABI (libqforest_b200.so).  Oracles: committed outputs of the unmodified reference
(tests/golden) and oracle/qrf_oracle.py.  Bars: leaf ids and QF_MODE_EXACT percentiles
bit-exact; means bit-exact (float64 tree-order sum); QF_MODE_FAST within 1e-4."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_almost_equal, assert_array_equal

import _golden
from oracle import qrf_oracle as O

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _dev(flat, **opts):
    import skgarden_b200 as sg
    d = sg.DeviceForest(flat, device=0)
    for k, v in opts.items():
        d.set_option(k, v)
    return d


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_native_library_is_loaded():
    from skgarden_b200 import _native
    assert _native.lib().qf_abi_version() == _native.ABI_VERSION == 2
    with open("/proc/self/maps") as fh:
        assert "libqforest_b200.so" in fh.read()


# ---- golden fixtures: the reference's own outputs ------------------------------------------------
@pytest.mark.parametrize("name", _golden.CASES)
def test_golden_apply_quantiles_mean(name):
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat)
    X = _cuda(g.X_test)
    assert_array_equal(dev.apply(X).cpu().numpy(), g.apply)                  # leaf ids bit-exact
    got = dev.predict_quantiles(X, g.q).cpu().numpy()
    assert_array_equal(got.T, g.pred)                                        # bit-exact, all q
    assert_array_equal(dev.predict_mean(X).cpu().numpy(), g.mean)            # float64, tree order
    host = dev.predict_quantiles_host(g.X_test, g.q)
    assert_array_equal(host, got)                                            # host-buffer path
    dev.close()


@pytest.mark.parametrize("name", _golden.CASES)
@pytest.mark.parametrize("warp_elems", [0, 2, 8])
def test_every_quantile_kernel_is_identical(name, warp_elems):
    # warp kernel / CTA kernel as the primary, each with its overflow chain
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, warp_elems=warp_elems)
    got = dev.predict_quantiles(_cuda(g.X_test), g.q).cpu().numpy()
    assert_array_equal(got.T, g.pred)
    dev.close()


@pytest.mark.parametrize("name", ["et_msl5", "et_depth2_large_sets", "c1_rfqr_boston"])
@pytest.mark.parametrize("cta2_capacity", [0, 40])
def test_dense_fallback_is_identical(name, cta2_capacity):
    # force rows out of shared memory into the global-scratch path: directly from the first
    # bucket stage (cta2_capacity=0) and through the whole chain warp -> bucket -> bucket -> dense
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, cta_capacity=8, cta2_capacity=cta2_capacity, dense_ctas=3,
               warp_elems=0 if cta2_capacity == 0 else 2)
    got = dev.predict_quantiles(_cuda(g.X_test), g.q).cpu().numpy()
    assert_array_equal(got.T, g.pred)
    dev.close()


@pytest.mark.parametrize("opts", [
    dict(apply_row_warps=1, apply_ilp=1), dict(apply_row_warps=2, apply_ilp=2),
    dict(apply_row_warps=8, apply_ilp=8), dict(rows_per_chunk=37),
    dict(apply_row_warps=4, apply_ilp=4),
    dict(sort_rows=0), dict(sort_rows=1), dict(sort_rows=1, rows_per_chunk=100, apply_row_warps=4),
    dict(sort_rows=1, warp_elems=0), dict(sort_rows=1, warp_elems=2, cta_capacity=64, dense_ctas=2),
    dict(warp_elems=0), dict(warp_elems=2), dict(warp_elems=4), dict(warp_elems=8),
    dict(warp_elems=2, cta_capacity=64), dict(warp_elems=4, cta_capacity=128, dense_ctas=2),
    dict(warp_elems=0, cta_capacity=300, cta2_capacity=1000), dict(warp_elems=0, cta2_capacity=0)])
def test_launch_shapes_do_not_change_results(opts):
    g = _golden.load("et_msl5")
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, **opts)
    X = _cuda(g.X_test)
    assert_array_equal(dev.apply(X).cpu().numpy(), g.apply)
    assert_array_equal(dev.predict_mean(X).cpu().numpy(), g.mean)
    assert_array_equal(dev.predict_quantiles(X, g.q).cpu().numpy().T, g.pred)
    assert_array_equal(dev.predict_quantiles_host(g.X_test, g.q, chunk_rows=41).T, g.pred)
    dev.close()


def test_many_quantiles_and_ragged_rows():
    g = _golden.load("c1_etqr_boston")
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat)
    qs = np.linspace(0, 100, 301)                              # > 128 per launch -> several launches
    leaves, weights = O.fit_attributes(g.estimators, g.X_train, g.bootstrap)
    index = O.SparseIndex(g.y_train, leaves, weights, g.sorter)
    for n in (1, 31, 33):
        got = dev.predict_quantiles(_cuda(g.X_test[:n]), qs).cpu().numpy()
        assert_array_equal(got, O.predict_sparse(index, g.apply[:n], qs))
    dev.close()


def test_error_paths():
    from skgarden_b200 import _native
    g = _golden.load("c1_rfqr_boston")
    dev = _dev(_golden.flat_from_golden(g))
    X = _cuda(g.X_test)
    with pytest.raises(_native.QFError, match="in-between 0 and 100"):
        dev.predict_quantiles(X, [50, 100.5])
    with pytest.raises(ValueError):
        dev.predict_quantiles(_cuda(g.X_test[:, :4]), [50])
    with pytest.raises(TypeError):
        dev.predict_quantiles(_cuda(g.X_test.astype(np.float64)), [50])
    with pytest.raises(_native.QFError):
        dev.set_option("no_such_option", 1)
    dev.close()


# ---- the reference's estimator-level tests, through the drop-in API ------------------------------
def test_estimator_matches_reference_outputs():
    import skgarden_b200 as sg
    for name, cls in (("c1_rfqr_boston", sg.RandomForestQuantileRegressor),
                      ("c1_etqr_boston", sg.ExtraTreesQuantileRegressor)):
        g = _golden.load(name)
        est = cls(n_estimators=10, random_state=0).fit(g.X_train, g.y_train)
        if not np.array_equal(np.concatenate([e.tree_.threshold for e in est.estimators_]), g.threshold):
            pytest.skip("scikit-learn on this box grows different trees than the fixture's")
        assert_array_equal(est.apply(g.X_test), g.apply)
        for k in (0, 4, 6, 11, 13):
            q = g.q[k]
            assert_array_equal(est.predict(g.X_test, quantile=int(q) if float(q).is_integer() else q), g.pred[k])
        assert_array_equal(est.predict(g.X_test), g.mean)
        assert_array_equal(est.predict(g.X_test, quantile=[5, 50, 95]), g.pred[[2, 6, 10]].T)
        assert est.predict(g.X_test, quantile=50).dtype == np.float64
        assert est.score(g.X_test, g.mean) == 1.0


def test_max_depth_None_rfqr():                                 # test_ensemble.py:69-83
    import skgarden_b200 as sg
    rng = np.random.RandomState(0)
    X = rng.randn(10, 1)
    y = np.linspace(0.0, 100.0, 10)
    rfqr = sg.RandomForestQuantileRegressor(random_state=0, bootstrap=False, max_depth=None).fit(X, y)
    for quantile in [20, 40, 50, 60, 80, 90]:
        assert_array_almost_equal(rfqr.predict(X, quantile=None), rfqr.predict(X, quantile=quantile), 5)


def test_forest_toy_data():                                     # test_ensemble.py:105-126
    import skgarden_b200 as sg
    rng = np.random.RandomState(1)
    x1 = rng.randn(1, 10)
    X1 = np.tile(x1, (10000, 1))
    x2 = 20.0 * rng.randn(1, 10)
    X2 = np.tile(x2, (10000, 1))
    X = np.vstack((X1, X2))
    y1 = rng.randn(10000)
    y2 = 5.0 + rng.randn(10000)
    y = np.concatenate((y1, y2))
    # bootstrap=False as in the reference run order (test_ensemble.py:42 leaves it off)
    for est in (sg.RandomForestQuantileRegressor(random_state=0, bootstrap=False, max_depth=1),
                sg.ExtraTreesQuantileRegressor(random_state=0, bootstrap=False, max_depth=1)):
        est.fit(X, y)
        for quantile in [20, 30, 40, 50, 60, 70, 80]:
            assert_array_almost_equal(est.predict(x1, quantile=quantile), [np.percentile(y1, quantile)], 3)
            assert_array_almost_equal(est.predict(x2, quantile=quantile), [np.percentile(y2, quantile)], 3)


def test_tree_forest_equivalence():                             # test_ensemble.py:51-66
    # forest(bootstrap=False, max_depth=2) == single tree: every tree of the forest sees
    # the same data, so the forest's percentile equals the one-tree percentile.
    import skgarden_b200 as sg
    g = _golden.load("c1_rfqr_boston")
    f10 = sg.RandomForestQuantileRegressor(n_estimators=10, random_state=0, bootstrap=False, max_depth=2)
    f1 = sg.RandomForestQuantileRegressor(n_estimators=1, random_state=0, bootstrap=False, max_depth=2)
    f10.fit(g.X_train, g.y_train)
    f1.fit(g.X_train, g.y_train)
    assert_array_almost_equal(f10.predict(g.X_test, quantile=10), f1.predict(g.X_test, quantile=10), 5)


# ---- randomized differential tests against the oracle at sizes it finishes in seconds -------------
@pytest.mark.parametrize("kind,bootstrap,msl,ties,depth", [
    ("rf", True, 1, True, None), ("et", True, 5, False, None), ("rf", False, 2, False, None),
    ("et", True, 40, True, None), ("rf", True, 1, False, 6), ("rf", True, 1, True, 4)])
def test_random_forests_bit_exact_vs_oracle(kind, bootstrap, msl, ties, depth):
    import skgarden_b200 as sg
    cls = sg.RandomForestQuantileRegressor if kind == "rf" else sg.ExtraTreesQuantileRegressor
    rng = np.random.RandomState(7)
    N, n, F = 6000, 3000, 9
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    if ties:
        y = np.round(y, 1)
    est = cls(n_estimators=30, random_state=3, bootstrap=bootstrap, min_samples_leaf=msl,
              max_depth=depth, n_jobs=-1).fit(X[:N], y[:N])
    Xt = X[N:]
    leaves = est.apply(Xt)
    assert_array_equal(leaves, O.forest_apply(est.estimators_, Xt))          # all rows, bit-exact
    qs = [0, 2.5, 5, 50, 95, 97.5, 100]
    got = est.predict(Xt, quantile=qs)
    rows = rng.choice(n, 400, replace=False)
    index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
    assert_array_equal(got[rows], O.predict_sparse(index, leaves[rows], qs))
    assert_array_equal(est.predict(Xt), O.predict_mean(est.estimators_, Xt, leaves))
    # size-independent properties on ALL rows: monotone in q (up to the float32 rounding
    # of the reference's interpolation), inside the training range, q=0 / q=100 are order
    # statistics (exact members of y_train)
    assert np.all(np.diff(got, axis=1) >= -1e-5 * np.abs(y).max())
    y32 = est.y_train_.astype(np.float32).astype(np.float64)
    assert np.isin(got[:, 0], y32).all() and np.isin(got[:, -1], y32).all()
    assert got.min() >= y32.min() and got.max() <= y32.max()


@pytest.mark.parametrize("name", _golden.CASES)
@pytest.mark.parametrize("warp_elems,select_bits", [(-1, 10), (0, 10), (0, 3), (0, 1)])
def test_fast_mode_against_reference_outputs(name, warp_elems, select_bits):
    # QF_MODE_FAST (histogram selection instead of the sort) against the reference's own
    # outputs: 1e-4 relative (BASELINE.json north_star, float32 arithmetic), order
    # statistics (q = 0 and q = 100) bit-exact
    from skgarden_b200 import _native
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, warp_elems=warp_elems, select_bits=select_bits)   # few bits: several refinement levels
    X = _cuda(g.X_test)
    scale = np.abs(g.y_train).max()
    for lo in range(0, len(g.q), 7):                    # batches of <= 16 percentiles take the selection kernel
        qs = g.q[lo:lo + 7]
        got = dev.predict_quantiles(X, qs, mode=_native.QF_MODE_FAST).cpu().numpy().T
        want = g.pred[lo:lo + 7]
        assert_allclose(got, want, rtol=1e-4, atol=1e-4 * scale)
        for k, q in enumerate(qs):
            if q in (0.0, 100.0):
                assert_array_equal(got[k], want[k])
    dev.close()


@pytest.mark.parametrize("kind,msl,depth,T", [("et", 20, None, 100), ("rf", 1, 7, 60), ("et", 5, None, 300)])
def test_fast_mode_large_member_sets(kind, msl, depth, T):
    # rows with thousands of members: the 128/256/512-thread selection kernels against the
    # bit-exact sorting kernels on the same forest
    import skgarden_b200 as sg
    cls = sg.RandomForestQuantileRegressor if kind == "rf" else sg.ExtraTreesQuantileRegressor
    rng = np.random.RandomState(11)
    N, n, F = 20000, 3000, 10
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    est = cls(n_estimators=T, random_state=2, min_samples_leaf=msl, max_depth=depth, n_jobs=-1).fit(X[:N], y[:N])
    qs = [0, 1, 5, 25, 50, 75, 95, 99, 100]
    exact = est.predict(X[N:], quantile=qs)
    est.mode = "fast"
    fast = est.predict(X[N:], quantile=qs)
    assert_allclose(fast, exact, rtol=1e-4, atol=1e-4 * np.abs(y).max())
    assert_array_equal(fast[:, [0, -1]], exact[:, [0, -1]])
    est._device_forest.set_option("select_bits", 4)              # same through several refinement levels
    assert_array_equal(est.predict(X[N:], quantile=qs), fast)    #   (multi-level kernel, select_kernel.cuh)
    est._device_forest.set_option("select_bits", 10)
    est._device_forest.set_option("select_impl", 1)              # multi-level kernel, one level
    assert_array_equal(est.predict(X[N:], quantile=qs), fast)
    est._device_forest.set_option("select_impl", 0)
    print("fast vs exact: max abs diff %.3g (|y| max %.3g), members/row %.0f" % (
        np.abs(fast - exact).max(), np.abs(y).max(), est._flat.expected_row_members))


def test_fast_mode_within_tolerance():
    import skgarden_b200 as sg
    rng = np.random.RandomState(5)
    N, n, F = 5000, 2000, 8
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=40, random_state=1, min_samples_leaf=5, n_jobs=-1)
    est.fit(X[:N], y[:N])
    qs = [5, 50, 95]
    exact = est.predict(X[N:], quantile=qs)
    est.mode = "fast"
    fast = est.predict(X[N:], quantile=qs)
    # tolerance stated by BASELINE.json north_star for float32 arithmetic: 1e-4 relative
    # (relative to the scale of y where the value itself is near zero)
    assert_allclose(fast, exact, rtol=1e-4, atol=1e-4 * np.abs(y).max())


def test_streaming_front_end_and_pickle_round_trip():
    # chunked streaming predict, flattened forest on disk, estimator pickling
    import pickle
    import tempfile
    import skgarden_b200 as sg
    rng = np.random.RandomState(9)
    X = rng.randn(3000, 7).astype(np.float32)
    y = 2 * X[:, 0] + np.sin(X[:, 1]) + rng.randn(3000)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=12, min_samples_leaf=3, random_state=0, n_jobs=-1)
    est.fit(X[:2000], y[:2000])
    Xt = X[2000:]
    want = est.predict(Xt, quantile=[10, 50, 90])
    chunks = [Xt[:1], Xt[1:300], Xt[300:300], Xt[300:]]
    got = np.concatenate(list(est.predict_stream(chunks, [10, 50, 90])))
    assert_array_equal(got, want)
    got1 = np.concatenate(list(est.predict_stream(iter(chunks), 50)))
    assert_array_equal(got1, want[:, 1])
    est2 = pickle.loads(pickle.dumps(est))                       # the device handle never travels
    assert_array_equal(est2.predict(Xt, quantile=[10, 50, 90]), want)
    with tempfile.TemporaryDirectory() as d:                     # serving without sklearn objects
        path = d + "/forest.npz"
        est._flat.save(path)
        dev = sg.DeviceForest(sg.FlatForest.load(path), device=0)
        assert_array_equal(dev.predict_quantiles_host(Xt, [10, 50, 90]), want)
        dev.close()


def test_one_member_per_leaf_gather_is_identical():
    # deep bootstrap trees on continuous targets: every leaf holds exactly one in-bag sample
    # (BASELINE.json configs 2 and 4); the warp kernel then skips the offset scan
    import skgarden_b200 as sg
    rng = np.random.RandomState(3)
    X = rng.randn(900, 6).astype(np.float32)
    y = 3 * X[:, 0] + np.sin(3 * X[:, 1]) + rng.randn(900)
    est = sg.RandomForestQuantileRegressor(n_estimators=40, random_state=0, n_jobs=-1).fit(X[:600], y[:600])
    hi = (est._flat.nodes >> np.uint64(32)).astype(np.uint32)
    leaf_sizes = hi[(hi & np.uint32(0x80000000)) != 0] & np.uint32(0x7FFFFFFF)
    assert leaf_sizes.min() == 1 and leaf_sizes.max() == 1
    qs = [0, 5, 33.3, 50, 95, 100]
    got = est.predict(X[600:], quantile=qs)
    leaves = est.apply(X[600:])
    index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
    assert_array_equal(got, O.predict_sparse(index, leaves, qs))
    est._device_forest.set_option("warp_unit", 0)
    assert_array_equal(est.predict(X[600:], quantile=qs), got)


@pytest.mark.parametrize("T,msl,bootstrap", [(1, 8, False), (3, 30, True), (600, 12, True), (1030, 6, False)])
def test_fast_mode_tree_counts_around_the_kernel_limits(T, msl, bootstrap):
    # select2 keeps T / 256 leaves per lane (1..4): one tree, a handful, more than 512, and more
    # than the 1024 it supports (-> multi-level kernel); all against the exact kernels
    import skgarden_b200 as sg
    rng = np.random.RandomState(T)
    N, n, F = 6000, 700, 5
    X = rng.randn(N + n, F).astype(np.float32)
    y = 2 * X[:, 0] + np.sin(2 * X[:, 1]) + rng.randn(N + n)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=T, random_state=0, min_samples_leaf=msl,
                                         bootstrap=bootstrap, max_depth=8, n_jobs=-1).fit(X[:N], y[:N])
    qs = [0, 10, 50, 90, 100]
    est._on_device().set_option("warp_elems", 0)
    exact = est.predict(X[N:], quantile=qs)
    est.mode = "fast"
    fast = est.predict(X[N:], quantile=qs)
    # 1e-4 of the scale of y; the 1030-tree forest gathers 24k members per row (total weight
    # 1030): there the REFERENCE's float32 running sum has drifted by ~1e-3 of a weight by the
    # time it reaches the tails, which moves its interpolation point -- fast mode (exact integer
    # sums) is the more accurate of the two, and the bar is loosened accordingly
    tol = (1e-3 if T > 1000 else 1e-4) * np.abs(y).max()
    assert_allclose(fast, exact, rtol=1e-4, atol=tol)
    assert_array_equal(fast[:, [0, -1]], exact[:, [0, -1]])
    est._device_forest.set_option("select_impl", 1)
    assert_array_equal(est.predict(X[N:], quantile=qs), fast)


def test_one_member_per_leaf_forest_larger_than_the_warp_kernel():
    # T = 300 single-member leaves: more than the 256 keys of the warp kernel -> overflow chain
    import skgarden_b200 as sg
    rng = np.random.RandomState(8)
    X = rng.randn(500, 4).astype(np.float32)
    y = X[:, 0] + rng.randn(500)
    est = sg.RandomForestQuantileRegressor(n_estimators=300, random_state=0, n_jobs=-1).fit(X[:400], y[:400])
    qs = [5, 50, 95]
    got = est.predict(X[400:], quantile=qs)
    index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
    assert_array_equal(got, O.predict_sparse(index, est.apply(X[400:]), qs))


def test_single_row_and_single_feature():
    import skgarden_b200 as sg
    rng = np.random.RandomState(2)
    X = rng.randn(300, 1).astype(np.float32)
    y = np.round(X[:, 0] + 0.3 * rng.randn(300), 1)
    est = sg.RandomForestQuantileRegressor(n_estimators=7, random_state=0, min_samples_leaf=4).fit(X, y)
    index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
    for rows in (X[:1], X[5:6], X[:33]):
        leaves = est.apply(rows)
        assert_array_equal(leaves, O.forest_apply(est.estimators_, rows))
        assert_array_equal(est.predict(rows, quantile=[25, 75]), O.predict_sparse(index, leaves, [25, 75]))
        assert_array_equal(est.predict(rows), O.predict_mean(est.estimators_, rows, leaves))


def test_fast_mode_on_more_than_a_million_training_rows():
    # N in (2^20, 2^21]: the 4096-bucket instantiation of the single-refinement selection kernel
    # (512 ranks per bucket), against the multi-level kernel (bit for bit) and the exact kernels
    import skgarden_b200 as sg
    rng = np.random.RandomState(12)
    N, n, F = 1_200_000, 1500, 4
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + np.sin(3 * X[:, 1]) + rng.randn(N + n)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=6, random_state=0, min_samples_leaf=5, n_jobs=-1).fit(X[:N], y[:N])
    assert est._flat.n_train > (1 << 20)
    dev = est._on_device()
    dev.set_option("warp_elems", 0)
    qs = [0, 5, 50, 95, 100]
    exact = est.predict(X[N:], quantile=qs)
    est.mode = "fast"
    fast = est.predict(X[N:], quantile=qs)
    assert_allclose(fast, exact, rtol=1e-4, atol=1e-4 * np.abs(y).max())
    assert_array_equal(fast[:, [0, -1]], exact[:, [0, -1]])
    dev.set_option("select_impl", 1)
    assert_array_equal(est.predict(X[N:], quantile=qs), fast)

