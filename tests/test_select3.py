"""GPU tests of the QF_MODE_FAST kernel on compact leaf lists (csrc/select3_kernel.cuh,
csrc/compact_lists.cuh) and of the kernels at the sizes bench.py times them (``pytest -m gpu``).
Oracles: committed outputs of the unmodified reference (tests/golden), oracle/qrf_oracle.py and
scikit-learn's own ``tree_.apply``.  Bars: QF_MODE_FAST within 1e-4 of the scale of y (BASELINE.json,
float32), order statistics (q = 0, 100) and QF_MODE_EXACT bit-exact."""
import threading

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

import _atsize
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


FAST = 1


@pytest.mark.parametrize("name", _golden.CASES)
@pytest.mark.parametrize("capw,variant", [(0, 1), (4, 1), (0, 0), (4, 0)])
def test_select3_against_reference_outputs(name, capw, variant):
    # every golden forest through the CTA-level fast path with <= 3 percentiles per call (the
    # select3 kernel wherever the forest is compactable); capw = 4 pushes rows out of the
    # shared-memory stage into the direct-from-global path
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, warp_elems=0, select3_capw=capw, select_impl=3, select3_variant=variant)
    X = _cuda(g.X_test)
    scale = np.abs(g.y_train).max()
    names = set()
    for lo in range(0, len(g.q), 3):
        qs = g.q[lo:lo + 3]
        names.add(dev.quantile_kernel_name(len(g.X_test), len(qs), FAST))
        got = dev.predict_quantiles(X, qs, mode=FAST).cpu().numpy().T
        want = g.pred[lo:lo + 3]
        assert_allclose(got, want, rtol=1e-4, atol=1e-4 * scale)
        for k, q in enumerate(qs):
            if q in (0.0, 100.0):
                assert_array_equal(got[k], want[k])
    if name in ("c1_rfqr_boston", "c1_etqr_boston", "et_msl5", "rf_noboot_msl3_ties", "rf_maxleaf50", "rf_pure_leaves"):
        assert all(n.startswith("quantile_select3_kernel") for n in names), names
    dev.close()


@pytest.mark.parametrize("rows", [1, 2, 3, 4, 5, 7, 8, 9, 33])
def test_select3_ragged_row_counts_and_chunks(rows):
    # row groups of four: every tail length, and chunks that are not multiples of four
    g = _golden.load("et_msl5")
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat, warp_elems=0, select_impl=3)
    qs = [5.0, 50.0, 95.0]
    ref = _dev(flat, warp_elems=0, select_impl=2)
    want = ref.predict_quantiles(_cuda(g.X_test), qs, mode=FAST).cpu().numpy()
    assert dev.quantile_kernel_name(rows, 3, FAST).startswith("quantile_select3_kernel")
    got = dev.predict_quantiles(_cuda(g.X_test[:rows]), qs, mode=FAST).cpu().numpy()
    assert_allclose(got, want[:rows], rtol=2e-5, atol=2e-5)
    dev.set_option("rows_per_chunk", 5)
    got = dev.predict_quantiles(_cuda(g.X_test), qs, mode=FAST).cpu().numpy()
    assert_allclose(got, want, rtol=2e-5, atol=2e-5)
    host = dev.predict_quantiles_host(g.X_test, qs, mode=FAST, chunk_rows=37)
    assert_array_equal(host, got)
    dev.close()
    ref.close()


@pytest.mark.parametrize("T,msl,bootstrap,N", [(1, 8, False, 6000), (3, 30, True, 6000), (255, 5, True, 6000),
                                              (600, 12, True, 6000), (1000, 6, False, 3000)])
def test_select3_tree_counts(T, msl, bootstrap, N):
    # 1, 2 and 4 trees per thread; against the exact kernels (tolerance) and select2 (same fixed-point
    # idea, other rounding of the per-member weight)
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(N, 500, 5, seed=T)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=T, random_state=0, min_samples_leaf=msl,
                                         bootstrap=bootstrap, max_depth=9, n_jobs=-1).fit(Xtr, ytr)
    dev = est._on_device()
    dev.set_option("warp_elems", 0)
    dev.set_option("select_impl", 3)
    for qs in ([0, 50, 100], [2.5], [33.3, 99.9]):
        exact = est.predict(Xt, quantile=qs)
        est.mode = "fast"
        name = dev.quantile_kernel_name(len(Xt), len(qs), FAST)
        fast = est.predict(Xt, quantile=qs)
        est.mode = "exact"
        if T <= 3:                               # few trees: fixed point is precise whatever the leaf sizes (long lists included)
            assert name.startswith("quantile_select3_kernel"), name
        tol = (1e-3 if T >= 1000 else 1e-4) * np.abs(ytr).max()
        assert_allclose(fast, exact, rtol=1e-4, atol=tol)
        if qs[0] == 0:
            assert_array_equal(fast[:, [0, -1]], exact[:, [0, -1]])


def test_select3_more_than_a_million_training_rows():
    # N in (2^20, 2^21]: the 4096-bucket instantiation
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(1_200_000, 1500, 4, seed=12)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=6, random_state=0, min_samples_leaf=5, n_jobs=-1).fit(Xtr, ytr)
    dev = est._on_device()
    dev.set_option("warp_elems", 0)
    dev.set_option("select_impl", 3)
    qs = [0, 50, 100]
    exact = est.predict(Xt, quantile=qs)
    est.mode = "fast"
    from skgarden_b200 import _native
    name = dev.quantile_kernel_name(len(Xt), 3, FAST)
    assert name .startswith("quantile_select3_kernel<12, 1,"), (name, _native.last_error())
    fast = est.predict(Xt, quantile=qs)
    assert_allclose(fast, exact, rtol=1e-4, atol=1e-4 * np.abs(ytr).max())
    assert_array_equal(fast[:, [0, -1]], exact[:, [0, -1]])


def test_tree_level_estimator_fast_mode_unit_weights():
    # unit member weights (quantile/tree.py:45-47): s = 1, count = 1 in the compact lists
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(4000, 300, 4, seed=3, ties=True)
    est = sg.DecisionTreeQuantileRegressor(min_samples_leaf=30, random_state=0).fit(Xtr, ytr)
    exact = est.predict(Xt, quantile=50)
    dev = est._on_device()
    dev.set_option("warp_elems", 0)
    dev.set_option("select_impl", 3)
    assert dev.quantile_kernel_name(len(Xt), 3, FAST).startswith("quantile_select3_kernel")
    got = dev.predict_quantiles(_cuda(Xt), [5.0, 50.0, 95.0], mode=FAST).cpu().numpy()
    assert_allclose(got[:, 1], exact, rtol=1e-4, atol=1e-4 * np.abs(ytr).max())


# ---- the kernels at the sizes bench.py times them -------------------------
def test_c2_shaped_forest_against_oracle():
    # RFQR T=100 msl=1 bootstrap, 100k x 20 train, 100k test rows: warp kernel with the
    # one-member-per-leaf gather and the 8192-bucket locality order, 2000 sampled rows bit-exact
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(100_000, 100_000, 20, seed=0)
    est = sg.RandomForestQuantileRegressor(n_estimators=100, min_samples_leaf=1, random_state=0, n_jobs=-1).fit(Xtr, ytr)
    qs = [5, 33.3, 50, 95, 99.9]
    got = est.predict(Xt, quantile=qs)
    rows = np.random.RandomState(1).choice(len(Xt), 2000, replace=False)
    leaves = np.array([e.tree_.apply(Xt[rows]) for e in est.estimators_]).T      # scikit-learn's own apply
    assert_array_equal(est.apply(Xt[rows]), leaves)
    want = O.predict_sparse(_atsize.thin_index(est, leaves), leaves, qs)
    assert_array_equal(got[rows], want)
    est.mode = "fast"
    fast = est.predict(Xt, quantile=[5, 50, 95])
    assert_allclose(fast[rows], want[:, [0, 2, 3]], rtol=1e-4, atol=1e-4 * np.abs(ytr).max())


def test_c3_shaped_forest_against_oracle():
    # ETQR T=500 msl=5 bootstrap, 250k x 50 train, 250k test rows (C3 with a quarter of the rows:
    # same trees per thread, same leaf sizes, 262144-row chunks): select2 and select3 (fast) within 1e-4
    # and the bucket kernel (exact) bit-exact on 500 sampled rows
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(250_000, 250_000, 50, seed=0)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=500, min_samples_leaf=5, random_state=0, n_jobs=-1).fit(Xtr, ytr)
    qs = [5, 50, 95]
    rows = np.random.RandomState(2).choice(len(Xt), 500, replace=False)
    leaves = np.array([e.tree_.apply(Xt[rows]) for e in est.estimators_]).T
    assert_array_equal(est.apply(Xt[rows]), leaves)
    want = O.predict_sparse(_atsize.thin_index(est, leaves), leaves, qs)
    exact = est.predict(Xt, quantile=qs)
    assert_array_equal(exact[rows], want)
    est.mode = "fast"
    dev = est._on_device()
    assert dev.quantile_kernel_name(len(Xt), 3, FAST) == "quantile_select2_kernel<11>"       # the default fast kernel
    fast2 = est.predict(Xt, quantile=qs)
    scale = np.abs(ytr).max()
    assert_allclose(fast2[rows], want, rtol=1e-4, atol=1e-4 * scale)
    assert_allclose(fast2, exact, rtol=1e-4, atol=2e-4 * scale)       # all rows against the exact kernels
    dev.set_option("select_impl", 3)                                   # select3 on the compact leaf lists
    assert dev.quantile_kernel_name(len(Xt), 3, FAST) .startswith("quantile_select3_kernel<11, 2,")
    fast = est.predict(Xt, quantile=qs)
    assert_allclose(fast[rows], want, rtol=1e-4, atol=1e-4 * scale)
    assert_allclose(fast, exact, rtol=1e-4, atol=2e-4 * scale)
    assert_allclose(fast, fast2, rtol=1e-5, atol=1e-5 * scale)


def test_two_host_threads_on_one_handle():
    # the handle's per-call state lives in the call (csrc/qforest_abi.cu CallScope): two threads,
    # each with its own stream, workspace and output, give the single-threaded results
    import skgarden_b200 as sg
    g = _golden.load("et_msl5")
    flat = _golden.flat_from_golden(g)
    dev = _dev(flat)
    X = _cuda(np.tile(g.X_test, (20, 1)))
    want = dev.predict_quantiles(X, g.q).cpu().numpy()
    want_host = dev.predict_quantiles_host(np.tile(g.X_test, (20, 1)), g.q)
    assert_array_equal(want_host, want)
    results, errors = {}, []

    def device_worker(k):
        try:
            from skgarden_b200 import _native
            import ctypes as C
            st = torch.cuda.Stream()
            n, q = X.shape[0], np.ascontiguousarray(g.q, dtype=np.float64)
            need = int(dev._lib.qf_forest_workspace_bytes(dev._h, n, len(q)))
            with torch.cuda.stream(st):
                ws = torch.empty(need, dtype=torch.uint8, device="cuda")
                for _ in range(5):
                    out = torch.empty((n, len(q)), dtype=torch.float64, device="cuda")
                    _native.check(dev._lib.qf_forest_predict_quantiles(
                        dev._h, C.c_void_p(X.data_ptr()), n, X.stride(0), q.ctypes.data_as(C.c_void_p), len(q),
                        C.c_void_p(out.data_ptr()), 0, C.c_void_p(ws.data_ptr()), ws.numel(), C.c_void_p(st.cuda_stream)))
                st.synchronize()
            results[k] = out.cpu().numpy()
        except Exception as exc:                                   # pragma: no cover
            errors.append(exc)

    def host_worker(k):
        try:
            for _ in range(3):
                results[k] = dev.predict_quantiles_host(np.tile(g.X_test, (20, 1)), g.q, chunk_rows=64)
        except Exception as exc:                                   # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=device_worker, args=(0,)), threading.Thread(target=device_worker, args=(1,)),
               threading.Thread(target=host_worker, args=(2,)), threading.Thread(target=host_worker, args=(3,))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        assert_array_equal(results[k], want)
    dev.close()


def test_more_features_than_the_shared_memory_tile():
    # 1700 features x 32 rows x 4 bytes > 200 KB: the leaf lookup falls back to one thread per row
    # reading features from global memory (used to fail with QF_ERR_LIMIT)
    import skgarden_b200 as sg
    rng = np.random.RandomState(4)
    X = rng.randn(700, 1700).astype(np.float32)
    y = X[:, 0] + X[:, 1500] + 0.1 * rng.randn(700)
    est = sg.RandomForestQuantileRegressor(n_estimators=5, random_state=0, min_samples_leaf=3,
                                           max_features=0.5).fit(X[:500], y[:500])
    Xt = X[500:]
    leaves = est.apply(Xt)
    assert_array_equal(leaves, O.forest_apply(est.estimators_, Xt))
    index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
    assert_array_equal(est.predict(Xt, quantile=[10, 50, 90]), O.predict_sparse(index, leaves, [10, 50, 90]))
    assert_array_equal(est.predict(Xt), O.predict_mean(est.estimators_, Xt, leaves))


def test_devices_list_shards_rows_over_replicas():
    # devices=[...]: forest replicated, rows of every call sharded, one host thread per GPU; the
    # result does not depend on the split (ensemble.py:137: rows are independent).  Uses two GPUs
    # when the box has them, else the one GPU as a single replica.
    import skgarden_b200 as sg
    n_dev = torch.cuda.device_count()
    devs = [0, 1] if n_dev >= 2 else [0]
    Xtr, ytr, Xt = _atsize.make_data(5000, 3001, 6, seed=5)
    kw = dict(n_estimators=20, min_samples_leaf=4, random_state=0, n_jobs=-1)
    one = sg.ExtraTreesQuantileRegressor(**kw).fit(Xtr, ytr)
    many = sg.ExtraTreesQuantileRegressor(devices=devs, **kw).fit(Xtr, ytr)
    qs = [5, 50, 95]
    assert_array_equal(many.predict(Xt, quantile=qs), one.predict(Xt, quantile=qs))
    assert_array_equal(many.predict(Xt), one.predict(Xt))
    assert_array_equal(many.apply(Xt), one.apply(Xt))
    chunks = [Xt[:1], Xt[1:700], Xt[700:700], Xt[700:2000], Xt[2000:]]
    got = np.concatenate(list(many.predict_stream(chunks, qs)))
    assert_array_equal(got, one.predict(Xt, quantile=qs))
    std_one = sg.RandomForestRegressor(n_estimators=8, random_state=0, min_samples_leaf=3).fit(Xtr, ytr)
    std_many = sg.RandomForestRegressor(n_estimators=8, random_state=0, min_samples_leaf=3, devices=devs).fit(Xtr, ytr)
    m1, s1 = std_one.predict(Xt, return_std=True)
    m2, s2 = std_many.predict(Xt, return_std=True)
    assert_array_equal(m1, m2)
    assert_array_equal(s1, s2)


# ---- select4: the warp-per-row form of the same arithmetic (csrc/select4_kernel.cuh) ---------------
@pytest.mark.parametrize("name", _golden.CASES)
@pytest.mark.parametrize("hits,extras", [(256, 0), (0, 0), (256, 32), (3, 32)])
def test_select4_equals_select3_bit_for_bit(name, hits, extras):
    # same fixed-point weights and the same definition of the crossing entry and its neighbours: the
    # sorted-hit-list path (hits=256), the exact narrowing path (hits=0: every row overflows the list;
    # hits=3: some do) and the leaf-by-leaf second sweep (extras=32 entries only) must return the bits
    # select3 returns, which tests above pin to the reference's outputs within 1e-4
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    d3 = _dev(flat, warp_elems=0, select_impl=3)
    d4 = _dev(flat, warp_elems=0, select_impl=4, select4_hits=hits, select4_list=extras)
    X = _cuda(g.X_test)
    scale = np.abs(g.y_train).max()
    for lo in range(0, len(g.q), 3):
        qs = g.q[lo:lo + 3]
        n3 = d3.quantile_kernel_name(len(g.X_test), len(qs), FAST)
        n4 = d4.quantile_kernel_name(len(g.X_test), len(qs), FAST)
        got = d4.predict_quantiles(X, qs, mode=FAST).cpu().numpy()
        if n3.startswith("quantile_select3") and n4.startswith("quantile_select4"):
            assert_array_equal(got, d3.predict_quantiles(X, qs, mode=FAST).cpu().numpy())
        assert_allclose(got.T, g.pred[lo:lo + 3], rtol=1e-4, atol=1e-4 * scale)
    if name in ("c1_rfqr_boston", "et_msl5", "rf_noboot_msl3_ties"):
        assert d4.quantile_kernel_name(len(g.X_test), 3, FAST).startswith("quantile_select4_kernel")
    d3.close()
    d4.close()


@pytest.mark.parametrize("T,msl,N", [(1, 8, 6000), (40, 5, 6000), (600, 12, 6000), (1000, 6, 3000)])
@pytest.mark.parametrize("rows", [1, 33, 500])
def test_select4_tree_counts_and_row_counts(T, msl, N, rows):
    import skgarden_b200 as sg
    Xtr, ytr, Xt = _atsize.make_data(N, 500, 5, seed=T)
    est = sg.ExtraTreesQuantileRegressor(n_estimators=T, random_state=0, min_samples_leaf=msl, max_depth=9,
                                         n_jobs=-1, mode="fast").fit(Xtr, ytr)
    dev = est._on_device()
    dev.set_option("warp_elems", 0)
    qs = [5, 50, 95]
    dev.set_option("select_impl", 3)
    want = est.predict(Xt[:rows], quantile=qs)
    is3 = dev.quantile_kernel_name(rows, 3, FAST).startswith("quantile_select3")
    dev.set_option("select_impl", 4)
    if dev.quantile_kernel_name(rows, 3, FAST) != "quantile_select4_kernel":
        pytest.skip("fast mode is not precise enough for this forest (leaves too heavy): the exact kernels serve it")
    scale = np.abs(ytr).max()

    def same(got, ref):
        if is3:                                         # same integers, same float32 formulas
            assert_array_equal(got, ref)
        else:                                           # select3 left this forest to select2 (other rounding of the weights)
            assert_allclose(got, ref, rtol=2e-5, atol=2e-5 * scale)
    for ctas, hints in ((4, 0), (3, 1), (1, 1)):
        dev.set_option("select4_ctas", ctas)
        dev.set_option("select4_unroll", 2 if hints else 4)
        dev.set_option("select4_hints", hints)
        dev.set_option("select4_hits", 256)
        same(est.predict(Xt[:rows], quantile=qs), want)
        dev.set_option("select4_hits", 16)
        same(est.predict(Xt[:rows], quantile=qs), want)
        same(est.predict(Xt[:rows], quantile=[50]), want[:, 1:2])


@pytest.mark.parametrize("name", _golden.CASES)
@pytest.mark.parametrize("impl", [3, 4])
def test_compact_lists_aligned_to_64_bytes_same_bits(name, impl):
    # option compact_align = 1 only moves multi-octet lists to even octets (one spare octet each): same results
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    d0 = _dev(flat, warp_elems=0, select_impl=impl)
    d1 = _dev(flat, warp_elems=0, select_impl=impl, compact_align=1)
    X = _cuda(g.X_test)
    qs = g.q[:3]
    assert d0.quantile_kernel_name(len(g.X_test), 3, FAST) == d1.quantile_kernel_name(len(g.X_test), 3, FAST)
    assert_array_equal(d1.predict_quantiles(X, qs, mode=FAST).cpu().numpy(),
                       d0.predict_quantiles(X, qs, mode=FAST).cpu().numpy())
    d0.close()
    d1.close()
