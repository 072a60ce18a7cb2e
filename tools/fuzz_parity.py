#!/usr/bin/env python
"""Randomised parity sweep on the GPU box (compute-sanitizer is not available on this pool):
random forest shapes through every quantile kernel.
  exact mode      == oracle (sparse restatement of the reference), bit for bit, on sampled rows
  fast, select2   == fast, multi-level selection kernel, bit for bit, all rows
  fast, select3   ~= fast, select2 (2e-5, or inside the float64 bracket for leaves of hundreds of members); order statistics equal
  fast, select4   == fast, select3, bit for bit (also through the narrowing path)
  fast            ~= exact within the float32 tolerance
  device builder  == host flattener, bit for bit (when no leaf exceeds 4096 members)
Prints one line per case; exits non-zero on the first mismatch."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def float64_percentile(index, est, row_leaves, q, T):
    """utils.py:68-81 in float64 on the float32 weights of one row (what fast mode approximates)."""
    js = np.concatenate([index.leaf_members(t, row_leaves[t]) for t in range(T)])
    ws = np.concatenate([index.weights[t, index.leaf_members(t, row_leaves[t])] for t in range(T)]).astype(np.float64)
    uniq, inv = np.unique(js, return_inverse=True)
    W = np.bincount(inv, weights=ws)
    order = np.argsort(index.rank[uniq])
    a = np.asarray(est.y_train_, dtype=np.float32)[uniq][order].astype(np.float64)
    W = W[order]
    keep = W != 0
    a, W = a[keep], W[keep]
    cum = np.cumsum(W)
    part = 100.0 / cum[-1] * (cum - W / 2.0)
    st = np.searchsorted(part, q) - 1
    if st == len(cum) - 1:
        return a[-1]
    if st == -1:
        return a[0]
    return a[st] + (q - part[st]) / (part[st + 1] - part[st]) * (a[st + 1] - a[st])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", type=int, default=-1, help="run just this case (same random stream)")
    args = ap.parse_args()
    import skgarden_b200 as sg
    from skgarden_b200 import _native
    from oracle import qrf_oracle as O
    rng = np.random.RandomState(args.seed)
    for case in range(args.cases):
        N = int(rng.choice([60, 300, 2000, 9000, 40000]))
        n = 257
        F = int(rng.choice([1, 3, 8, 40]))
        T = int(rng.choice([1, 2, 7, 33, 100, 260, 530]))
        if N * T > 6_000_000:
            T = max(1, 6_000_000 // N)
        kind = rng.choice(["rf", "et"])
        msl = int(rng.choice([1, 2, 5, 30]))
        depth = rng.choice([None, 3, 9])
        bootstrap = bool(rng.rand() < 0.6)
        ties = bool(rng.rand() < 0.3)
        X = rng.randn(N + n, F).astype(np.float32)
        y = 3 * X[:, 0] + np.sin(2 * X[:, min(1, F - 1)]) + rng.randn(N + n)
        if ties:
            y = np.round(y, 1)
        cls = sg.RandomForestQuantileRegressor if kind == "rf" else sg.ExtraTreesQuantileRegressor
        forest_seed = int(rng.randint(1 << 30))
        nq = int(rng.choice([1, 3, 4, 9, 16]))
        qs = sorted(set([float(q) for q in rng.choice([0, 0.5, 1, 5, 25, 33.3, 50, 62.5, 75, 95, 99, 99.9, 100], nq)]))
        sample_rows = rng.choice(n, 12, replace=False)
        mv = float(rng.choice([0.0, 0.01, 1.0]))
        chunk = int(rng.choice([131072, 100, 33]))       # several kernel chunks per call
        sort_rows = int(rng.choice([-1, 0, 1]))
        if args.only >= 0 and case != args.only:
            continue
        est = cls(n_estimators=T, random_state=forest_seed, min_samples_leaf=msl, max_depth=depth,
                  bootstrap=bootstrap, n_jobs=-1).fit(X[:N], y[:N])
        Xt = X[N:]
        dev = est._on_device()
        dev.set_option("rows_per_chunk", chunk)
        dev.set_option("sort_rows", sort_rows)
        tag = "case %3d %s N=%5d T=%3d F=%2d msl=%2d depth=%s boot=%d ties=%d Q=%d chunk=%d sort=%d members/row=%.0f" % (
            case, kind, N, T, F, msl, depth, bootstrap, ties, len(qs), chunk, sort_rows, est._flat.expected_row_members)
        leaves = est.apply(Xt)
        assert np.array_equal(leaves, O.forest_apply(est.estimators_, Xt)), tag + ": apply"
        index = O.SparseIndex(est.y_train_, est.y_train_leaves_, est.y_weights_, est._flat.sorter)
        rows = sample_rows
        want = O.predict_sparse(index, leaves[rows], qs)
        results = {}
        hi = (est._flat.nodes >> np.uint64(32)).astype(np.uint32)
        biggest = int((hi[(hi >> 31) == 1] & np.uint32(0x7FFFFFFF)).max())
        for warp in (-1, 0):                       # warp kernel first / CTA kernels only
            dev.set_option("warp_elems", warp)
            est.mode = "exact"
            exact = est.predict(Xt, quantile=qs)
            assert np.array_equal(exact[rows], want), tag + ": exact vs oracle (warp_elems=%d)" % warp
            est.mode = "fast"
            dev.set_option("select_impl", 2)
            fast = est.predict(Xt, quantile=qs)
            dev.set_option("select_impl", 1)
            fast1 = est.predict(Xt, quantile=qs)
            assert np.array_equal(fast, fast1), tag + ": select2 vs multi-level kernel (warp_elems=%d)" % warp
            dev.set_option("select_impl", 3)       # select3 on compact lists when <= 3 percentiles
            fast3 = est.predict(Xt, quantile=qs)
            is3 = dev.quantile_kernel_name(len(Xt), len(qs), 1).startswith("quantile_select3")
            # same fixed-point idea, other rounding of the per-member weight (count * U instead of rn(w * scale))
            if not np.allclose(fast3, fast, rtol=2e-5, atol=2e-5 * np.abs(y).max()):
                # Leaves of hundreds of members: U = rn(scale / count sum) carries a relative error of up to
                # eps = count sum * T / 2^33, which moves every knot of the piecewise-linear percentile by at
                # most 200 * eps percent points; the result must then lie between the float64 percentiles
                # at q - delta and q + delta (the function is monotone in q).
                delta = 100.0 * 4.0 * biggest * T / 2.0 ** 32 + 1e-6
                d = np.abs(fast3 - fast)
                small = 1e-5 * np.abs(y).max()
                for flat_i in np.argsort(d, axis=None)[::-1][:12]:
                    i, k = np.unravel_index(flat_i, d.shape)
                    lo = float64_percentile(index, est, leaves[i], max(qs[k] - delta, 0.0), T)
                    hi_ = float64_percentile(index, est, leaves[i], min(qs[k] + delta, 100.0), T)
                    for nm, v in (("select3", fast3[i, k]), ("select2", fast[i, k])):
                        assert lo - small <= v <= hi_ + small, (
                            "%s: row %d q=%g: %s %.9g outside [%.9g, %.9g] (delta %.3g; select2 %.9g select3 %.9g)"
                            % (tag, i, qs[k], nm, v, lo, hi_, delta, fast[i, k], fast3[i, k]))
                tag += " [select3/select2 differ by %.2g on large leaves, both inside the float64 bracket]" % d.max()
            dev.set_option("select_impl", 4)       # warp per row, the same integers as select3
            if is3 and dev.quantile_kernel_name(len(Xt), len(qs), 1).startswith("quantile_select4"):
                fast4 = est.predict(Xt, quantile=qs)
                assert np.array_equal(fast4, fast3), tag + ": select4 vs select3 (warp_elems=%d), max diff %g" % (
                    warp, np.abs(fast4 - fast3).max())
                dev.set_option("select4_hits", 8)
                assert np.array_equal(est.predict(Xt, quantile=qs), fast3), tag + ": select4 narrowing path"
                dev.set_option("select4_hits", 256)
            ends3 = [k for k, q in enumerate(qs) if q in (0.0, 100.0)]
            assert np.array_equal(fast3[:, ends3], fast[:, ends3]), tag + ": select3 order statistics"
            dev.set_option("select_impl", 0)
            # float32 drift of the reference's own running sum grows with the total weight of a row
            tol = max(1e-4, 2e-6 * T) * np.abs(y).max()
            if not np.allclose(fast, exact, rtol=1e-4, atol=tol):
                # Many tiny weights (leaves of thousands of members): the REFERENCE's float32 running
                # sum drifts far enough to move its crossing entry; with tie-heavy y that is a whole
                # step of y.  Fast mode sums exactly: it must then agree with float64 arithmetic.
                d = np.abs(fast - exact)
                for flat_i in np.argsort(d, axis=None)[::-1][:8]:
                    i, k = np.unravel_index(flat_i, d.shape)
                    if d[i, k] <= tol:
                        break
                    v64 = float64_percentile(index, est, leaves[i], qs[k], T)
                    assert abs(fast[i, k] - v64) <= tol, (
                        "%s: row %d q=%g: fast %.9g exact %.9g float64 %.9g" % (tag, i, qs[k], fast[i, k], exact[i, k], v64))
                tag += " [reference float32 drift on %d elements]" % int((d > tol).sum())
            ends = [k for k, q in enumerate(qs) if q in (0.0, 100.0)]
            assert np.array_equal(fast[:, ends], exact[:, ends]), tag + ": order statistics"
            results[warp] = exact
        assert np.array_equal(results[-1], results[0]), tag + ": warp vs CTA kernels"
        built = "-"
        if biggest <= 4096:
            est.builder = "device"
            fl = est._flatten(X[:N], est.y_train_)
            assert np.array_equal(fl.nodes, est._flat.nodes) and np.array_equal(fl.members, est._flat.members), tag + ": K4"
            est.builder = "host"
            built = "ok"
        dev.close()
        # tree-level estimator (quantile/tree.py) and return_std forest (forest.py:133-178) on the same data
        tcls = sg.DecisionTreeQuantileRegressor if kind == "rf" else sg.ExtraTreeQuantileRegressor
        tq = tcls(random_state=forest_seed, min_samples_leaf=msl, max_depth=depth).fit(X[:N], y[:N])
        tl = tq.apply(Xt)
        assert np.array_equal(tl, O.tree_apply(tq.tree_, Xt)), tag + ": tree apply"
        for q in qs[:3]:
            qq = int(q) if float(q).is_integer() else q
            assert np.array_equal(tq.predict(Xt, quantile=qq),
                                  O.tree_predict_quantile(tq.y_train_, tq.y_train_leaves_, tl, qq)), tag + ": tree quantile %g" % q
        scls = sg.RandomForestRegressor if kind == "rf" else sg.ExtraTreesRegressor
        sf = scls(n_estimators=min(T, 20), random_state=forest_seed, min_samples_leaf=msl, max_depth=depth,
                  bootstrap=bootstrap, min_variance=mv, n_jobs=-1).fit(X[:N], y[:N])
        mean, std = sf.predict(Xt, return_std=True)
        sl = sf.apply(Xt)
        assert np.array_equal(mean, O.predict_mean(sf.estimators_, Xt, sl)), tag + ": forest mean"
        assert np.array_equal(std, O.forest_return_std(sf.estimators_, sl, mean, mv)), tag + ": return_std"
        print(tag, "| largest leaf %d | K4 %s | tree + return_std ok | ok" % (biggest, built), flush=True)
    print("all %d cases ok" % args.cases)


if __name__ == "__main__":
    main()
