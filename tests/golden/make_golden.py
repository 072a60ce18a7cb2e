"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

For every case it fits the reference's own estimator (loaded through
oracle/reference_shim.py, no reference file edited), and stores the inputs, the fitted
trees as raw arrays, the reference's fit attributes and the reference's outputs:

    X_train f32, y_train f64, X_test f32
    tree arrays (concatenated, with node offsets): children_left/right, feature,
        threshold f64, value f64; per-tree integer seeds (est.random_state)
    sorter        = np.argsort(y_train_)            (ensemble.py:133, same process)
    apply         = est.apply(X_test)               (forest.py:583-604)       int64 (n,T)
    y_train_leaves_, y_weights_                     (ensemble.py:83-101)
    quantiles q[] and predict(X_test, q) for each   (ensemble.py:104-143)     f64 (Q,n)
    mean          = est.predict(X_test)             (forest.py:1093-1131)

The fixtures travel to the GPU box (the reference does not), so the GPU parity tests
compare the CUDA path with the reference's real outputs without needing /root/reference.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import reference_shim as rs  # noqa: E402

# the last three are not float32-representable: the searchsorted compare of utils.py:73 runs in float64
QS = [0, 1, 5, 10, 25, 30, 50, 62.5, 75, 90, 95, 97.5, 99, 100, 0.1, 33.3, 99.9]


def boston_shaped(seed=0):
    """506x13 stand-in for sklearn's removed load_boston (test_ensemble.py:13-16)."""
    rng = np.random.RandomState(seed)
    scale = np.array([8.6, 23.3, 6.9, 0.25, 0.12, 0.7, 28.1, 2.1, 8.7, 168.5, 2.2, 91.3, 7.1])
    shift = np.array([3.6, 11.4, 11.1, 0.07, 0.55, 6.3, 68.6, 3.8, 9.5, 408.2, 18.5, 356.7, 12.7])
    Z = rng.randn(506, 13)
    X = np.abs(Z * scale + shift)
    X[:, 3] = (rng.rand(506) < 0.07).astype(float)      # CHAS-like binary column
    X[:, 8] = np.round(X[:, 8])                          # RAD-like integer column
    y = 22.5 + 4.0 * Z[:, 5] - 3.5 * Z[:, 12] + 1.5 * Z[:, 0] * Z[:, 7] + 2.0 * rng.randn(506)
    y = np.round(np.clip(y, 5.0, 50.0), 1)               # tie-heavy targets
    return X, y


def synth(N, n, F, seed, ties=False):
    rng = np.random.RandomState(seed)
    X = rng.randn(N + n, F).astype(np.float32)
    y = 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(N + n)
    if ties:
        y = np.round(y, 1)
    return X[:N], y[:N], X[N:]


def dump(name, est, X_train, y_train, X_test, bootstrap, kind):
    trees = [e.tree_ for e in est.estimators_]
    off = np.cumsum([0] + [t.node_count for t in trees])
    out = dict(
        kind=np.array(kind), bootstrap=np.array(bootstrap),
        X_train=np.asarray(X_train, dtype=np.float32), y_train=np.asarray(y_train, dtype=np.float64),
        X_test=np.asarray(X_test, dtype=np.float32),
        node_offsets=off,
        children_left=np.concatenate([t.children_left for t in trees]),
        children_right=np.concatenate([t.children_right for t in trees]),
        feature=np.concatenate([t.feature for t in trees]),
        threshold=np.concatenate([t.threshold for t in trees]),
        value=np.concatenate([t.value[:, 0, 0] for t in trees]),
        seeds=np.array([e.random_state for e in est.estimators_], dtype=np.int64),
        sorter=np.argsort(est.y_train_),
        apply=est.apply(X_test),
        y_train_leaves_=est.y_train_leaves_, y_weights_=est.y_weights_,
        q=np.array(QS, dtype=np.float64),
        pred=np.array([est.predict(X_test, quantile=(int(q) if float(q).is_integer() else q))
                       for q in QS]),
        mean=est.predict(X_test),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s T=%3d N=%5d n=%4d nodes=%6d  %7.1f KB" % (
        name, len(trees), len(y_train), len(X_test), off[-1], os.path.getsize(path) / 1024))


def main():
    ref = rs.load()
    KW = rs.MODERN_KW
    RF, ET = ref.RandomForestQuantileRegressor, ref.ExtraTreesQuantileRegressor

    # C1 (BASELINE.json configs[0]): RFQR n_estimators=10 on Boston-shaped 506x13, 60/40
    Xb, yb = boston_shaped()
    perm = np.random.RandomState(0).permutation(506)
    tr, te = perm[:303], perm[303:]
    Xtr, Xte = Xb[tr].astype(np.float32), Xb[te].astype(np.float32)
    est = RF(n_estimators=10, random_state=0, **KW).fit(Xtr, yb[tr])
    dump("c1_rfqr_boston", est, Xtr, yb[tr], Xte, True, "rf")
    est = ET(n_estimators=10, random_state=0, **KW).fit(Xtr, yb[tr])
    dump("c1_etqr_boston", est, Xtr, yb[tr], Xte, True, "et")

    # bootstrap off, small leaves, tie-heavy y
    Xtr, ytr, Xte = synth(1200, 150, 6, seed=1, ties=True)
    est = RF(n_estimators=8, random_state=1, bootstrap=False, min_samples_leaf=3, **KW).fit(Xtr, ytr)
    dump("rf_noboot_msl3_ties", est, Xtr, ytr, Xte, False, "rf")

    # C3-shaped in miniature: extra trees, msl=5, bootstrap, continuous y
    Xtr, ytr, Xte = synth(2000, 200, 10, seed=2)
    est = ET(n_estimators=20, random_state=2, min_samples_leaf=5, **KW).fit(Xtr, ytr)
    dump("et_msl5", est, Xtr, ytr, Xte, True, "et")

    # best-first builder: node order is not depth-first (max_leaf_nodes)
    Xtr, ytr, Xte = synth(1500, 150, 5, seed=3)
    est = RF(n_estimators=6, random_state=3, max_leaf_nodes=50, **KW).fit(Xtr, ytr)
    dump("rf_maxleaf50", est, Xtr, ytr, Xte, True, "rf")

    # shallow trees: every row gathers about T*N/4 members (large candidate sets)
    Xtr, ytr, Xte = synth(3000, 60, 4, seed=4, ties=True)
    est = ET(n_estimators=5, random_state=4, max_depth=2, **KW).fit(Xtr, ytr)
    dump("et_depth2_large_sets", est, Xtr, ytr, Xte, True, "et")

    # pure leaves, bootstrap off: quantile == mean (test_ensemble.py:69-83)
    rng = np.random.RandomState(0)
    Xp = rng.randn(10, 1).astype(np.float32)
    yp = np.linspace(0.0, 100.0, 10)
    est = RF(n_estimators=10, random_state=0, bootstrap=False, max_depth=None, **KW).fit(Xp, yp)
    dump("rf_pure_leaves", est, Xp, yp, Xp, False, "rf")


if __name__ == "__main__":
    main()
