"""Synthetic workloads of BASELINE.md section 5 (there is no network for real data)."""
import numpy as np

# name -> (estimator, n_estimators, min_samples_leaf, N_train, n_test, F, quantiles)
CONFIGS = {
    # BASELINE.json configs[0]: reference plumbing case
    "c1": dict(kind="rf", n_estimators=10, min_samples_leaf=1, n_train=303, n_test=203, n_features=13,
               quantiles=[50.0], label="RFQR T=10, 506x13 Boston-shaped (60/40 split), q=50"),
    # configs[1]: the configuration the headline metric is quoted on at 1 GPU
    "c2": dict(kind="rf", n_estimators=100, min_samples_leaf=1, n_train=100_000, n_test=100_000,
               n_features=20, quantiles=[5.0, 50.0, 95.0],
               label="RFQR T=100 msl=1 bootstrap, 100k x 20 train, 100k test rows, q={5,50,95}"),
    # configs[2]
    "c3": dict(kind="et", n_estimators=500, min_samples_leaf=5, n_train=1_000_000, n_test=1_000_000,
               n_features=50, quantiles=[5.0, 50.0, 95.0],
               label="ETQR T=500 msl=5 bootstrap, 1M x 50 train, 1M test rows, q={5,50,95}"),
    # configs[3]
    "c4": dict(kind="rf", n_estimators=200, min_samples_leaf=1, n_train=1_000_000, n_test=1_000_000,
               n_features=20, quantiles=[float(q) for q in range(1, 100)],
               label="RFQR T=200 msl=1 bootstrap, 1M x 20 train, 1M test rows, q=1..99"),
    # configs[3] stress variant (hard part 6): depth-10 trees, leaves of ~1000 members,
    # ~125k members per row -- genuinely large candidate sets (exact kernels: dense fallback)
    "c4_depth10": dict(kind="rf", n_estimators=200, min_samples_leaf=1, max_depth=10, n_train=1_000_000,
                       n_test=20_000, n_features=20, quantiles=[5.0, 50.0, 95.0],
                       label="RFQR T=200 max_depth=10 bootstrap, 1M x 20 train, 20k test rows, q={5,50,95} (large candidate sets)"),
    # configs[4]: streamed test rows (bench.py keeps the 10M x 64 rows in one pinned host array and streams
    # them through the host-buffer call; tools/bench_stream.py streams chunks through predict_stream)
    "c5": dict(kind="et", n_estimators=1000, min_samples_leaf=5, n_train=1_000_000, n_test=10_000_000,
               n_features=64, quantiles=[5.0, 50.0, 95.0],
               label="ETQR T=1000 msl=5 bootstrap, 1M x 64 train, 10M streamed test rows, q={5,50,95}"),
    # reduced-size stand-ins used by tests and quick runs (NOT bench lines)
    "c2_small": dict(kind="rf", n_estimators=20, min_samples_leaf=1, n_train=20_000, n_test=20_000,
                     n_features=20, quantiles=[5.0, 50.0, 95.0], label="C2 at 1/5 scale"),
    "c3q": dict(kind="et", n_estimators=500, min_samples_leaf=5, n_train=250_000, n_test=250_000,
                n_features=50, quantiles=[5.0, 50.0, 95.0], label="C3 with a quarter of the rows (same T, leaf size)"),
    "c4q": dict(kind="rf", n_estimators=200, min_samples_leaf=1, n_train=250_000, n_test=250_000,
                n_features=20, quantiles=[float(q) for q in range(1, 100)],
                label="C4 with a quarter of the rows (RFQR T=200 msl=1, 99-point quantile grid)"),
    "c5q": dict(kind="et", n_estimators=1000, min_samples_leaf=5, n_train=250_000, n_test=1_000_000,
                n_features=64, quantiles=[5.0, 50.0, 95.0],
                label="C5 reduced (ETQR T=1000 msl=5, 250k x 64 train, 1M test rows per GPU through the host pipeline)"),
    "c3_small": dict(kind="et", n_estimators=50, min_samples_leaf=5, n_train=100_000, n_test=50_000,
                     n_features=50, quantiles=[5.0, 50.0, 95.0], label="C3 at 1/10 scale"),
}


def target(X, rng):
    """y = 3*x0 + 2*sin(3*x1) + x2*x3 + noise."""
    return 3 * X[:, 0] + 2 * np.sin(3 * X[:, 1]) + X[:, 2] * X[:, 3] + rng.randn(X.shape[0])


def make_train(n_train, n_features, seed=0, ties=False):
    rng = np.random.RandomState(seed)
    X = rng.randn(n_train, n_features).astype(np.float32)
    y = target(X, rng)
    if ties:
        y = np.round(y, 1)
    return X, y


def make_test(n_test, n_features, seed=1):
    """Test rows from the same distribution; ``seed`` differs per shard / rank."""
    rng = np.random.RandomState(1_000_003 + seed)
    return rng.randn(n_test, n_features).astype(np.float32)
