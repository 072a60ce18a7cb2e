#!/usr/bin/env python
"""Generates tests/golden_std/*.npz: outputs of the UNMODIFIED reference's ``return_std``
forests (skgarden/forest.py:181-548) on small synthetic problems.  Runs only in the build
container (needs /root/reference; imports it through oracle/reference_shim.py).

The reference spells the criterion 'mse' and checks for exactly that string before returning
the std (forest.py:355-358); scikit-learn 1.9 only accepts 'squared_error'.  The forests are
therefore fitted with 'squared_error' and the attribute is set to 'mse' before predict --
no reference file is edited.

    python tests/golden_std/make_golden_std.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import reference_shim as rs  # noqa: E402
from make_golden import boston_shaped, synth  # noqa: E402


def dump(name, est, X_train, y_train, X_test, min_variances):
    trees = [e.tree_ for e in est.estimators_]
    off = np.cumsum([0] + [t.node_count for t in trees])
    est.criterion = "mse"
    means, stds = [], []
    for mv in min_variances:
        est.min_variance = mv
        m, s = est.predict(X_test, return_std=True)
        means.append(m)
        stds.append(s)
    out = dict(
        X_train=np.asarray(X_train, dtype=np.float32), y_train=np.asarray(y_train, dtype=np.float64),
        X_test=np.asarray(X_test, dtype=np.float32), node_offsets=off,
        children_left=np.concatenate([t.children_left for t in trees]),
        children_right=np.concatenate([t.children_right for t in trees]),
        feature=np.concatenate([t.feature for t in trees]),
        threshold=np.concatenate([t.threshold for t in trees]),
        value=np.concatenate([t.value[:, 0, 0] for t in trees]),
        impurity=np.concatenate([t.impurity for t in trees]),
        apply=est.apply(X_test), min_variance=np.array(min_variances, dtype=np.float64),
        mean=np.array(means), std=np.array(stds),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-24s T=%3d N=%5d n=%4d nodes=%6d  %7.1f KB" % (name, len(trees), len(y_train), len(X_test), off[-1],
                                                          os.path.getsize(path) / 1024))


def main():
    rs.load()
    fm = sys.modules["skgarden.forest"]
    kw = dict(criterion="squared_error", max_features=1.0)
    Xb, yb = boston_shaped()
    perm = np.random.RandomState(0).permutation(506)
    tr, te = perm[:303], perm[303:]
    Xtr, Xte = Xb[tr].astype(np.float32), Xb[te].astype(np.float32)
    # n_jobs=1: the reference's mean is a threaded sum whose order is fixed only with one job
    dump("rf_boston", fm.RandomForestRegressor(n_estimators=10, random_state=0, n_jobs=1, **kw).fit(Xtr, yb[tr]),
         Xtr, yb[tr], Xte, [0.0, 0.5])
    dump("et_boston", fm.ExtraTreesRegressor(n_estimators=10, random_state=0, n_jobs=1, **kw).fit(Xtr, yb[tr]),
         Xtr, yb[tr], Xte, [0.0, 2.0])
    Xtr, ytr, Xte = synth(2000, 200, 6, seed=5)
    dump("rf_msl10", fm.RandomForestRegressor(n_estimators=12, random_state=1, min_samples_leaf=10, n_jobs=1, **kw).fit(Xtr, ytr),
         Xtr, ytr, Xte, [0.0, 0.01])
    dump("et_depth3", fm.ExtraTreesRegressor(n_estimators=7, random_state=2, max_depth=3, n_jobs=1, **kw).fit(Xtr, ytr),
         Xtr, ytr, Xte, [0.0])


if __name__ == "__main__":
    main()
