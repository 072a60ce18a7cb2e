#!/usr/bin/env python
"""Generates tests/golden_tree/*.npz: outputs of the UNMODIFIED reference's tree-level
estimators (skgarden/quantile/tree.py: DecisionTreeQuantileRegressor,
ExtraTreeQuantileRegressor) on small synthetic problems.  Runs only in the build container
(needs /root/reference; imports it through oracle/reference_shim.py).

    python tests/golden_tree/make_golden_tree.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from oracle import reference_shim as rs  # noqa: E402
from make_golden import QS, boston_shaped, synth  # noqa: E402


def dump(name, est, X_train, y_train, X_test):
    t = est.tree_
    out = dict(
        X_train=np.asarray(X_train, dtype=np.float32), y_train=np.asarray(y_train, dtype=np.float64),
        X_test=np.asarray(X_test, dtype=np.float32),
        children_left=t.children_left, children_right=t.children_right, feature=t.feature,
        threshold=t.threshold, value=t.value[:, 0, 0],
        y_train_leaves_=est.y_train_leaves_, apply=est.apply(X_test),
        q=np.array(QS, dtype=np.float64),
        pred=np.array([est.predict(X_test, quantile=(int(q) if float(q).is_integer() else q)) for q in QS]),
        mean=est.predict(X_test),
    )
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("%-28s N=%5d n=%4d nodes=%6d  %7.1f KB" % (name, len(y_train), len(X_test), t.node_count,
                                                    os.path.getsize(path) / 1024))


def main():
    ref = rs.load()
    DT, ET = ref.DecisionTreeQuantileRegressor, ref.ExtraTreeQuantileRegressor
    kw_dt = dict(criterion="squared_error")
    kw_et = dict(criterion="squared_error", max_features=1.0)

    # test_tree.py:9-19: Boston-shaped 60/40 split, default trees (pure leaves) and depth 1
    Xb, yb = boston_shaped()
    perm = np.random.RandomState(0).permutation(506)
    tr, te = perm[:303], perm[303:]
    Xtr, Xte = Xb[tr].astype(np.float32), Xb[te].astype(np.float32)
    dump("dt_boston_full", DT(random_state=0, **kw_dt).fit(Xtr, yb[tr]), Xtr, yb[tr], Xte)
    dump("et_boston_full", ET(random_state=0, **kw_et).fit(Xtr, yb[tr]), Xtr, yb[tr], Xte)
    dump("dt_boston_depth1", DT(random_state=0, max_depth=1, **kw_dt).fit(Xtr, yb[tr]), Xtr, yb[tr], Xte)
    dump("et_boston_depth3", ET(random_state=0, max_depth=3, **kw_et).fit(Xtr, yb[tr]), Xtr, yb[tr], Xte)

    # larger leaves, tie-heavy targets
    Xtr, ytr, Xte = synth(2500, 200, 6, seed=7, ties=True)
    dump("dt_msl20_ties", DT(random_state=1, min_samples_leaf=20, **kw_dt).fit(Xtr, ytr), Xtr, ytr, Xte)
    # best-first builder (node order is not depth-first)
    dump("et_maxleaf40", ET(random_state=2, max_leaf_nodes=40, **kw_et).fit(Xtr, ytr), Xtr, ytr, Xte)


if __name__ == "__main__":
    main()
