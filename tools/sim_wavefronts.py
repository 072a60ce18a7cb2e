#!/usr/bin/env python
"""CPU simulation of the L1TEX wavefronts of the leaf-lookup kernel (no GPU needed): fits
a small forest with scikit-learn, walks 32-row warps through it in lock step and counts
the distinct 128-byte lines per warp-level node load, with the rows in arrival order and
in the locality order (sorted by the leaf reached in tree 0).

    python tools/sim_wavefronts.py rf 100000 20      # kind, n_train, n_features
"""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from sklearn.ensemble import RandomForestRegressor, ExtraTreesRegressor
from skgarden_b200.datasets import make_train, make_test

kind = sys.argv[1] if len(sys.argv) > 1 else "rf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
F = int(sys.argv[3]) if len(sys.argv) > 3 else 20
T = 8
n = 32768
Xtr, ytr = make_train(N, F, seed=0)
Xte = make_test(n, F, seed=0)
t0 = time.time()
if kind == "rf":
    est = RandomForestRegressor(n_estimators=T, min_samples_leaf=1, bootstrap=True, random_state=0, n_jobs=8, max_features=1.0).fit(Xtr, ytr)
else:
    est = ExtraTreesRegressor(n_estimators=T, min_samples_leaf=5, bootstrap=True, random_state=0, n_jobs=8, max_features=1.0).fit(Xtr, ytr)
print("fit", time.time() - t0, "nodes", [e.tree_.node_count for e in est.estimators_][:3])

def walk(tree, X):
    """returns list per level of node index (sklearn ids, preorder) for active rows (-1 inactive)"""
    left, right, feat, thr = tree.children_left, tree.children_right, tree.feature, tree.threshold
    node = np.zeros(X.shape[0], dtype=np.int64)
    levels = []
    active = np.ones(X.shape[0], bool)
    while active.any():
        levels.append(np.where(active, node, -1))
        r = np.flatnonzero(active)
        nd = node[r]
        isleaf = left[nd] == -1
        active[r[isleaf]] = False
        r = r[~isleaf]; nd = nd[~isleaf]
        go_left = X[r, feat[nd]] <= thr[nd]
        node[r] = np.where(go_left, left[nd], right[nd])
    return levels, node

def wavefronts(levels, bytes_per_node=8, line=128):
    tot = 0; visits = 0; per_level = []
    for lv in levels:
        a = lv.reshape(-1, 32)
        lines = np.where(a >= 0, a * bytes_per_node // line, -1)
        s = np.sort(lines, axis=1)
        distinct = (s[:, 1:] != s[:, :-1]).sum(axis=1) + 1
        distinct -= (s[:, 0] == -1)   # the -1 group is not a load
        tot += distinct.sum(); visits += (a >= 0).sum()
        per_level.append((distinct.sum(), (a >= 0).sum()))
    return tot, visits, per_level

# key for sorting: leaf of tree 0
lv0, leaf0 = walk(est.estimators_[0].tree_, Xte)
order = np.argsort(leaf0, kind="stable")
for name, X in (("unsorted", Xte), ("sorted-by-tree0-leaf", Xte[order])):
    tw = tv = 0
    agg = {}
    for e in est.estimators_[1:]:
        levels, _ = walk(e.tree_, X)
        w, v, pl = wavefronts(levels)
        tw += w; tv += v
        for d, (a, b) in enumerate(pl):
            x = agg.setdefault(d, [0, 0]); x[0] += a; x[1] += b
    print(name, "wavefronts/visit = %.3f" % (tw / tv), "visits/row/tree = %.1f" % (tv / n / (T - 1)))
    print("  per level wf/visit:", " ".join("%d:%.2f" % (d, agg[d][0] / max(1, agg[d][1])) for d in sorted(agg) if d % 2 == 0))
