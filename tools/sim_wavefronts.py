#!/usr/bin/env python
"""CPU simulation of the L1TEX cost of the leaf-lookup kernel (no GPU needed): fits a small
forest with scikit-learn, walks 32-row warps through it in lock step and counts the distinct
32-byte sectors / 128-byte lines per warp-level node load

* with the rows in arrival order and in the locality order (sorted by the leaf reached in tree 0),
* for several node layouts: depth-first preorder, level order with sibling pairs (the layout of
  include/qforest_b200.h), sibling pairs in depth-first order, and level order for the first
  levels followed by depth-first subtrees.

    python tools/sim_wavefronts.py rf 100000 20 100000      # kind, n_train, n_features, n_test
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor

from skgarden_b200.datasets import make_test, make_train

kind = sys.argv[1] if len(sys.argv) > 1 else "rf"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
F = int(sys.argv[3]) if len(sys.argv) > 3 else 20
n = int(sys.argv[4]) if len(sys.argv) > 4 else 32768
T = 6
Xtr, ytr = make_train(N, F, seed=0)
Xte = make_test(n, F, seed=0)
t0 = time.time()
if kind == "rf":
    est = RandomForestRegressor(n_estimators=T, min_samples_leaf=1, bootstrap=True, random_state=0,
                                n_jobs=8, max_features=1.0).fit(Xtr, ytr)
else:
    est = ExtraTreesRegressor(n_estimators=T, min_samples_leaf=5, bootstrap=True, random_state=0,
                              n_jobs=8, max_features=1.0).fit(Xtr, ytr)
print("fit %.1f s, nodes per tree" % (time.time() - t0), [e.tree_.node_count for e in est.estimators_][:3], flush=True)


def walk(tree, X):
    """per level: the node (scikit-learn id) every row sits on, -1 once it reached its leaf"""
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
        r, nd = r[~isleaf], nd[~isleaf]
        node[r] = np.where(X[r, feat[nd]] <= thr[nd], left[nd], right[nd])
    return levels, node


def layout_preorder(tree):
    return np.arange(tree.node_count)          # the depth-first builder numbers in preorder


def layout_pairs(tree, level_order_depth):
    """slot of every node: root 0, padding 1, sibling pairs; level order for the first
    `level_order_depth` levels, every subtree below in depth-first order"""
    left, right = tree.children_left, tree.children_right
    pos = np.zeros(tree.node_count, np.int64)
    nxt, frontier, d = 2, [0], 0
    while frontier and d < level_order_depth:
        new = []
        for u in frontier:
            if left[u] != -1:
                pos[left[u]], pos[right[u]] = nxt, nxt + 1
                nxt += 2
                new += [left[u], right[u]]
        frontier, d = new, d + 1
    for root in frontier:
        stack = [root]
        while stack:
            u = stack.pop()
            if left[u] != -1:
                pos[left[u]], pos[right[u]] = nxt, nxt + 1
                nxt += 2
                stack += [right[u], left[u]]
    return pos


LAYOUTS = [("preorder", layout_preorder),
           ("level order, pairs", lambda t: layout_pairs(t, 1 << 30)),
           ("pairs, depth-first", lambda t: layout_pairs(t, 0)),
           ("12 levels, then depth-first", lambda t: layout_pairs(t, 12))]


def count(levels, pos, unit):
    """distinct `unit`-byte blocks over all warp loads, and the number of warp loads"""
    tot = req = 0
    for lv in levels:
        a = lv.reshape(-1, 32)
        s = np.sort(np.where(a >= 0, pos[np.maximum(a, 0)] * 8 // unit, -1), axis=1)
        distinct = (s[:, 1:] != s[:, :-1]).sum(axis=1) + 1 - (s[:, 0] == -1)
        tot += distinct.sum()
        req += (distinct > 0).sum()
    return tot, req


_, leaf0 = walk(est.estimators_[0].tree_, Xte)
order = np.argsort(leaf0, kind="stable")
rows = (n // 32) * 32
for oname, X in (("arrival order", Xte[:rows]), ("locality order (leaf of tree 0)", Xte[order][:rows])):
    res = {}
    for e in est.estimators_[1:]:
        levels, _ = walk(e.tree_, X)
        for lname, fn in LAYOUTS:
            pos = fn(e.tree_)
            for unit in (32, 128):
                t, r = count(levels, pos, unit)
                x = res.setdefault((lname, unit), [0, 0])
                x[0] += t
                x[1] += r
    print(oname)
    for lname, _ in LAYOUTS:
        s, l = res[(lname, 32)], res[(lname, 128)]
        print("  %-28s sectors per warp load %5.2f   lines %5.2f   sectors per row and tree %5.2f"
              % (lname, s[0] / s[1], l[0] / l[1], s[0] / rows / (T - 1)))
