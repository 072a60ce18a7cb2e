#!/usr/bin/env python
"""Fit-side bookkeeping: host flattener (tree_.apply on the CPU + qf_pack_tree) against the
device builder (qf_build_csr) on one of the bench configs.  Not a bench line."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    args = ap.parse_args()
    import torch
    import bench
    from skgarden_b200.datasets import CONFIGS
    cfg = CONFIGS[args.config]
    est, Xtr, ytr = bench.fit_forest(cfg, product=False)          # plain scikit-learn fit
    import skgarden_b200 as sg
    cls = sg.RandomForestQuantileRegressor if cfg["kind"] == "rf" else sg.ExtraTreesQuantileRegressor
    q = cls(n_estimators=cfg["n_estimators"], min_samples_leaf=cfg["min_samples_leaf"], bootstrap=True,
            random_state=0, n_jobs=-1)
    q.estimators_ = est.estimators_
    X32 = np.ascontiguousarray(Xtr, dtype=np.float32)
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    from concurrent.futures import ThreadPoolExecutor
    from skgarden_b200.flatten import bootstrap_counts, flatten_forest, flatten_forest_device
    trees = [e.tree_ for e in est.estimators_]
    N = len(ytr)
    t0 = time.perf_counter()
    counts = [bootstrap_counts(e.random_state, N) for e in est.estimators_]      # ensemble.py:88-94 (both builders)
    t_counts = time.perf_counter() - t0
    sorter = np.argsort(ytr)
    workers = os.cpu_count() or 1
    print("bootstrap multiplicities (shared by both builders): %.2f s" % t_counts, flush=True)

    def host():
        with ThreadPoolExecutor(workers) as ex:                                   # tree.py:101 on the CPU
            leaves = list(ex.map(lambda t: trees[t].apply(X32), range(len(trees))))
        return flatten_forest(trees, leaves, counts, ytr, X32.shape[1], sorter=sorter, n_jobs=workers)

    def device():
        return flatten_forest_device(trees, X32, counts, ytr, X32.shape[1], sorter=sorter, n_jobs=workers)

    res = {}
    for name, fn in (("host", host), ("device", device), ("device", device)):
        t0 = time.perf_counter()
        flat = fn()
        res[name] = (time.perf_counter() - t0, flat)
        print("%-6s builder: %.2f s  (nodes %d, members %d)" % (name, res[name][0], flat.n_nodes, flat.n_members), flush=True)
    same = np.array_equal(res["host"][1].nodes, res["device"][1].nodes) and \
        np.array_equal(res["host"][1].members, res["device"][1].members)
    print("identical arrays:", same, " speed-up: %.1fx" % (res["host"][0] / res["device"][0]))


if __name__ == "__main__":
    main()
