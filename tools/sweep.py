#!/usr/bin/env python
"""Kernel tuning sweep on the GPU box: times the leaf-lookup and weight/percentile kernels
of one config under different launch options (CUDA events inside the library, L2 flushed
between repetitions).  Not a bench line -- bench.py is the contract."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--forest-cache", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rows", type=int, default=0, help="test rows (default: the config's)")
    ap.add_argument("--grid", default="default", help="default | json list of option dicts")
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    args = ap.parse_args()
    import torch
    import skgarden_b200 as sg
    from skgarden_b200.datasets import CONFIGS, make_test
    from skgarden_b200.flatten import FlatForest
    import bench
    cfg = CONFIGS[args.config]
    if args.forest_cache and os.path.isfile(args.forest_cache):
        flat = FlatForest.load(args.forest_cache)
    else:
        est, _, _ = bench.fit_forest(cfg)
        flat = est._flat
        if args.forest_cache:
            flat.save(args.forest_cache)
    n, F, qs = args.rows or cfg["n_test"], cfg["n_features"], cfg["quantiles"]
    print("forest: T=%d nodes=%d members=%d max_depth=%d expected_row_members=%.1f max_row_members=%d" % (
        flat.n_trees, flat.n_nodes, flat.n_members, flat.max_depth, flat.expected_row_members,
        flat.max_row_members), flush=True)
    X = torch.from_numpy(make_test(n, F, seed=0)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dev = sg.DeviceForest(flat, device=0)
    defaults = dict(select_impl=0, sort_rows=-1, apply_row_warps=0, apply_ilp=4, warp_elems=-1, rows_per_chunk=262144,
                    cta_capacity=4096)
    if args.grid == "default":
        grid = [dict(sort_rows=s, apply_ilp=i, apply_row_warps=rw)
                for s in (0, 1) for i in (2, 4, 8) for rw in (1, 2, 4)]
        grid += [dict(rows_per_chunk=c) for c in (32768, 65536, 262144)]
        grid += [dict(warp_elems=w) for w in (0, 4, 8)]
    else:
        grid = json.loads(args.grid)
    from skgarden_b200 import _native
    mode = _native.QF_MODE_FAST if args.mode == "fast" else _native.QF_MODE_EXACT
    ref_out = None
    results = []
    for opts in grid:
        full = dict(defaults)
        full.update(opts)
        for k, v in full.items():
            dev.set_option(k, v)
        dev.set_option("profile", 1)
        a_ms, q_ms, t_ms = [], [], []
        out = None
        try:
            for _ in range(args.reps + 1):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = dev.predict_quantiles(X, qs, mode=mode, out=out)
                e1.record()
                a, q = dev.last_timings()
                torch.cuda.synchronize()
                a_ms.append(a)
                q_ms.append(q)
            dev.set_option("profile", 0)              # whole-call time without the per-kernel events
            t_ms = []
            for _ in range(args.reps + 1):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = dev.predict_quantiles(X, qs, mode=mode, out=out)
                e1.record()
                torch.cuda.synchronize()
                t_ms.append(e0.elapsed_time(e1))
        except Exception as exc:
            print("%-70s -> %s" % (json.dumps(opts), str(exc)[:80]), flush=True)
            continue
        if ref_out is None:
            ref_out = out.clone()
            if args.mode == "fast":                      # deviation of the selection kernel from the sorting one
                for k, v in full.items():
                    dev.set_option(k, v)
                ex = dev.predict_quantiles(X, qs, mode=_native.QF_MODE_EXACT)
                torch.cuda.synchronize()
                diff = (ex - ref_out).abs()
                print("fast vs exact: max abs diff %.3g, max rel diff %.3g, |value| max %.3g" % (
                    float(diff.max()), float((diff / ex.abs().clamp_min(1e-3)).max()), float(ex.abs().max())), flush=True)
        ok = bool(torch.equal(out, ref_out))
        r = dict(opts=opts, apply_ms=float(np.median(a_ms[1:])), quantile_ms=float(np.median(q_ms[1:])),
                 total_ms=float(np.median(t_ms[1:])), same=ok)
        results.append(r)
        print("%-70s apply=%.4f quantile=%.4f total=%.4f ms same=%s" % (
            json.dumps(opts), r["apply_ms"], r["quantile_ms"], r["total_ms"], ok), flush=True)
    dev.close()
    best = sorted([r for r in results if r["same"]], key=lambda r: r["total_ms"])[:3]
    print("BEST:", json.dumps(best))


if __name__ == "__main__":
    main()
