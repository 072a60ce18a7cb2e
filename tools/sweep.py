#!/usr/bin/env python
"""Kernel tuning sweep on the GPU box: times the leaf-lookup and weight/percentile kernels
of one config under different launch options (CUDA events inside the library, L2 flushed
between repetitions).  Not a bench line -- bench.py is the contract."""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--forest-cache", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--what", default="apply,quantile")
    ap.add_argument("--top-levels", default="10")
    args = ap.parse_args()
    import torch
    import skgarden_b200 as sg
    from skgarden_b200.datasets import CONFIGS, make_test
    from skgarden_b200.flatten import FlatForest
    sys.path.insert(0, ROOT)
    import bench
    cfg = CONFIGS[args.config]
    if args.forest_cache and os.path.isfile(args.forest_cache):
        flat = FlatForest.load(args.forest_cache)
    else:
        est, _, _ = bench.fit_forest(cfg)
        flat = est._flat
        if args.forest_cache:
            flat.save(args.forest_cache)
    n, F, qs = cfg["n_test"], cfg["n_features"], cfg["quantiles"]
    X = torch.from_numpy(make_test(n, F, seed=0)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    results = []

    def time_opts(dev, opts):
        for k, v in opts.items():
            dev.set_option(k, v)
        dev.set_option("profile", 1)
        a_ms, q_ms = [], []
        out = None
        for _ in range(args.reps + 1):
            flush.zero_()
            out = dev.predict_quantiles(X, qs, out=out)
            a, q = dev.last_timings()
            a_ms.append(a)
            q_ms.append(q)
        return float(np.median(a_ms[1:])), float(np.median(q_ms[1:])), out

    ref_out = None
    for L in [int(x) for x in args.top_levels.split(",")]:
        flat.build_top_blocks(L)
        dev = sg.DeviceForest(flat, device=0)
        if "apply" in args.what:
            grid = [dict(apply_variant=1, apply_ilp=i, apply_row_warps=rw) for i in (2, 4) for rw in (2, 8)]
            grid += [dict(apply_variant=2, apply_ilp=i, apply_rows_per_cta=r, apply_chunk_log2=c)
                     for i in (1, 2, 4, 8) for r in (128, 256) for c in (1, 2, 3)]
            for opts in grid:
                try:
                    a, q, out = time_opts(dev, opts)
                except Exception as exc:  # shape does not fit
                    print("L=%d %s -> %s" % (L, opts, str(exc)[:60]))
                    continue
                if ref_out is None:
                    ref_out = out.clone()
                ok = bool(torch.equal(out, ref_out))
                print("L=%2d apply %-80s apply_ms=%.4f same=%s" % (L, json.dumps(opts), a, ok), flush=True)
                results.append(dict(L=L, opts=opts, apply_ms=a, same=ok))
            dev.set_option("apply_variant", 0); dev.set_option("apply_rows_per_cta", 0)
            dev.set_option("apply_chunk_log2", -1); dev.set_option("apply_ilp", 4)
        if "quantile" in args.what:
            for we in (-1, 0, 4, 8):
                a, q, out = time_opts(dev, dict(warp_elems=we))
                if ref_out is None:
                    ref_out = out.clone()
                print("L=%2d quantile warp_elems=%2d quantile_ms=%.4f same=%s" % (
                    L, we, q, bool(torch.equal(out, ref_out))), flush=True)
        dev.close()
    best = sorted([r for r in results if r["same"]], key=lambda r: r["apply_ms"])[:5]
    print("BEST apply:", json.dumps(best))


if __name__ == "__main__":
    main()
