#!/usr/bin/env python
"""Times the host-buffer C-ABI call (pinned X in, host result out) for several pipeline
chunk sizes.  Tuning aid, not a bench line."""
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
    ap.add_argument("--forest-cache", default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--mode", default="exact")
    ap.add_argument("--chunks", default="0,100000,50000,33344,25000,16384")
    args = ap.parse_args()
    import torch
    import skgarden_b200 as sg
    from skgarden_b200 import _native
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
    n, F, qs = cfg["n_test"], cfg["n_features"], cfg["quantiles"]
    Xh = torch.from_numpy(make_test(n, F, seed=0)).pin_memory()
    outh = torch.empty((n, len(qs)), dtype=torch.float64).pin_memory()
    dev = sg.DeviceForest(flat, device=0)
    mode = _native.QF_MODE_FAST if args.mode == "fast" else _native.QF_MODE_EXACT
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for chunk in [int(c) for c in args.chunks.split(",")]:
        ts = []
        for _ in range(args.reps + 3):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev.predict_quantiles_host(Xh.numpy(), qs, out_host=outh.numpy(), mode=mode, chunk_rows=chunk)
            ts.append(time.perf_counter() - t0)
        print("chunk_rows=%-7d  median %.4f ms  min %.4f ms" % (chunk, np.median(ts[3:]) * 1e3, np.min(ts[3:]) * 1e3), flush=True)
    dev.close()


if __name__ == "__main__":
    main()
