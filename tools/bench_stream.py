#!/usr/bin/env python
"""Streaming / multi-device throughput through the PRODUCT API (not the bench contract):

    python tools/bench_stream.py --config c5 --devices 0,1 --chunk-rows 1000000 --rows 10000000

Fits (or loads from bench.py's /tmp cache) the config's forest, then
  * est.predict(X, quantile=qs) with devices=[...]   (one call, rows sharded over the GPUs)
  * est.predict_stream(chunks, qs) with devices=[...] (chunks generated on the fly, staged into
    page-locked buffers by two host threads per GPU while the previous chunk is in flight)
and prints one JSON line with rows*quantiles/s, the host-feed rate in GB/s and a parity check of
the sharded result against a single-device call on a sample of rows."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c5")
    ap.add_argument("--devices", default="0")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--chunk-rows", type=int, default=1_000_000)
    ap.add_argument("--mode", default="fast")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import bench
    import skgarden_b200 as sg
    from skgarden_b200.datasets import CONFIGS, make_test
    cfg = CONFIGS[args.config]
    qs = cfg["quantiles"]
    n = args.rows or cfg["n_test"]
    F = cfg["n_features"]
    devices = [int(d) for d in args.devices.split(",")]
    wl = bench.load_workload(args.config, cfg)
    flat = wl["flat"]

    def shell(devs):
        est = sg.ExtraTreesQuantileRegressor(mode=args.mode, devices=devs)
        est._set_flat(flat)
        est.n_features_ = est.n_features_in_ = flat.n_features
        return est

    est = shell(devices)
    est._on_device()
    n_chunks = -(-n // args.chunk_rows)
    base = make_test(min(n, args.chunk_rows), F, seed=0)          # one block of rows, re-used with a per-chunk shift

    # a few distinct chunks, generated before the clock starts and cycled through: the stream is fed from
    # host memory that already exists (what a data loader hands over), pageable, not pinned
    pool = [np.ascontiguousarray(np.roll(base, 7919 * k, axis=0)) for k in range(min(4, n_chunks))]

    def chunks():
        for k in range(n_chunks):
            m = min(args.chunk_rows, n - k * args.chunk_rows)
            yield pool[k % len(pool)][:m]

    out = {"config": args.config, "devices": devices, "rows": n, "chunk_rows": args.chunk_rows, "mode": args.mode,
           "quantiles": qs, "forest_bytes_per_gpu": est._device_forest.device_bytes // len(devices)}
    # warm-up (pinned buffers, compact lists, workspaces)
    for _ in est.predict_stream(list(chunks())[:min(2 * len(devices), n_chunks)], qs):
        pass
    best = None
    for _ in range(args.reps):
        t0 = time.perf_counter()
        rows_done = 0
        first = None
        for res in est.predict_stream(chunks(), qs):
            rows_done += res.shape[0]
            if first is None:
                first = res
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    out["stream"] = {"seconds": round(best, 3), "rows_quantiles_per_s": round(n * len(qs) / best, 1),
                     "host_feed_GBps": round(n * F * 4 / best / 1e9, 2),
                     "api": "est.predict_stream(chunks, quantile=[...]) with devices=%r" % devices}
    # one sharded call over an in-memory array (bounded to 4M rows of host memory)
    m = min(n, 4_000_000)
    X = np.concatenate([pool[k % len(pool)] for k in range(-(-m // base.shape[0]))])[:m]
    est.predict(X[:4096], quantile=qs)
    t0 = time.perf_counter()
    got = est.predict(X, quantile=qs)
    dt = time.perf_counter() - t0
    out["predict"] = {"rows": m, "seconds": round(dt, 3), "rows_quantiles_per_s": round(m * len(qs) / dt, 1),
                      "api": "est.predict(X, quantile=[...]) with devices=%r (pageable X, check_array included)" % devices}
    single = shell(None)
    idx = np.random.RandomState(0).choice(m, 4096, replace=False)
    want = single.predict(X[idx], quantile=qs)
    out["sharded_equals_single_device"] = bool(np.array_equal(got[idx], want))
    out["stream_first_chunk_equals_predict"] = bool(np.array_equal(first[:1024], est.predict(base[:1024], quantile=qs)))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
