#!/usr/bin/env python
"""Benchmark of the quantile-forest predict path (BASELINE.json metric:
test rows x quantiles / sec; % of HBM roofline; reference CPU predict beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (leaf lookup + fused weight/percentile kernels) over
one batch of synthetic test rows.  Prints ONE JSON line (rank 0).

  value     : rows x quantiles / s, all ranks, inputs resident in HBM, CUDA-event timed,
              L2 flushed between timed steps
  e2e       : same metric through the C-ABI host-buffer call (pinned host X in, host result
              out; H2D and D2H inside the timed region)
  roofline  : HBM roofline of the dominant kernel (by measured time) + both kernels listed
  cpu_baseline : the reference's algorithm (dense O(T*N)/row, ensemble.py:137-142) on the
              host cores on a bounded row sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "predict throughput (test rows x quantiles per second)"
UNIT = "rows*quantiles/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def fit_forest(cfg, product=True, n_jobs=-1):
    """Fit the config's forest on synthetic data.  product=True -> our estimator (fit +
    flatten); product=False -> plain scikit-learn (what the reference's fit is)."""
    from skgarden_b200.datasets import make_train
    Xtr, ytr = make_train(cfg["n_train"], cfg["n_features"], seed=0)
    kw = dict(n_estimators=cfg["n_estimators"], min_samples_leaf=cfg["min_samples_leaf"],
              bootstrap=True, random_state=0, n_jobs=n_jobs)
    t0 = time.time()
    if product:
        import skgarden_b200 as sg
        cls = sg.RandomForestQuantileRegressor if cfg["kind"] == "rf" else sg.ExtraTreesQuantileRegressor
        est = cls(**kw).fit(Xtr, ytr)
    else:
        from sklearn.ensemble import ExtraTreesRegressor, RandomForestRegressor
        cls = RandomForestRegressor if cfg["kind"] == "rf" else ExtraTreesRegressor
        est = cls(criterion="squared_error", max_features=1.0, **kw).fit(Xtr, ytr)
    log("[bench] fit %s T=%d N=%d F=%d: %.1f s" % (cfg["kind"], cfg["n_estimators"], cfg["n_train"],
                                                    cfg["n_features"], time.time() - t0))
    return est, Xtr, ytr


# --------------------------------------------------------------------------------------------
# CPU baseline: the reference's per-row algorithm on host cores (bounded sample)
# --------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    rows, qs = args
    from oracle import qrf_oracle as O
    t0 = time.perf_counter()
    for q in qs:                                      # scalar-q API: one call per quantile
        O.predict_dense(_CPU["y"], _CPU["leaves"], _CPU["weights"], _CPU["X_leaves"][rows], q, _CPU["sorter"])
    return time.perf_counter() - t0


def cpu_reference_setup(estimators, Xtr, ytr, Xte_sample, bootstrap=True):
    """Dense fit attributes (ensemble.py:83-101, O(T*N) bincount form) + leaf ids of the
    sampled test rows via scikit-learn's own apply (what forest.py:601 calls)."""
    from oracle import qrf_oracle as O
    leaves, weights = O.fit_attributes(estimators, Xtr, bootstrap)
    X_leaves = np.array([e.tree_.apply(Xte_sample) for e in estimators]).T
    _CPU.update(y=ytr, leaves=leaves, weights=weights, X_leaves=X_leaves, sorter=np.argsort(ytr))


def cpu_reference_throughput(qs, rows_per_worker, workers):
    """rows*quantiles/s of the dense reference algorithm with `workers` forked processes
    on disjoint row chunks (the per-row loop itself is single-threaded by construction,
    ensemble.py:137)."""
    import multiprocessing as mp
    n = _CPU["X_leaves"].shape[0]
    chunks = [np.arange(w * rows_per_worker, (w + 1) * rows_per_worker) % n for w in range(workers)]
    t0 = time.perf_counter()
    if workers == 1:
        _cpu_worker((chunks[0], qs))
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            pool.map(_cpu_worker, [(c, qs) for c in chunks])
    dt = time.perf_counter() - t0
    return workers * rows_per_worker * len(qs) / dt, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every few ms while the GPU is
    busy (background thread; falls back to one nvidia-smi query if NVML is missing)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
               "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as exc:
            log("[bench] NVML unavailable (%s)" % exc)

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.003)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def row_member_counts(flat, leaf_ids):
    """Members gathered per row (P_row), exact, from the leaf ids actually hit."""
    hi = (flat.nodes >> np.uint64(32)).astype(np.uint32)
    cnt = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7FFFFFFF), 0).astype(np.int64)
    if flat.node_ids is None:
        packed = leaf_ids + flat.tree_offsets[:-1][None, :]
    else:
        packed = np.empty_like(leaf_ids)
        for t in range(flat.n_trees):
            n0, n1 = flat.tree_offsets[t], flat.tree_offsets[t + 1]
            ids = flat.node_ids[n0:n1]
            inv = np.empty(n1 - n0, dtype=np.int64)        # sklearn id -> packed slot
            inv[ids[ids >= 0]] = n0 + np.flatnonzero(ids >= 0)
            packed[:, t] = inv[leaf_ids[:, t]]
    return cnt[packed].sum(axis=1)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(config):
    """DRAM bytes per launch of each kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        try:
            return {k: v for k, v in json.load(open(path)).get(config, {}).items() if not k.startswith("_")}
        except Exception:
            pass
    return {}


def run_b200(args, rank, world, local_rank):
    from skgarden_b200.datasets import CONFIGS, make_test
    cfg = CONFIGS[args.config]
    qs = cfg["quantiles"]
    n, F, T = cfg["n_test"], cfg["n_features"], cfg["n_estimators"]

    # ---- forest (rank 0 fits; CPU baseline before any CUDA work so fork is safe) ----
    flat, cpu_baseline = None, None
    if rank == 0:
        if args.forest_cache and os.path.isfile(args.forest_cache):
            from skgarden_b200.flatten import FlatForest
            flat = FlatForest.load(args.forest_cache)          # profiling runs: skip the refit
            log("[bench] loaded flattened forest from", args.forest_cache)
        else:
            est, Xtr, ytr = fit_forest(cfg, product=True)
            flat = est._flat
            if args.forest_cache:
                flat.save(args.forest_cache)
            if world == 1 and not args.no_cpu_baseline:
                cpu_baseline = measure_cpu_baseline(est.estimators_, Xtr, ytr, cfg, budget_s=args.cpu_seconds)
            del Xtr
    import torch
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        from skgarden_b200.sharding import broadcast_flat
        flat = broadcast_flat(flat, rank, device)
    import skgarden_b200 as sg
    from skgarden_b200 import _native
    dev = sg.DeviceForest(flat, device=local_rank)
    mode = _native.QF_MODE_FAST if args.mode == "fast" else _native.QF_MODE_EXACT

    # ---- this rank's shard of test rows (weak scaling: n rows per GPU) ----
    Xh = torch.from_numpy(make_test(n, F, seed=rank)).pin_memory()
    Xd = Xh.to(device)
    outd = torch.empty((n, len(qs)), dtype=torch.float64, device=device)
    outh = torch.empty((n, len(qs)), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_steps(k, profile=False):
        """k steps, each bracketed by CUDA events on the launching (current) stream; L2 is
        flushed between steps outside the events.  Returns (sum ms, apply ms, quantile ms)."""
        dev.set_option("profile", 1 if profile else 0)
        dev._workspace(n, len(qs))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        a_ms = q_ms = 0.0
        for s, e in evs:
            flush.zero_()
            s.record()
            dev.predict_quantiles(Xd, qs, mode=mode, out=outd)
            e.record()
            if profile:
                a, q = dev.last_timings()
                a_ms += a
                q_ms += q
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs), a_ms, q_ms

    def host_steps(k):
        """k end-to-end steps through the host-buffer C-ABI call (H2D + kernels + D2H)."""
        total = 0.0
        for _ in range(k):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev.predict_quantiles_host(Xh.numpy(), qs, out_host=outh.numpy(), mode=mode)
            total += time.perf_counter() - t0
        return total * 1e3

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    device_steps(args.warmup)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    total_ms, _, _ = device_steps(args.steps)
    barrier()
    launches_per_step = dev.last_launches
    total_ms = max_over_ranks(total_ms)
    _, apply_ms, quant_ms = device_steps(args.steps, profile=True)
    barrier()
    host_steps(min(args.warmup, 2))
    barrier()
    e2e_ms = max_over_ranks(host_steps(args.steps))
    clock_info = clocks.stop() if rank == 0 else None
    barrier()

    # ---- accounting ----
    ms_per_step = total_ms / args.steps
    value = world * n * len(qs) / (ms_per_step * 1e-3)
    e2e_value = world * n * len(qs) / (e2e_ms / args.steps * 1e-3)
    if rank == 0:
        leaf_ids = dev.apply(Xd).cpu().numpy()
        P_rows = row_member_counts(flat, leaf_ids)
        peak, peak_src = load_peaks()
        traffic = load_traffic(args.config)
        Q = len(qs)
        bytes_k2 = float(8 * T * n + 8 * P_rows.sum() + 16 * Q * n)        # K2 bytes/row
        bytes_k1 = float(4 * n * F + 8 * n * T + 8 * flat.n_nodes)         # K1 bytes/launch
        k1_ms, k2_ms = apply_ms / args.steps, quant_ms / args.steps

        def roof(name, nbytes, ms):
            ach = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(name), "peak_source": peak_src,
                    "algorithmic_bytes": nbytes, "ms_per_launch": round(ms, 4)}

        plan = dev.plan(n)                                                  # which kernels the library ran
        if plan["warp_elems"]:
            k2_name = "quantile_warp_kernel<%d>" % plan["warp_elems"]
        elif args.mode == "fast" and Q <= 16:
            nbits = max(1, int(flat.n_train - 1).bit_length())
            k2_name = ("quantile_select2_kernel<%d>" % (11 if nbits <= 20 else 12)) if (
                nbits <= 21 and T <= 1024) else ("quantile_select_kernel<%d>" % plan["s1_threads"])
        else:
            k2_name = "quantile_bucket_kernel<%d>" % plan["s1_threads"]
        r1 = roof("apply_kernel", bytes_k1, k1_ms)
        r2 = roof(k2_name, bytes_k2, k2_ms)
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, cfg["label"]), "rows_per_gpu": n,
                       "quantiles": qs, "mode": args.mode, "l2": "flushed between timed steps (256 MiB write)",
                       "forest_bytes": dev.device_bytes, "mean_row_members": float(P_rows.mean()),
                       "sharding": "test rows sharded over GPUs, forest replicated, no data-path collective"},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": n * F * 4,
                    "d2h_bytes_per_step": n * Q * 8, "ms_per_step": round(e2e_ms / args.steps, 4),
                    "api": "qf_forest_predict_quantiles_host (pinned host X -> host result)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": r1 if k1_ms >= k2_ms else r2,
            "roofline_kernels": {"apply_kernel": r1, k2_name: r2},
            "kernel_ms": {"apply": round(k1_ms, 4), "quantile": round(k2_ms, 4),
                          "locality_order_and_gaps": round(max(0.0, ms_per_step - k1_ms - k2_ms), 4)},
            "clocks": clock_info,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    dev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _memory_bound_workers(cfg):
    """The dense algorithm materialises a (T,N) bool mask and a masked float32 copy per row
    (ensemble.py:138-140): ~6 bytes x T x N transient per worker on top of the shared 8 x T x N
    fit attributes.  Keep the workers inside half of the available host memory."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        return 4
    tn = cfg["n_estimators"] * cfg["n_train"]
    return max(1, int((avail * 0.5 - 8 * tn) // (6 * tn)))


def measure_cpu_baseline(estimators, Xtr, ytr, cfg, budget_s=12.0, workers=None):
    """Dense reference algorithm on a bounded row sample, all host cores."""
    from skgarden_b200.datasets import make_test
    workers = workers or max(1, min(host_cores(), 64))
    workers = min(workers, _memory_bound_workers(cfg))
    qs = cfg["quantiles"][:3]
    sample = make_test(min(cfg["n_test"], 4096), cfg["n_features"], seed=0)
    cpu_reference_setup(estimators, Xtr, ytr, sample)
    _, dt = cpu_reference_throughput(qs, 1, workers)              # calibrate: 1 row per worker
    rows = int(max(1, min(256, budget_s / max(dt, 1e-3))))
    thr, dt = cpu_reference_throughput(qs, rows, workers)
    return {"value": round(thr, 3), "unit": UNIT, "cores": workers, "kind": "port",
            "sample": "%d rows x %d quantiles (%d rows per worker, %d forked workers), %.1f s wall; "
                      "dense O(T*N)-per-row algorithm of ensemble.py:137-142 + utils.py:46-81 "
                      "(oracle.predict_dense, bit-identical to the reference)" % (
                          rows * workers, len(qs), rows, workers, dt)}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm, all host threads, same config/metric
# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from skgarden_b200.datasets import CONFIGS, make_test
    cfg = CONFIGS[args.config]
    qs = cfg["quantiles"][:3] if len(cfg["quantiles"]) > 3 else cfg["quantiles"]
    est, Xtr, ytr = fit_forest(cfg, product=False)
    workers = max(1, min(host_cores(), 64, _memory_bound_workers(cfg)))
    sample = make_test(min(cfg["n_test"], 4096), cfg["n_features"], seed=0)
    cpu_reference_setup(est.estimators_, Xtr, ytr, sample)
    _, dt1 = cpu_reference_throughput(qs, 1, workers)
    # bounded: the whole --steps/--warmup run should end within a few minutes
    rows = int(max(1, min(64, (120.0 / max(1, args.steps + args.warmup)) / max(dt1, 1e-3))))
    for _ in range(args.warmup):
        cpu_reference_throughput(qs, rows, workers)
    t_total = 0.0
    for _ in range(args.steps):
        _, dt = cpu_reference_throughput(qs, rows, workers)
        t_total += dt
    per_step = rows * workers * len(qs)
    value = per_step * args.steps / t_total
    sample_txt = ("each step = %d rows x %d quantiles (%d rows per worker, %d forked workers) of the same "
                  "workload; dense O(T*N)-per-row algorithm (ensemble.py:137-142 + utils.py:46-81) as "
                  "restated in oracle/qrf_oracle.py (bit-identical to the reference; the reference is "
                  "Python and cannot travel to the GPU box); leaf ids from scikit-learn's own apply"
                  % (rows * workers, len(qs), rows, workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_total / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.config, cfg["label"]), "rows_per_step": rows * workers,
                   "quantiles": qs},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": sample_txt},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--forest-cache", default=None,
                    help="npz path: load the flattened forest if present, else fit and save it")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
