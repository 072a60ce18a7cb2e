#!/usr/bin/env python
"""Benchmark of the quantile-forest predict path (BASELINE.json metric:
test rows x quantiles / sec; % of HBM roofline; reference CPU predict beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload: C3 = BASELINE.json configs[2], the 1M-row configuration the north_star
targets are quoted on (ETQR T=500 msl=5, 1M x 50 train, 1M test rows, q = {5, 50, 95}).
One "step" = one pass of the hot path (leaf lookup + fused weight/percentile kernels) over
the WHOLE test set; with N GPUs the rows of that ONE test set are split into N contiguous
shards (strong scaling), the forest is replicated and the (n, Q) result is gathered on
every rank inside the timed step.  Prints ONE JSON line (rank 0).

  value        : rows x quantiles / s, all ranks, inputs resident in HBM, CUDA-event timed,
                 L2 flushed between timed steps
  e2e          : same metric through the C-ABI host-buffer call (pinned host X in, host result
                 out; H2D and D2H inside the timed region)
  roofline     : HBM roofline of the dominant kernel (by measured time) + both kernels listed,
                 on the ALGORITHMIC bytes of SURVEY.md section 8(d)
  parity_sample: rows of the TIMED output compared with oracle.predict_sparse / scikit-learn's
                 own apply (the run fails when they differ)
  cpu_baseline : the reference's algorithm (dense O(T*N)/row, ensemble.py:137-142) on the
                 host cores on a bounded row sample; 1-core and sparse-restatement figures beside it
"""
import argparse
import hashlib
import json
import os
import platform
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
KERNEL_TAG = "r2"                # profiles/traffic.json entries made with another binary read as null
SAMPLE_ROWS = 512                # rows of the test set with scikit-learn leaf ids kept for parity / CPU legs
CACHE_DIR = os.environ.get("QF_BENCH_CACHE", "/tmp/qf_b200_cache")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------------
# workload: fitted + flattened forest (cached under /tmp so that back-to-back runs fit once)
# --------------------------------------------------------------------------------------------
def _cache_paths(name, cfg):
    from skgarden_b200 import _native
    from skgarden_b200.flatten import LAYOUT_VERSION
    import sklearn
    key = json.dumps([name, {k: cfg[k] for k in sorted(cfg) if k != "label"}, LAYOUT_VERSION,
                      _native.ABI_VERSION, sklearn.__version__, np.__version__, SAMPLE_ROWS], sort_keys=True)
    h = hashlib.sha1(key.encode()).hexdigest()[:12]
    base = os.path.join(CACHE_DIR, "%s-%s" % (name, h))
    return base + "-flat.npz", base + "-extra.npz"


def sample_row_ids(n_test):
    rng = np.random.RandomState(7)
    return np.sort(rng.choice(n_test, size=min(SAMPLE_ROWS, n_test), replace=False))


def fit_forest(cfg, n_jobs=-1):
    """Fit the config's forest on synthetic data with the product estimator (= scikit-learn's CPU
    forest fit, which is what the reference's fit is, plus one flattening pass)."""
    import skgarden_b200 as sg
    from skgarden_b200.datasets import make_train
    Xtr, ytr = make_train(cfg["n_train"], cfg["n_features"], seed=0)
    kw = dict(n_estimators=cfg["n_estimators"], min_samples_leaf=cfg["min_samples_leaf"],
              bootstrap=True, random_state=0, n_jobs=n_jobs)
    if cfg.get("max_depth") is not None:
        kw["max_depth"] = cfg["max_depth"]
    t0 = time.time()
    cls = sg.RandomForestQuantileRegressor if cfg["kind"] == "rf" else sg.ExtraTreesQuantileRegressor
    est = cls(**kw).fit(Xtr, ytr)
    log("[bench] fit + flatten %s T=%d N=%d F=%d: %.1f s" % (cfg["kind"], cfg["n_estimators"], cfg["n_train"],
                                                             cfg["n_features"], time.time() - t0))
    return est, Xtr, ytr


def load_workload(name, cfg, use_cache=True):
    """-> dict(flat, y_train, sample_rows, sample_leaves, fit_s, cached).  ``sample_leaves`` are
    scikit-learn's own ``tree_.apply`` leaf ids (what forest.py:601 calls) of SAMPLE_ROWS test rows."""
    from skgarden_b200.datasets import make_test, make_train
    from skgarden_b200.flatten import FlatForest
    flat_path, extra_path = _cache_paths(name, cfg)
    rows = sample_row_ids(cfg["n_test"])
    if use_cache and os.path.isfile(flat_path) and os.path.isfile(extra_path):
        try:
            t0 = time.time()
            flat = FlatForest.load(flat_path)
            extra = np.load(extra_path)
            _, ytr = make_train(cfg["n_train"], cfg["n_features"], seed=0)
            log("[bench] loaded the flattened forest from %s (%.1f s)" % (flat_path, time.time() - t0))
            return dict(flat=flat, y_train=ytr, sample_rows=rows, sample_leaves=extra["sample_leaves"],
                        fit_s=float(extra["fit_s"]), cached=True, attrs_checked=bool(extra["attrs_checked"]))
        except Exception as exc:                           # a torn file from a killed run: refit
            log("[bench] cache unreadable (%s); refitting" % exc)
    t0 = time.time()
    est, Xtr, ytr = fit_forest(cfg)
    fit_s = time.time() - t0
    Xs = make_test(cfg["n_test"], cfg["n_features"], seed=0)[rows]
    sample_leaves = np.array([e.tree_.apply(Xs) for e in est.estimators_]).T
    # the dense fit attributes the CPU legs use are rebuilt from the member lists; pin them here,
    # while the scikit-learn trees exist, against the oracle's ensemble.py:83-101 on two trees
    from oracle import qrf_oracle as O
    pick = [0, len(est.estimators_) - 1]
    lv, wt = O.fit_attributes([est.estimators_[t] for t in pick], Xtr, bootstrap=True)
    dl, dw = dense_attributes(est._flat, trees=pick)
    attrs_checked = bool(np.array_equal(lv, dl) and np.array_equal(wt, dw))
    if not attrs_checked:
        raise RuntimeError("dense fit attributes rebuilt from the member lists differ from the oracle's")
    flat = est._flat
    if use_cache:
        try:
            os.makedirs(CACHE_DIR, exist_ok=True)
            tmp = flat_path + ".tmp.npz"
            flat.save(tmp)
            os.replace(tmp, flat_path)
            np.savez(extra_path + ".tmp.npz", sample_leaves=sample_leaves, fit_s=fit_s, attrs_checked=attrs_checked)
            os.replace(extra_path + ".tmp.npz", extra_path)
            log("[bench] cached the flattened forest under %s" % CACHE_DIR)
        except Exception as exc:
            log("[bench] could not write the forest cache (%s)" % exc)
            for p in (flat_path, extra_path):
                try:
                    os.remove(p)
                except OSError:
                    pass
    del est, Xtr
    return dict(flat=flat, y_train=ytr, sample_rows=rows, sample_leaves=sample_leaves, fit_s=fit_s,
                cached=False, attrs_checked=attrs_checked)


def dense_attributes(flat, trees=None):
    """(T', N) ``y_train_leaves_`` int32 / ``y_weights_`` float32 (ensemble.py:84-101) of the
    given trees, rebuilt from the flattened member lists (out-of-bag: -1 / 0)."""
    import skgarden_b200 as sg
    shell = sg.RandomForestQuantileRegressor()
    shell._set_flat(flat)
    return shell._dense_fit_attributes(trees=trees)


# --------------------------------------------------------------------------------------------
# CPU legs: the reference's per-row algorithm on host cores (bounded sample)
# --------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_worker(args):
    rows, qs = args
    from oracle import qrf_oracle as O
    t0 = time.perf_counter()
    for q in qs:                                      # scalar-q API: one call per quantile
        O.predict_dense(_CPU["y"], _CPU["leaves"], _CPU["weights"], _CPU["X_leaves"][rows], q, _CPU["sorter"])
    return time.perf_counter() - t0


def cpu_reference_setup(wl):
    """Dense fit attributes (ensemble.py:83-101) + scikit-learn leaf ids of the sampled rows."""
    t0 = time.time()
    leaves, weights = dense_attributes(wl["flat"])
    _CPU.update(y=wl["y_train"], leaves=leaves, weights=weights, X_leaves=wl["sample_leaves"],
                sorter=wl["flat"].sorter)
    log("[bench] dense (T,N) fit attributes for the CPU legs: %.1f s" % (time.time() - t0))


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


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


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
    return max(1, int((avail * 0.5) // (6 * tn)))


def sparse_index_for(wl, rows_idx):
    """oracle.SparseIndex restricted to the leaves the given sample rows reach (the other
    in-bag entries are masked to -1 = "out of bag", which never matches): the oracle's own
    class on a thinned input, so that building it at T x N = 5e8 takes seconds."""
    from oracle import qrf_oracle as O
    leaves, weights = _CPU["leaves"], _CPU["weights"]
    XL = wl["sample_leaves"][rows_idx]
    thin = np.full_like(leaves, -1)
    for t in range(leaves.shape[0]):
        keep = np.isin(leaves[t], np.unique(XL[:, t]))
        thin[t, keep] = leaves[t, keep]
    return O.SparseIndex(wl["y_train"], thin, weights, wl["flat"].sorter), XL


def measure_cpu_legs(wl, cfg, qs, budget_s=12.0):
    """-> (cpu_baseline all cores, 1-core figure, sparse restatement figure)."""
    from oracle import qrf_oracle as O
    cores = host_cores()
    workers = max(1, min(cores, 64, _memory_bound_workers(cfg)))
    q3 = qs[:3]
    cpu_reference_setup(wl)
    # (b) all host cores: forked workers on disjoint row chunks
    _, dt = cpu_reference_throughput(q3, 1, workers)              # calibrate: 1 row per worker
    rows = int(max(1, min(256, np.ceil(budget_s / max(dt, 1e-3)))))
    thr, dt = cpu_reference_throughput(q3, rows, workers)
    what = ("dense O(T*N)-per-row algorithm of ensemble.py:137-142 + utils.py:46-81 (oracle.predict_dense, "
            "pinned bit-identical to the reference; the reference is Python and cannot travel to the GPU box); "
            "leaf ids from scikit-learn's own apply, argsort outside the timed region")
    base = {"value": round(thr, 3), "unit": UNIT, "cores": workers, "kind": "port",
            "host_cpus": os.cpu_count(), "cpu_model": cpu_model(),
            "sample": "%d rows x %d quantiles (%d rows per worker, %d forked workers), %.1f s wall; %s"
                      % (rows * workers, len(q3), rows, workers, dt, what)}
    # (a) one core, as shipped: the per-row loop of ensemble.py:137 is single-threaded
    _, dt1 = cpu_reference_throughput(q3, 1, 1)
    rows1 = int(max(1, min(256, 0.5 * budget_s / max(dt1, 1e-3))))
    thr1, dt1 = cpu_reference_throughput(q3, rows1, 1)
    one = {"value": round(thr1, 3), "unit": UNIT, "cores": 1, "kind": "port",
           "sample": "%d rows x %d quantiles, %.1f s wall; the reference as shipped (ensemble.py:137 loop, 1 thread)"
                     % (rows1, len(q3), dt1)}
    # (c) NOT the reference: the sparse restatement (same arithmetic, only the non-zero weights)
    m = min(64, wl["sample_leaves"].shape[0])
    index, XL = sparse_index_for(wl, np.arange(m))
    t0 = time.perf_counter()
    O.predict_sparse(index, XL, q3)
    dts = time.perf_counter() - t0
    sparse = {"value": round(m * len(q3) / dts, 3), "unit": UNIT, "cores": 1, "kind": "port-sparse",
              "sample": "%d rows x %d quantiles, %.2f s wall; NOT the reference: oracle.predict_sparse "
                        "(numpy restatement restricted to the non-zero weights), 1 thread" % (m, len(q3), dts)}
    return base, one, sparse


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
    packed = np.empty_like(leaf_ids)
    for t in range(flat.n_trees):
        n0, n1 = flat.tree_offsets[t], flat.tree_offsets[t + 1]
        if flat.node_ids is None:
            packed[:, t] = leaf_ids[:, t] + n0
        else:
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


def load_traffic(config, mode):
    """DRAM bytes per step of each kernel from the committed ncu capture of THIS kernel version
    (profiles/traffic.json); entries made with another binary are ignored (-> null)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(path):
        try:
            ent = json.load(open(path)).get("%s_%s" % (config, mode), {})
            if ent.get("_binary") == KERNEL_TAG:
                return {k: v for k, v in ent.items() if not k.startswith("_")}
        except Exception:
            pass
    return {}


def run_b200(args, rank, world, local_rank):
    from skgarden_b200.datasets import CONFIGS, make_test
    from skgarden_b200.sharding import shard_rows
    cfg = CONFIGS[args.config]
    qs = cfg["quantiles"]
    n, F, T, Q = cfg["n_test"], cfg["n_features"], cfg["n_estimators"], len(cfg["quantiles"])

    # ---- forest (rank 0 fits or loads; CPU legs before any CUDA work so that fork is safe) ----
    wl, cpu_legs = None, None
    if rank == 0:
        wl = load_workload(args.config, cfg, use_cache=not args.no_cache)
        if world == 1 and not args.no_cpu_baseline:
            cpu_legs = measure_cpu_legs(wl, cfg, qs, budget_s=args.cpu_seconds)
        elif not args.no_parity:
            cpu_reference_setup(wl)                        # the parity sample needs the dense attributes
    import torch
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    flat = wl["flat"] if rank == 0 else None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        from skgarden_b200.sharding import broadcast_flat
        flat = broadcast_flat(flat, rank, device)
    import skgarden_b200 as sg
    from skgarden_b200 import _native
    dev = sg.DeviceForest(flat, device=local_rank)
    for kv in args.option:
        k, v = kv.split("=")
        dev.set_option(k, int(v))
    MODES = {"fast": _native.QF_MODE_FAST, "exact": _native.QF_MODE_EXACT}

    # ---- this rank's shard of the ONE test set (strong scaling) ----
    X_all = make_test(n, F, seed=0)
    lo, hi = shard_rows(n, world, rank)
    rows = hi - lo
    rows_max = -(-n // world)
    Xh = torch.from_numpy(X_all[lo:hi]).pin_memory()
    if rank != 0:
        del X_all                                          # rank 0 keeps it for the parity sample / estimator call
    Xd = Xh.to(device)
    # the gathered (n, Q) result of all ranks (padded to equal shards for all_gather)
    gathered = torch.empty((world * rows_max, Q), dtype=torch.float64, device=device)
    outd = gathered[rank * rows_max: rank * rows_max + rows]
    outh = torch.empty((rows, Q), dtype=torch.float64).pin_memory()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(mode):
        dev.predict_quantiles(Xd, qs, mode=mode, out=outd)
        if world > 1:                                       # the only collective: gather the outputs
            dist.all_gather_into_tensor(gathered, gathered[rank * rows_max:(rank + 1) * rows_max])

    def device_steps(k, mode, profile=False):
        """k steps, each bracketed by CUDA events on the launching (current) stream; L2 is
        flushed between steps outside the events.  Returns (sum ms, apply ms, quantile ms)."""
        dev.set_option("profile", 1 if profile else 0)
        dev._workspace(rows, Q)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        a_ms = q_ms = 0.0
        for s, e in evs:
            flush.zero_()
            s.record()
            one_step(mode)
            e.record()
            if profile:
                a, q = dev.last_timings()
                a_ms += a
                q_ms += q
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs), a_ms, q_ms

    def host_steps(k, mode):
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

    head = args.mode
    other = "exact" if head == "fast" else "fast"
    clocks = ClockSampler(local_rank)
    res = {}
    for mode_name in ([head] if args.single_mode else [other, head]):      # the headline mode last: its output is checked
        mode = MODES[mode_name]
        steps = args.steps if mode_name == head else max(3, args.steps // 4)
        device_steps(args.warmup if mode_name == head else 3, mode)
        barrier()
        if rank == 0 and mode_name == head:
            clocks.start()
        total_ms, _, _ = device_steps(steps, mode)
        barrier()
        launches = dev.last_launches
        total_ms = max_over_ranks(total_ms)
        _, apply_ms, quant_ms = device_steps(steps, mode, profile=True)
        barrier()
        host_steps(2, mode)
        barrier()
        e2e_ms = max_over_ranks(host_steps(steps, mode))
        barrier()
        res[mode_name] = dict(steps=steps, ms=total_ms / steps, k1=apply_ms / steps, k2=quant_ms / steps,
                              e2e=e2e_ms / steps, launches=launches)
    clock_info = clocks.stop() if rank == 0 else None
    # the timed output of the headline mode (device-resident steps ran last with profile on; run
    # one more plain step so that `gathered` is what a timed step leaves behind)
    dev.set_option("profile", 0)
    one_step(MODES[head])
    barrier()

    # ---- accounting ----
    r = res[head]
    ms_per_step = r["ms"]
    value = n * Q / (ms_per_step * 1e-3)
    e2e_value = n * Q / (r["e2e"] * 1e-3)
    if rank == 0:
        out_all = torch.cat([gathered[k * rows_max: k * rows_max + (shard_rows(n, world, k)[1] - shard_rows(n, world, k)[0])]
                             for k in range(world)]).cpu().numpy()
        srows = wl["sample_rows"]
        parity = None
        if not args.no_parity:
            parity = parity_sample(wl, dev, out_all, srows, qs, head, X_all[srows], lo, hi, device)
        # P_row on the sample (exact, from the leaves hit), scaled to all rows
        P_rows = row_member_counts(flat, wl["sample_leaves"])
        P_mean = float(P_rows.mean())
        peak, peak_src = load_peaks()
        traffic = load_traffic(args.config, head)
        # per row K2 = 4T (leaf ids) + 8T (CSR pair) + 8 P_row + 8Q (two y reads) + 8Q (out);
        # K1 = 4 n F + 4 n T + 8 * nodes
        # (this rank's shard: `rows` rows; every rank reads the whole forest once)
        bytes_k2 = float(rows) * (12 * T + 8 * P_mean + 16 * Q)
        bytes_k1 = float(4 * rows * F + 4 * rows * T + 8 * flat.n_nodes)
        k1_ms, k2_ms = r["k1"] * 1.0, r["k2"] * 1.0          # per step, this rank's shard

        def roof(name, nb, ms):
            ach = nb / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(ach / peak, 4), "traffic": traffic.get(name), "peak_source": peak_src,
                    "algorithmic_bytes": nb, "ms_per_step": round(ms, 4)}

        plan = dev.plan(rows)
        k2_name = dev.quantile_kernel_name(rows, Q, MODES[head]) if hasattr(dev, "quantile_kernel_name") else "quantile"
        r1 = roof("apply_kernel", bytes_k1, k1_ms)
        r2 = roof(k2_name, bytes_k2, k2_ms)
        chunks = -(-rows // plan["chunk_rows"])
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.config, cfg["label"]), "rows_total": n, "rows_per_gpu": rows,
                       "quantiles": qs, "mode": head, "l2": "flushed between timed steps (256 MiB write)",
                       "forest_bytes": dev.device_bytes, "mean_row_members": P_mean,
                       "kernel_chunks_per_step": chunks,
                       "sharding": "ONE test set, rows split into contiguous shards over the GPUs, forest replicated; "
                                   "the only collective is the all_gather of the (n, Q) result inside the timed step",
                       "forest_fit_s": round(wl["fit_s"], 1), "forest_cached": wl["cached"]},
            "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": rows * F * 4,
                    "d2h_bytes_per_step": rows * Q * 8, "ms_per_step": round(r["e2e"], 4),
                    "api": "qf_forest_predict_quantiles_host (pinned host X -> host result), each rank its shard"},
            "gpu_launches": r["launches"] * r["steps"],
            "roofline": r2 if k2_ms >= k1_ms or args.roofline_kernel == "k2" else r1,
            "roofline_kernels": {"apply_kernel": r1, k2_name: r2},
            "kernel_ms": {"apply": round(k1_ms, 4), "quantile": round(k2_ms, 4),
                          "locality_order_gather_and_gaps": round(max(0.0, ms_per_step - k1_ms - k2_ms), 4)},
            "clocks": clock_info,
            "parity_sample": parity,
        }
        if other in res:
            o = res[other]
            line["other_mode"] = {"mode": other, "value": round(n * Q / (o["ms"] * 1e-3), 1), "unit": UNIT,
                                  "ms_per_step": round(o["ms"], 4), "steps": o["steps"],
                                  "kernel_ms": {"apply": round(o["k1"], 4), "quantile": round(o["k2"], 4)},
                                  "e2e_ms_per_step": round(o["e2e"], 4)}
        if cpu_legs is not None:
            line["cpu_baseline"], line["cpu_baseline_1core"], line["cpu_baseline_sparse"] = cpu_legs
        if world == 1 and not args.no_estimator_e2e:
            line["e2e_estimator"] = estimator_e2e(flat, dev, X_all, qs, head, out_all)
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            log("[bench] PARITY FAILURE on the timed output: %s" % json.dumps(parity))
            dev.close()
            sys.exit(3)
    dev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_sample(wl, dev, out_all, srows, qs, mode, Xs, lo, hi, device):
    """BASELINE.md section 3.5 on the TIMED rows: sampled rows of the timed output against
    oracle.predict_sparse (exact: bit-equal; fast: <= 1e-4 of max|y|), and the library's apply
    on those rows against scikit-learn's own leaf ids."""
    import torch
    from oracle import qrf_oracle as O
    t0 = time.time()
    m = min(256, len(srows))
    index, XL = sparse_index_for(wl, np.arange(m))
    want = O.predict_sparse(index, XL, qs)
    got = out_all[srows[:m]]
    scale = float(np.abs(wl["y_train"]).max())
    err = float(np.abs(got - want).max())
    exact_equal = bool(np.array_equal(got, want))
    ok_q = exact_equal if mode == "exact" else bool(err <= 1e-4 * scale)
    leaves = dev.apply(torch.from_numpy(np.ascontiguousarray(Xs)).to(device)).cpu().numpy()
    ok_apply = bool(np.array_equal(leaves, wl["sample_leaves"]))
    return {"ok": ok_q and ok_apply, "rows": m, "quantiles": len(qs), "mode": mode,
            "bit_equal": exact_equal, "max_abs_err": err, "tolerance": 0.0 if mode == "exact" else 1e-4 * scale,
            "apply_rows": int(len(srows)), "apply_bit_equal": ok_apply,
            "oracle": "oracle.predict_sparse on the oracle's SparseIndex; leaf ids: scikit-learn tree_.apply",
            "seconds": round(time.time() - t0, 1)}


def estimator_e2e(flat, dev, X, qs, mode, out_all):
    """The drop-in call itself: est.predict(X, quantile=[...]) on a pageable numpy array
    (check_array + H2D + kernels + D2H), one repetition after one warm-up call of the same size."""
    import skgarden_b200 as sg
    est = sg.ExtraTreesQuantileRegressor(mode=mode)
    est._set_flat(flat)
    est._device_forest = dev
    est.n_features_ = est.n_features_in_ = flat.n_features
    est.predict(X, quantile=qs)                        # warm-up at size: staging buffers, workspaces
    t0 = time.perf_counter()
    got = est.predict(X, quantile=qs)
    dt = time.perf_counter() - t0
    est._device_forest = None
    return {"value": round(X.shape[0] * len(qs) / dt, 1), "unit": UNIT, "ms": round(dt * 1e3, 2),
            "api": "est.predict(X, quantile=[...]) on a pageable float32 array (check_array, H2D, kernels, D2H)",
            "equals_timed_output": bool(np.array_equal(got, out_all))}


# --------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm, all host threads, same config/metric
# --------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    from skgarden_b200.datasets import CONFIGS
    cfg = CONFIGS[args.config]
    qs = cfg["quantiles"][:3] if len(cfg["quantiles"]) > 3 else cfg["quantiles"]
    wl = load_workload(args.config, cfg, use_cache=not args.no_cache)
    workers = max(1, min(host_cores(), 64, _memory_bound_workers(cfg)))
    cpu_reference_setup(wl)
    # one step = `rows` rows per worker for ONE of the workload's quantiles (the reference's API takes a scalar
    # q, ensemble.py:104; the quantile rotates from step to step); bounded so that the whole --steps/--warmup
    # run ends within a few minutes even at config 3 (2 s per row and quantile)
    _, dt1 = cpu_reference_throughput(qs[:1], 1, workers)
    rows = int(max(1, min(64, (90.0 / max(1, args.steps + args.warmup)) / max(dt1, 1e-3))))
    for k in range(args.warmup):
        cpu_reference_throughput([qs[k % len(qs)]], rows, workers)
    t_total = 0.0
    for k in range(args.steps):
        _, dt = cpu_reference_throughput([qs[k % len(qs)]], rows, workers)
        t_total += dt
    per_step = rows * workers
    value = per_step * args.steps / t_total
    sample_txt = ("each step = %d rows x 1 quantile (of %d, rotating; %d rows per worker, %d forked workers) of the same "
                  "workload; dense O(T*N)-per-row algorithm (ensemble.py:137-142 + utils.py:46-81) as "
                  "restated in oracle/qrf_oracle.py (pinned bit-identical to the reference; the reference is "
                  "Python and cannot travel to the GPU box); leaf ids from scikit-learn's own apply"
                  % (rows * workers, len(qs), rows, workers))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t_total / args.steps * 1e3, 2),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.config, cfg["label"]), "rows_per_step": rows * workers,
                   "quantiles": qs},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": workers, "kind": "port",
                         "host_cpus": os.cpu_count(), "cpu_model": cpu_model(), "sample": sample_txt},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3")
    ap.add_argument("--mode", default="fast", choices=["exact", "fast"],
                    help="headline mode (the other one is timed too and reported under other_mode)")
    ap.add_argument("--single-mode", action="store_true", help="time only the headline mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-estimator-e2e", action="store_true")
    ap.add_argument("--no-cache", action="store_true", help="do not read or write the /tmp forest cache")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--roofline-kernel", default="auto", choices=["auto", "k2"])
    ap.add_argument("--option", action="append", default=[], help="library option key=value (sweeps)")
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
