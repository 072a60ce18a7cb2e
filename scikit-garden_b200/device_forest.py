"""The forest resident in HBM + the calls that run the CUDA path on it.

Thin Python over the C ABI (include/qforest_b200.h).  torch is plumbing only: it owns
the device tensors for X / outputs / workspace and the stream; every kernel is ours.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native
from .flatten import FlatForest


def _torch():
    import torch
    return torch


class DeviceForest:
    """A flattened forest uploaded to one GPU (replicated per GPU when row-sharding)."""

    def __init__(self, flat: FlatForest, device: int = 0, with_leaf_values: bool = True,
                 with_impurity: bool = False):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceForest needs a CUDA device (sm_100); there is no CPU fallback")
        self._lib = _native.lib()
        self.flat_meta = dict(n_trees=flat.n_trees, n_features=flat.n_features, n_train=flat.n_train,
                              n_nodes=flat.n_nodes, n_members=flat.n_members,
                              max_row_members=flat.max_row_members,
                              expected_row_members=flat.expected_row_members)
        self.device = int(device)
        self.n_trees = flat.n_trees
        self.n_features = flat.n_features
        desc = _native.ForestDesc()
        desc.device = self.device
        desc.n_trees = flat.n_trees
        desc.n_features = flat.n_features
        desc.feature_bits = flat.feature_bits
        desc.n_train = flat.n_train
        desc.n_nodes = flat.n_nodes
        desc.n_members = flat.n_members
        keep = [np.ascontiguousarray(flat.nodes), np.ascontiguousarray(flat.tree_offsets),
                None if flat.node_ids is None else np.ascontiguousarray(flat.node_ids),
                np.ascontiguousarray(flat.members), np.ascontiguousarray(flat.y_sorted),
                np.ascontiguousarray(flat.leaf_values) if with_leaf_values else None]
        desc.nodes = keep[0].ctypes.data
        desc.tree_offsets = keep[1].ctypes.data
        desc.node_ids = None if keep[2] is None else keep[2].ctypes.data
        desc.members = keep[3].ctypes.data
        desc.y_sorted = keep[4].ctypes.data
        desc.leaf_values = None if keep[5] is None else keep[5].ctypes.data
        desc.max_row_members = flat.max_row_members
        desc.expected_row_members = float(flat.expected_row_members)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.cuda.current_stream().synchronize()      # make sure the primary context exists
        handle = C.c_void_p()
        _native.check(self._lib.qf_forest_create(C.byref(desc), C.byref(handle)))
        self._h = handle
        self._ws = None
        # tree_.impurity per node is only needed by the return_std forests (predict_mean_std)
        self.has_impurity = with_impurity and flat.leaf_impurity is not None and with_leaf_values
        if self.has_impurity:
            imp = np.ascontiguousarray(flat.leaf_impurity, dtype=np.float64)
            _native.check(self._lib.qf_forest_set_leaf_impurity(self._h, C.c_void_p(imp.ctypes.data)))

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.qf_forest_destroy(self._h)
            self._h = None
            self._ws = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(self._lib.qf_forest_device_bytes(self._h))

    def set_option(self, key: str, value: int) -> None:
        _native.check(self._lib.qf_forest_set_option(self._h, key.encode(), int(value)))
        self._ws = None

    @property
    def last_launches(self) -> int:
        return int(self._lib.qf_forest_last_launches(self._h))

    def plan(self, n_rows: int) -> dict:
        """Which kernels a call over ``n_rows`` rows runs (qf_forest_plan)."""
        vals = (C.c_int32 * 8)()
        _native.check(self._lib.qf_forest_plan(self._h, int(n_rows), vals))
        keys = ("chunk_rows", "warp_elems", "s1_threads", "s1_limit", "s2_limit", "dense", "sort_rows", "reserved")
        return dict(zip(keys, [int(v) for v in vals]))

    def quantile_kernel_name(self, n_rows: int, n_q: int, mode: int) -> str:
        """Name of the weight+percentile kernel that takes the bulk of the rows of a call."""
        buf = C.create_string_buffer(96)
        _native.check(self._lib.qf_forest_quantile_kernel_name(self._h, int(n_rows), int(n_q), int(mode), buf, 96))
        return buf.value.decode()

    def last_timings(self):
        """(apply_ms, quantile_ms) of the most recent call; needs set_option("profile", 1)."""
        a, q = C.c_double(0.0), C.c_double(0.0)
        _native.check(self._lib.qf_forest_last_timings(self._h, C.byref(a), C.byref(q)))
        return a.value, q.value

    # -- helpers ----------------------------------------------------------------------
    def _workspace(self, n_rows: int, n_q: int):
        torch = _torch()
        need = int(self._lib.qf_forest_workspace_bytes(self._h, n_rows, n_q))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device="cuda:%d" % self.device)
        return self._ws, need

    def _check_X(self, X):
        torch = _torch()
        if not (isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float32 and X.dim() == 2):
            raise TypeError("X must be a 2-D float32 CUDA tensor")
        if X.device.index != self.device:
            raise ValueError("X is on cuda:%s, forest on cuda:%d" % (X.device.index, self.device))
        if X.shape[1] != self.n_features:
            raise ValueError("X has %d features, forest expects %d" % (X.shape[1], self.n_features))
        if X.stride(1) != 1:
            X = X.contiguous()
        return X

    def _stream(self):
        torch = _torch()
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- device-resident calls ----------------------------------------------------------
    def apply(self, X):
        """(n, T) int64 CUDA tensor of scikit-learn leaf ids (forest.py:583-604)."""
        torch = _torch()
        X = self._check_X(X)
        n = X.shape[0]
        out = torch.empty((n, self.n_trees), dtype=torch.int64, device=X.device)
        ws, _ = self._workspace(n, 1)
        _native.check(self._lib.qf_forest_apply(
            self._h, C.c_void_p(X.data_ptr()), n, X.stride(0), C.c_void_p(out.data_ptr()),
            C.c_void_p(ws.data_ptr()), ws.numel(), self._stream()))
        return out

    def predict_quantiles(self, X, quantiles: Sequence[float], mode: int = _native.QF_MODE_EXACT, out=None):
        """(n, Q) float64 CUDA tensor (ensemble.py:104-143 for Q percentiles at once)."""
        torch = _torch()
        X = self._check_X(X)
        n = X.shape[0]
        q = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
        if out is None:
            out = torch.empty((n, q.shape[0]), dtype=torch.float64, device=X.device)
        ws, _ = self._workspace(n, q.shape[0])
        _native.check(self._lib.qf_forest_predict_quantiles(
            self._h, C.c_void_p(X.data_ptr()), n, X.stride(0), q.ctypes.data_as(C.c_void_p),
            q.shape[0], C.c_void_p(out.data_ptr()), int(mode),
            C.c_void_p(ws.data_ptr()), ws.numel(), self._stream()))
        return out

    def predict_mean(self, X):
        """(n,) float64 CUDA tensor (forest.py:1093-1131)."""
        torch = _torch()
        X = self._check_X(X)
        n = X.shape[0]
        out = torch.empty((n,), dtype=torch.float64, device=X.device)
        ws, _ = self._workspace(n, 1)
        _native.check(self._lib.qf_forest_predict_mean(
            self._h, C.c_void_p(X.data_ptr()), n, X.stride(0), C.c_void_p(out.data_ptr()),
            C.c_void_p(ws.data_ptr()), ws.numel(), self._stream()))
        return out

    def predict_mean_std(self, X, min_variance: float = 0.0):
        """((n,), (n,)) float64 CUDA tensors: forest mean and std(y | x) of the reference's
        ``return_std`` forests (forest.py:133-178, :332-362)."""
        torch = _torch()
        if not self.has_impurity:
            raise RuntimeError("create the DeviceForest with with_impurity=True (and a forest flattened "
                               "from trees that carry tree_.impurity)")
        X = self._check_X(X)
        n = X.shape[0]
        mean = torch.empty((n,), dtype=torch.float64, device=X.device)
        std = torch.empty((n,), dtype=torch.float64, device=X.device)
        ws, _ = self._workspace(n, 1)
        _native.check(self._lib.qf_forest_predict_mean_std(
            self._h, C.c_void_p(X.data_ptr()), n, X.stride(0), float(min_variance),
            C.c_void_p(mean.data_ptr()), C.c_void_p(std.data_ptr()),
            C.c_void_p(ws.data_ptr()), ws.numel(), self._stream()))
        return mean, std

    # -- host-buffer end-to-end call ------------------------------------------------------
    def predict_quantiles_host(self, X_host: np.ndarray, quantiles: Sequence[float],
                               out_host: Optional[np.ndarray] = None,
                               mode: int = _native.QF_MODE_EXACT, chunk_rows: int = 0) -> np.ndarray:
        """Host float32 (n,F) in, host float64 (n,Q) out; H2D, kernels and D2H run inside
        the C call, pipelined on three streams.  Pass pinned arrays for full PCIe bandwidth."""
        if X_host.dtype != np.float32 or X_host.ndim != 2 or X_host.strides[1] != 4:
            raise TypeError("X_host must be a 2-D float32 array with contiguous rows")
        if X_host.shape[1] != self.n_features:
            raise ValueError("X has %d features, forest expects %d" % (X_host.shape[1], self.n_features))
        q = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
        n = X_host.shape[0]
        if out_host is None:
            out_host = np.empty((n, q.shape[0]), dtype=np.float64)
        _native.check(self._lib.qf_forest_predict_quantiles_host(
            self._h, C.c_void_p(X_host.ctypes.data), n, X_host.strides[0] // 4,
            q.ctypes.data_as(C.c_void_p), q.shape[0], C.c_void_p(out_host.ctypes.data),
            int(mode), int(chunk_rows)))
        return out_host

    # -- host-array conveniences (what the estimators call) -------------------------------
    def _to_device(self, X_host):
        torch = _torch()
        return torch.from_numpy(np.ascontiguousarray(X_host, dtype=np.float32)).to("cuda:%d" % self.device)

    def apply_host(self, X_host: np.ndarray) -> np.ndarray:
        torch = _torch()
        with torch.cuda.device(self.device):
            return self.apply(self._to_device(X_host)).cpu().numpy()

    def predict_mean_host(self, X_host: np.ndarray) -> np.ndarray:
        torch = _torch()
        with torch.cuda.device(self.device):
            return self.predict_mean(self._to_device(X_host)).cpu().numpy()

    def predict_mean_std_host(self, X_host: np.ndarray, min_variance: float = 0.0):
        torch = _torch()
        with torch.cuda.device(self.device):
            mean, std = self.predict_mean_std(self._to_device(X_host), min_variance)
            return mean.cpu().numpy(), std.cpu().numpy()

    def predict_stream(self, chunks, quantiles: Sequence[float], mode: int = _native.QF_MODE_EXACT,
                       pinned: bool = True, prepare=None):
        """Streaming front end (BASELINE.json config 5: test rows that do not fit one host
        array): ``chunks`` is any iterable of (m_i, F) float32 arrays; yields one (m_i, Q)
        float64 array per chunk, in order.  See :func:`stream_predict`."""
        return stream_predict([self], chunks, quantiles, mode=mode, pinned=pinned, prepare=prepare)


def assert_all_finite_parallel(X: np.ndarray, workers: int = 8, pool=None) -> None:
    """The finite check of ``check_array`` (ensemble.py:129 rejects NaN / inf), on row slices in
    parallel (numpy releases the GIL inside the scan).  Raises scikit-learn's own ValueError."""
    from sklearn.utils import assert_all_finite
    n = X.shape[0]
    if n * max(1, X.shape[1]) < (1 << 20) or workers <= 1:
        assert_all_finite(X, input_name="X")
        return
    from concurrent.futures import ThreadPoolExecutor
    bounds = np.linspace(0, n, workers + 1).astype(np.int64)
    jobs = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    own = pool is None
    pool = pool or ThreadPoolExecutor(len(jobs))
    try:
        for f in [pool.submit(assert_all_finite, X[a:b], input_name="X") for a, b in jobs]:
            f.result()
    finally:
        if own:
            pool.shutdown(wait=True)


def stream_predict(replicas, chunks, quantiles, mode=_native.QF_MODE_EXACT, pinned=True, prepare=None):
    """Chunks of test rows through one or several forest replicas, results in order.

    Three worker threads per replica: while one chunk is inside the pipelined host-buffer call
    (H2D / kernels / D2H on three streams, csrc/qforest_abi.cu), the next ones are being validated and
    staged into their own page-locked buffers, so the GPU does not wait for the host copy; chunk k goes
    to replica k mod G.  ``pinned=False`` hands the caller's arrays to the C call as they are.
    ``prepare`` (chunk -> float32 C-contiguous array, may raise) runs in the worker threads."""
    import collections
    import threading
    from concurrent.futures import ThreadPoolExecutor
    torch = _torch()
    q = [float(v) for v in quantiles]
    F = replicas[0].n_features
    tls = threading.local()

    def work(dev, X):
        if prepare is not None:                         # the caller's validation, off the feeding thread
            X = prepare(X)
        m = X.shape[0]
        if not pinned:
            return dev.predict_quantiles_host(X, q, mode=mode)
        st = getattr(tls, "stage", None)
        if st is None or st[0].shape[0] < m:
            st = (torch.empty((m, F), dtype=torch.float32).pin_memory(),
                  torch.empty((m, len(q)), dtype=torch.float64).pin_memory())
            tls.stage = st
        xs, os_ = st[0][:m].numpy(), st[1][:m].numpy()
        xs[...] = X
        dev.predict_quantiles_host(xs, q, out_host=os_, mode=mode)
        return os_.copy()

    pending = collections.deque()
    depth = 3 * len(replicas)                           # chunks in flight: one in the GPU call, two being validated / staged
    with ThreadPoolExecutor(depth) as pool:
        k = 0
        for X in chunks:
            if np.ndim(X) != 2 or np.shape(X)[1] != F:
                raise ValueError("chunk has shape %r, forest expects (m, %d)" % (np.shape(X), F))
            if prepare is None:
                X = np.ascontiguousarray(X, dtype=np.float32)
            if np.shape(X)[0] == 0:
                pending.append(None)
            else:
                pending.append(pool.submit(work, replicas[k % len(replicas)], X))
                k += 1
            while len(pending) > depth or (pending and pending[0] is None):
                head = pending.popleft()
                yield np.empty((0, len(q)), dtype=np.float64) if head is None else head.result()
        while pending:
            head = pending.popleft()
            yield np.empty((0, len(q)), dtype=np.float64) if head is None else head.result()


class ForestReplicas:
    """The forest replicated on several GPUs of one box; every call splits its rows into
    contiguous shards (skgarden/quantile/ensemble.py:137: rows are independent), one host
    thread per GPU, results written straight into one output array.  No collective: the only
    traffic is each GPU's own H2D / D2H."""

    def __init__(self, flat: FlatForest, devices: Sequence[int], with_impurity: bool = False):
        from concurrent.futures import ThreadPoolExecutor
        devices = [int(d) for d in devices]
        if not devices or len(set(devices)) != len(devices):
            raise ValueError("devices must be a non-empty list of distinct CUDA ordinals")
        self.devices = devices
        self.replicas = [DeviceForest(flat, device=d, with_impurity=with_impurity) for d in devices]
        self.n_trees, self.n_features = flat.n_trees, flat.n_features
        self.device = devices[0]
        self._pool = ThreadPoolExecutor(len(devices))

    def close(self):
        for r in self.replicas:
            r.close()
        self._pool.shutdown(wait=True)

    def set_option(self, key: str, value: int) -> None:
        for r in self.replicas:
            r.set_option(key, value)

    @property
    def device_bytes(self) -> int:
        return sum(r.device_bytes for r in self.replicas)

    def _sharded(self, n, fn):
        from .sharding import shard_rows
        G = len(self.replicas)
        jobs = [(self.replicas[g],) + shard_rows(n, G, g) for g in range(G)]
        futures = [self._pool.submit(fn, dev, lo, hi) for dev, lo, hi in jobs if hi > lo]
        for f in futures:
            f.result()

    def predict_quantiles_host(self, X_host, quantiles, out_host=None, mode=_native.QF_MODE_EXACT, chunk_rows=0):
        q = np.ascontiguousarray(quantiles, dtype=np.float64).ravel()
        n = X_host.shape[0]
        if out_host is None:
            out_host = np.empty((n, q.shape[0]), dtype=np.float64)
        self._sharded(n, lambda dev, lo, hi: dev.predict_quantiles_host(
            X_host[lo:hi], q, out_host=out_host[lo:hi], mode=mode, chunk_rows=chunk_rows))
        return out_host

    def apply_host(self, X_host):
        out = np.empty((X_host.shape[0], self.n_trees), dtype=np.int64)

        def fn(dev, lo, hi):
            out[lo:hi] = dev.apply_host(X_host[lo:hi])
        self._sharded(X_host.shape[0], fn)
        return out

    def predict_mean_host(self, X_host):
        out = np.empty(X_host.shape[0], dtype=np.float64)

        def fn(dev, lo, hi):
            out[lo:hi] = dev.predict_mean_host(X_host[lo:hi])
        self._sharded(X_host.shape[0], fn)
        return out

    def predict_mean_std_host(self, X_host, min_variance=0.0):
        mean = np.empty(X_host.shape[0], dtype=np.float64)
        std = np.empty(X_host.shape[0], dtype=np.float64)

        def fn(dev, lo, hi):
            mean[lo:hi], std[lo:hi] = dev.predict_mean_std_host(X_host[lo:hi], min_variance)
        self._sharded(X_host.shape[0], fn)
        return mean, std

    def predict_stream(self, chunks, quantiles, mode=_native.QF_MODE_EXACT, pinned=True, prepare=None):
        return stream_predict(self.replicas, chunks, quantiles, mode=mode, pinned=pinned, prepare=prepare)
