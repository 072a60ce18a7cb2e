"""Host-side forest flattener: fitted scikit-learn trees -> the HBM layout.

Replaces the fit bookkeeping of the reference (skgarden/quantile/ensemble.py:83-101: the
dense (T,N) ``y_train_leaves_`` / ``y_weights_`` arrays, built by an O(T*L*N) loop) and
scikit-learn's 64-byte AoS node records (sklearn/tree/_tree.pxd:15-25) by

* 8-byte packed nodes in level order with sibling pairs (one padding slot per tree, see
  ``include/qforest_b200.h``), trees back to back,
* per-tree leaf -> in-bag member lists (CSR) whose entries are
  (rank in ``np.argsort(y_train_)``, float32 weight), sorted by rank inside a leaf, the
  leaves from left to right,
* ``y_sorted = y_train_.astype(float32)[sorter]`` (skgarden/quantile/utils.py:46,52).

The packing itself is native code (``qf_pack_tree`` in csrc/pack_forest.cpp, called
through the C ABI, one tree per thread); this module only orchestrates it.
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _native


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def bootstrap_counts(random_state, n_samples: int) -> np.ndarray:
    """In-bag multiplicities of one tree, regenerated from the tree's seed exactly as
    the reference does (ensemble.py:16-38, :88-94; same call scikit-learn's
    ``_generate_sample_indices`` makes when fitting with ``sample_weight=None``)."""
    rs = np.random.RandomState(random_state)
    idx = rs.randint(0, n_samples, n_samples)
    return np.bincount(idx, minlength=n_samples).astype(np.int64)


@dataclass
class FlatForest:
    """The flattened forest in host memory (what ``qf_forest_create`` uploads)."""
    n_trees: int
    n_features: int
    feature_bits: int
    n_train: int
    nodes: np.ndarray            # uint64 [n_nodes] packed slots (sklearn node_count + 1 per tree)
    tree_offsets: np.ndarray     # int64  [T+1]
    node_ids: Optional[np.ndarray]   # int32 [n_nodes] sklearn id of every slot (-1 = padding)
    members: np.ndarray          # uint64 [n_members]
    member_offsets: np.ndarray   # int64  [T+1]
    y_sorted: np.ndarray         # float32 [N]
    leaf_values: np.ndarray      # float64 [n_nodes]
    max_row_members: int         # sum over trees of the largest leaf (upper bound per row)
    expected_row_members: float  # sum over trees of E[leaf size] for a training-like row
    max_depth: int
    sorter: np.ndarray = field(repr=False, default=None)   # int64 [N] np.argsort(y_train_)
    leaf_impurity: Optional[np.ndarray] = field(repr=False, default=None)   # float64 [n_nodes] tree_.impurity

    @property
    def n_nodes(self) -> int:
        return int(self.nodes.shape[0])

    @property
    def n_members(self) -> int:
        return int(self.members.shape[0])

    def nbytes(self) -> int:
        return int(self.nodes.nbytes + self.members.nbytes + self.y_sorted.nbytes +
                   self.leaf_values.nbytes +
                   (0 if self.node_ids is None else self.node_ids.nbytes))

    # -- on-disk format: a plain .npz of the arrays above
    def save(self, path: str) -> None:
        # through an open handle: np.savez(str) would append ".npz" to a path without it and
        # load(path) with the same string would then fail
        with open(path, "wb") as fh:
            self._savez(fh)

    def _savez(self, fh) -> None:
        np.savez(fh, layout=LAYOUT_VERSION, n_trees=self.n_trees, n_features=self.n_features,
                 feature_bits=self.feature_bits, n_train=self.n_train, nodes=self.nodes,
                 tree_offsets=self.tree_offsets,
                 node_ids=np.zeros(0, np.int32) if self.node_ids is None else self.node_ids,
                 members=self.members, member_offsets=self.member_offsets,
                 y_sorted=self.y_sorted, leaf_values=self.leaf_values,
                 max_row_members=self.max_row_members,
                 expected_row_members=self.expected_row_members, max_depth=self.max_depth,
                 sorter=self.sorter,
                 leaf_impurity=np.zeros(0) if self.leaf_impurity is None else self.leaf_impurity)

    @classmethod
    def load(cls, path: str) -> "FlatForest":
        z = np.load(path)
        if "layout" not in z.files or int(z["layout"]) != LAYOUT_VERSION:
            raise ValueError("%s holds a forest in another packed layout (file %s, library %d); "
                             "flatten the forest again" % (path, int(z["layout"]) if "layout" in z.files else 1,
                                                          LAYOUT_VERSION))
        node_ids = z["node_ids"]
        return cls(int(z["n_trees"]), int(z["n_features"]), int(z["feature_bits"]),
                   int(z["n_train"]), z["nodes"], z["tree_offsets"],
                   None if node_ids.size == 0 else node_ids, z["members"],
                   z["member_offsets"], z["y_sorted"], z["leaf_values"],
                   int(z["max_row_members"]), float(z["expected_row_members"]),
                   int(z["max_depth"]), z["sorter"],
                   z["leaf_impurity"] if "leaf_impurity" in z.files and z["leaf_impurity"].size else None)


LAYOUT_VERSION = 2      # 1: depth-first preorder nodes; 2: level order with sibling pairs


def _per_slot(values, ids):
    """values[sklearn id] -> per packed slot (0 in the padding slot, id -1)."""
    return np.where(ids >= 0, values[np.maximum(ids, 0)], 0.0)


def _tree_arrays(tree):
    """children_left/right, feature (int64), threshold (f64), value (f64 per node)."""
    t = getattr(tree, "tree_", tree)
    left = np.ascontiguousarray(t.children_left, dtype=np.int64)
    right = np.ascontiguousarray(t.children_right, dtype=np.int64)
    feat = np.ascontiguousarray(t.feature, dtype=np.int64)
    thr = np.ascontiguousarray(t.threshold, dtype=np.float64)
    val = np.asarray(t.value, dtype=np.float64)
    val = np.ascontiguousarray(val.reshape(val.shape[0], -1)[:, 0])
    return left, right, feat, thr, val


def _tree_impurity(tree):
    """tree_.impurity (float64 per node) or None for duck-typed trees without it."""
    t = getattr(tree, "tree_", tree)
    imp = getattr(t, "impurity", None)
    return None if imp is None else np.ascontiguousarray(imp, dtype=np.float64)


def flatten_forest(trees: Sequence, train_leaves, counts, y_train, n_features: int,
                   sorter: Optional[np.ndarray] = None, n_jobs: Optional[int] = None,
                   unit_weights: bool = False) -> FlatForest:
    """Flatten ``trees`` (fitted sklearn estimators, ``Tree`` objects, or anything with
    children_left/right, feature, threshold, value).

    train_leaves : (T, N) int array or sequence of N-vectors: ``tree_.apply(X_train)``
                   per tree (tree.py:101), or a callable ``t -> N-vector``.
    counts       : sequence of per-tree in-bag multiplicity vectors (ensemble.py:88-94),
                   entries may be None (= every sample once, bootstrap=False), or None.
    y_train      : (N,) training targets as given to fit (ensemble.py:83).
    sorter       : np.argsort(y_train) computed by the caller; computed here if None.
                   It is never recomputed on the device (ties are ordered by numpy).
    unit_weights : every member weighs 1.0 instead of count / leaf count sum: the
                   tree-level estimators (quantile/tree.py:45-47 call weighted_percentile
                   without weights, utils.py:40-41).
    """
    L = _native.lib()
    T = len(trees)
    if T == 0:
        raise ValueError("empty forest")
    y_train = np.asarray(y_train)
    N = int(y_train.shape[0])
    if sorter is None:
        sorter = np.argsort(y_train)                      # ensemble.py:133
    sorter = np.ascontiguousarray(sorter, dtype=np.int64)
    y_sorted = np.ascontiguousarray(np.asarray(y_train, dtype=np.float32)[sorter])   # utils.py:46,52
    feature_bits = max(1, int(n_features - 1).bit_length())

    arrays = [_tree_arrays(t) for t in trees]
    node_counts = np.array([a[0].shape[0] for a in arrays], dtype=np.int64)
    tree_offsets = np.concatenate([[0], np.cumsum(node_counts + 1)]).astype(np.int64)    # + the padding slot

    def counts_of(t):
        if counts is None or counts[t] is None:
            return None
        return np.ascontiguousarray(counts[t], dtype=np.int64)

    inbag = np.array([N if counts_of(t) is None else int(np.count_nonzero(counts_of(t) > 0))
                      for t in range(T)], dtype=np.int64)
    member_offsets = np.concatenate([[0], np.cumsum(inbag)]).astype(np.int64)

    nodes = np.empty(int(tree_offsets[-1]), dtype=np.uint64)
    node_ids = np.empty(int(tree_offsets[-1]), dtype=np.int32)
    members = np.empty(int(member_offsets[-1]), dtype=np.uint64)
    leaf_values = np.empty(int(tree_offsets[-1]), dtype=np.float64)
    impurities = [_tree_impurity(t) for t in trees]
    leaf_impurity = None if any(i is None for i in impurities) else np.empty(int(tree_offsets[-1]), dtype=np.float64)
    stats = np.zeros((T, 3), dtype=np.float64)            # max_leaf_members, depth, sum_sq

    def pack(t):
        left, right, feat, thr, val = arrays[t]
        lv = train_leaves(t) if callable(train_leaves) else train_leaves[t]
        lv = np.ascontiguousarray(lv, dtype=np.int64)
        if lv.shape[0] != N:
            raise ValueError("train_leaves[%d] has %d entries, expected %d" % (t, lv.shape[0], N))
        cnt = counts_of(t)
        n0, n1 = int(tree_offsets[t]), int(tree_offsets[t + 1])
        m0, m1 = int(member_offsets[t]), int(member_offsets[t + 1])
        n_inbag, max_members, depth = C.c_int64(0), C.c_int32(0), C.c_int32(0)
        sum_sq = C.c_double(0.0)
        nodes_t, ids_t, members_t = nodes[n0:n1], node_ids[n0:n1], members[m0:m1]
        rc = L.qf_pack_tree(left.shape[0], _ptr(left), _ptr(right), _ptr(feat), _ptr(thr), feature_bits,
                            N, _ptr(lv), _ptr(cnt), _ptr(sorter), m0,
                            _ptr(nodes_t), _ptr(ids_t), _ptr(members_t) if m1 > m0 else None,
                            C.byref(n_inbag), C.byref(max_members), C.byref(depth), C.byref(sum_sq))
        if rc != 0:
            raise _native.QFError(rc, "tree %d: %s" % (t, _native.last_error()))
        if n_inbag.value != m1 - m0:
            raise RuntimeError("tree %d: in-bag count mismatch" % t)
        leaf_values[n0:n1] = _per_slot(val, ids_t)
        if leaf_impurity is not None:
            leaf_impurity[n0:n1] = _per_slot(impurities[t], ids_t)
        stats[t] = (max_members.value, depth.value, sum_sq.value)

    workers = n_jobs if n_jobs and n_jobs > 0 else min(T, os.cpu_count() or 1)
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(pack, range(T)))
    else:
        for t in range(T):
            pack(t)

    if unit_weights:
        members &= np.uint64(0xFFFFFFFF)
        members |= np.uint64(0x3F800000) << np.uint64(32)          # float32 1.0
    return FlatForest(
        n_trees=T, n_features=int(n_features), feature_bits=feature_bits, n_train=N,
        nodes=nodes, tree_offsets=tree_offsets, node_ids=node_ids,
        members=members, member_offsets=member_offsets, y_sorted=y_sorted,
        leaf_values=leaf_values, max_row_members=int(stats[:, 0].sum()),
        expected_row_members=float((stats[:, 2] / np.maximum(inbag, 1)).sum()),
        max_depth=int(stats[:, 1].max()), sorter=sorter, leaf_impurity=leaf_impurity)


def flatten_forest_device(trees: Sequence, X_train, counts, y_train, n_features: int,
                          sorter: Optional[np.ndarray] = None, n_jobs: Optional[int] = None,
                          unit_weights: bool = False, device: int = 0) -> FlatForest:
    """Same result as :func:`flatten_forest`, bit for bit, with the member lists built on
    the GPU (``qf_build_csr``, csrc/csr_builder.cuh): no ``tree_.apply(X_train)`` on the CPU
    (tree.py:98-101) and no per-sample host loop (ensemble.py:87-101).  Only the tree
    STRUCTURE is packed on the host.  Raises ``QFError`` (QF_ERR_LIMIT) when a leaf holds
    more than 4096 members -- use :func:`flatten_forest` for such shallow trees.

    X_train : (N, F) training rows as given to fit (cast to float32 like ensemble.py:79).
    counts  : as for :func:`flatten_forest` (per-tree multiplicity vectors or None).
    """
    import torch
    L = _native.lib()
    T = len(trees)
    if T == 0:
        raise ValueError("empty forest")
    y_train = np.asarray(y_train)
    N = int(y_train.shape[0])
    if sorter is None:
        sorter = np.argsort(y_train)                      # ensemble.py:133
    sorter = np.ascontiguousarray(sorter, dtype=np.int64)
    y_sorted = np.ascontiguousarray(np.asarray(y_train, dtype=np.float32)[sorter])   # utils.py:46,52
    feature_bits = max(1, int(n_features - 1).bit_length())
    arrays = [_tree_arrays(t) for t in trees]
    node_counts = np.array([a[0].shape[0] for a in arrays], dtype=np.int64)
    tree_offsets = np.concatenate([[0], np.cumsum(node_counts + 1)]).astype(np.int64)    # + the padding slot
    n_nodes = int(tree_offsets[-1])
    nodes = np.empty(n_nodes, dtype=np.uint64)
    node_ids = np.empty(n_nodes, dtype=np.int32)
    leaf_values = np.empty(n_nodes, dtype=np.float64)
    impurities = [_tree_impurity(t) for t in trees]
    leaf_impurity = None if any(i is None for i in impurities) else np.empty(n_nodes, dtype=np.float64)
    depths = np.zeros(T, dtype=np.int64)

    def pack(t):                                          # structure only (train_leaves = NULL)
        left, right, feat, thr, val = arrays[t]
        n0, n1 = int(tree_offsets[t]), int(tree_offsets[t + 1])
        depth = C.c_int32(0)
        nodes_t, ids_t = nodes[n0:n1], node_ids[n0:n1]
        # member_base = first leaf ordinal of the tree (a tree of k slots has k / 2 leaves)
        rc = L.qf_pack_tree(left.shape[0], _ptr(left), _ptr(right), _ptr(feat), _ptr(thr), feature_bits,
                            0, None, None, None, n0 // 2, _ptr(nodes_t), _ptr(ids_t), None,
                            None, None, C.byref(depth), None)
        if rc != 0:
            raise _native.QFError(rc, "tree %d: %s" % (t, _native.last_error()))
        leaf_values[n0:n1] = _per_slot(val, ids_t)
        if leaf_impurity is not None:
            leaf_impurity[n0:n1] = _per_slot(impurities[t], ids_t)
        depths[t] = depth.value

    workers = n_jobs if n_jobs and n_jobs > 0 else min(T, os.cpu_count() or 1)
    if workers > 1:
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(pack, range(T)))
    else:
        for t in range(T):
            pack(t)

    if counts is None or all(c is None for c in counts):
        counts_nt, capacity = None, N * T
    else:
        cm = np.empty((T, N), dtype=np.uint16)            # [T][N] here, transposed on the device
        capacity = 0
        for t, c in enumerate(counts):
            if c is None:
                cm[t] = 1
                capacity += N
                continue
            c = np.asarray(c)
            if c.min() < 0 or c.max() > 65535:
                raise ValueError("bootstrap multiplicities must be in [0, 65535] for the device builder")
            cm[t] = c
            capacity += int(np.count_nonzero(c))
        counts_nt = cm
    rank = np.empty(N, dtype=np.int32)
    rank[sorter] = np.arange(N, dtype=np.int32)

    dev = torch.device("cuda", device)
    with torch.cuda.device(dev):
        d_nodes = torch.from_numpy(nodes.view(np.int64)).to(dev)
        d_X = torch.from_numpy(np.ascontiguousarray(X_train, dtype=np.float32)).to(dev)
        d_rank = torch.from_numpy(rank).to(dev)
        d_counts = None if counts_nt is None else torch.from_numpy(counts_nt.view(np.int16)).to(dev).t().contiguous()   # [N][T]
        d_members = torch.empty(max(capacity, 1), dtype=torch.int64, device=dev)
        inbag = np.zeros(T, dtype=np.int64)
        max_leaf = np.zeros(T, dtype=np.int32)
        sum_sq = np.zeros(T, dtype=np.float64)
        desc = _native.CsrBuildDesc()
        desc.device = device
        desc.n_trees, desc.n_features, desc.feature_bits = T, int(n_features), feature_bits
        desc.n_train, desc.n_nodes = N, n_nodes
        desc.nodes_dev = d_nodes.data_ptr()
        desc.tree_offsets = tree_offsets.ctypes.data
        desc.X_train_dev = d_X.data_ptr()
        desc.rank_dev = d_rank.data_ptr()
        desc.counts_dev = None if d_counts is None else d_counts.data_ptr()
        desc.members_dev = d_members.data_ptr()
        desc.members_capacity = capacity
        desc.inbag_out = inbag.ctypes.data
        desc.max_leaf_members_out = max_leaf.ctypes.data
        desc.sum_sq_out = sum_sq.ctypes.data
        torch.cuda.synchronize()
        n_members = C.c_int64(0)
        _native.check(L.qf_build_csr(C.byref(desc), C.byref(n_members)))
        if n_members.value != capacity:
            raise RuntimeError("device builder produced %d members, expected %d" % (n_members.value, capacity))
        nodes = d_nodes.cpu().numpy().view(np.uint64)
        members = d_members[:capacity].cpu().numpy().view(np.uint64)
    if unit_weights:
        members = members.copy()
        members &= np.uint64(0xFFFFFFFF)
        members |= np.uint64(0x3F800000) << np.uint64(32)
    member_offsets = np.concatenate([[0], np.cumsum(inbag)]).astype(np.int64)
    return FlatForest(
        n_trees=T, n_features=int(n_features), feature_bits=feature_bits, n_train=N,
        nodes=nodes, tree_offsets=tree_offsets, node_ids=node_ids,
        members=members, member_offsets=member_offsets, y_sorted=y_sorted,
        leaf_values=leaf_values, max_row_members=int(max_leaf.sum()),
        expected_row_members=float((sum_sq / np.maximum(inbag, 1)).sum()),
        max_depth=int(depths.max()), sorter=sorter, leaf_impurity=leaf_impurity)

