"""The C-ABI library loads and exports every symbol include/qforest_b200.h declares
(no compute calls: this runs on the CPU-only box)."""
import ctypes
import os
import re

import numpy as np
import pytest

from skgarden_b200 import _native


def declared_symbols(repo_root):
    text = open(os.path.join(repo_root, "include", "qforest_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qf_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(repo_root):
    names = declared_symbols(repo_root)
    assert len(names) >= 12
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(_native.SIGNATURES) == names        # the binding covers the whole header


def test_abi_version_and_error_slot():
    lib = _native.lib()
    assert lib.qf_abi_version() == _native.ABI_VERSION
    rc = lib.qf_forest_set_option(None, b"apply_ilp", 4)
    assert rc == _native.QF_ERR_INVALID
    assert "null" in _native.last_error()


def test_pack_tree_argument_checks():
    lib = _native.lib()
    rc = lib.qf_pack_tree(0, None, None, None, None, 4, 0, None, None, None, 0,
                          None, None, None, None, None, None, None)
    assert rc == _native.QF_ERR_INVALID
    assert lib.qf_count_inbag(4, np.array([0, 2, 1, 0], dtype=np.int64).ctypes.data_as(ctypes.c_void_p)) == 2
    assert lib.qf_count_inbag(7, None) == 7


def _chain_tree(depth, left_heavy):
    """Internal nodes 2i with one leaf child and one chain child."""
    n = 2 * depth + 1
    left = np.full(n, -1, dtype=np.int64)
    right = np.full(n, -1, dtype=np.int64)
    for i in range(depth):
        if left_heavy:
            left[2 * i], right[2 * i] = 2 * i + 2, 2 * i + 1
        else:
            left[2 * i], right[2 * i] = 2 * i + 1, 2 * i + 2
    feat = np.where(left >= 0, 0, -2).astype(np.int64)
    thr = np.where(left >= 0, 0.5, -2.0)
    return n, left, right, feat, thr


def _complete_tree(depth):
    """Complete binary tree in heap numbering (node k has children 2k+1, 2k+2)."""
    n = 2 ** (depth + 1) - 1
    left = np.full(n, -1, dtype=np.int64)
    right = np.full(n, -1, dtype=np.int64)
    inner = np.arange(2 ** depth - 1)
    left[inner], right[inner] = 2 * inner + 1, 2 * inner + 2
    feat = np.where(left >= 0, 0, -2).astype(np.int64)
    thr = np.where(left >= 0, 0.5, -2.0)
    return n, left, right, feat, thr


def test_offset_limit_is_reported():
    # feature_bits = 24 leaves 7 bits for the child offset (max 127).  In level order the children
    # of a node lie about one level width behind it: chains of any depth pack fine (offsets of 1
    # or 2), a complete tree of depth 9 has levels of 256 nodes -> QF_ERR_LIMIT (never a silently
    # truncated offset).
    lib = _native.lib()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    cases = [(_chain_tree(100, False), _native.QF_OK), (_chain_tree(100, True), _native.QF_OK),
             (_complete_tree(6), _native.QF_OK), (_complete_tree(9), _native.QF_ERR_LIMIT)]
    for (n, left, right, feat, thr), expect in cases:
        leaves = np.full(5, n - 1, dtype=np.int64)
        sorter = np.arange(5, dtype=np.int64)
        nodes, ids, mem = np.zeros(n + 1, np.uint64), np.zeros(n + 1, np.int32), np.zeros(5, np.uint64)
        rc = lib.qf_pack_tree(n, p(left), p(right), p(feat), p(thr), 24, 5, p(leaves), None, p(sorter),
                              0, p(nodes), p(ids), p(mem), None, None, None, None)
        assert rc == expect, _native.last_error()
    rc = lib.qf_pack_tree(n, p(left), p(right), p(feat), p(thr), 25, 5, p(leaves), None, p(sorter), 0,
                          p(nodes), p(ids), p(mem), None, None, None, None)
    assert rc == _native.QF_ERR_INVALID          # feature_bits out of range


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_device_calls_fail_loudly_without_gpu():
    lib = _native.lib()
    desc = _native.ForestDesc()
    arr = np.zeros(4, dtype=np.uint64)
    desc.device, desc.n_trees, desc.n_features, desc.feature_bits = 0, 1, 1, 1
    desc.n_train, desc.n_nodes, desc.n_members = 1, 1, 1
    desc.nodes = desc.tree_offsets = desc.members = desc.y_sorted = arr.ctypes.data
    h = ctypes.c_void_p()
    rc = lib.qf_forest_create(ctypes.byref(desc), ctypes.byref(h))
    assert rc == _native.QF_ERR_CUDA and not h.value
    import skgarden_b200 as sg
    with pytest.raises(RuntimeError):
        sg.DeviceForest(None)
