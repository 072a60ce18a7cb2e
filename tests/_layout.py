"""Test-only decoder of the HBM layout (include/qforest_b200.h): walks the packed nodes
and member lists on the CPU exactly as the kernels are specified to, so the flattener
can be checked against the reference's outputs without a GPU.  Never used by the product."""
import numpy as np


def decode_apply(flat, X):
    """(n,T) packed node index of each row's leaf + (start, count) member segments."""
    X = np.asarray(X, dtype=np.float32)
    n, T = X.shape[0], flat.n_trees
    lo = (flat.nodes & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (flat.nodes >> np.uint64(32)).astype(np.uint32)
    fmask = np.uint32((1 << flat.feature_bits) - 1)
    node_idx = np.zeros((n, T), dtype=np.int64)
    for t in range(T):
        node = np.full(n, flat.tree_offsets[t], dtype=np.int64)
        active = (hi[node] & np.uint32(0x80000000)) == 0
        while active.any():
            r = np.flatnonzero(active)
            nd = node[r]
            thr = lo[nd].view(np.float32)                      # floor-f32 threshold
            f = (hi[nd] & fmask).astype(np.int64)
            off = (hi[nd] >> np.uint32(flat.feature_bits)).astype(np.int64)
            node[r] = nd + off + np.where(X[r, f] <= thr, 0, 1)   # float32 compare; right child = left + 1
            active[r] = (hi[node[r]] & np.uint32(0x80000000)) == 0
        node_idx[:, t] = node
    start = lo[node_idx].astype(np.int64)
    count = (hi[node_idx] & np.uint32(0x7FFFFFFF)).astype(np.int64)
    return node_idx, start, count


def sklearn_ids(flat, node_idx):
    if flat.node_ids is None:
        return node_idx - flat.tree_offsets[:-1][None, :]
    return flat.node_ids[node_idx].astype(np.int64)


def decode_predict(flat, start, count, quantiles):
    """Percentiles from the member lists with the float32 op order of the kernels
    (== SURVEY.md Appendix A)."""
    from oracle.qrf_oracle import _percentile_of_sorted
    rank = (flat.members & np.uint64(0xFFFFFFFF)).astype(np.int64)
    w = (flat.members >> np.uint64(32)).astype(np.uint32).view(np.float32)
    n, T = start.shape
    out = np.zeros((n, len(quantiles)))
    for i in range(n):
        idx = np.concatenate([np.arange(start[i, t], start[i, t] + count[i, t]) for t in range(T)])
        r, ww = rank[idx], w[idx]
        uniq, inv = np.unique(r, return_inverse=True)          # rank order
        W = np.zeros(len(uniq), dtype=np.float32)
        np.add.at(W, inv, ww)                                  # sequential, tree order
        keep = W != 0
        a = flat.y_sorted[uniq][keep]
        for k, q in enumerate(quantiles):
            out[i, k] = _percentile_of_sorted(a, W[keep], float(q))
    return out
