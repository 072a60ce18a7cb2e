"""Row sharding over the GPUs of one box: the forest is replicated, test rows are split
into contiguous blocks, and there is no collective on the data path
(skgarden/quantile/ensemble.py:137 -- the per-row loop has no cross-row state).

``broadcast_flat`` ships the flattened forest from rank 0 to every rank once, as raw byte
buffers over ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .flatten import FlatForest

_ARRAYS = ["nodes", "tree_offsets", "node_ids", "members", "member_offsets", "y_sorted",
           "leaf_values", "sorter", "leaf_impurity"]
_SCALARS = ["n_trees", "n_features", "feature_bits", "n_train", "max_row_members",
            "expected_row_members", "max_depth"]


def shard_rows(n_rows: int, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of rank ``rank``: sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_flat(flat, rank: int, device, src: int = 0) -> FlatForest:
    """Rank ``src`` -> all ranks.  ``flat`` is ignored on the other ranks (may be None)."""
    import torch
    import torch.distributed as dist
    meta = [None]
    if rank == src:
        meta[0] = dict(
            scalars={k: getattr(flat, k) for k in _SCALARS},
            arrays={k: (None if getattr(flat, k) is None
                        else (getattr(flat, k).dtype.str, tuple(getattr(flat, k).shape))) for k in _ARRAYS})
    dist.broadcast_object_list(meta, src=src)
    meta = meta[0]
    out = {}
    for k in _ARRAYS:
        spec = meta["arrays"][k]
        if spec is None:
            out[k] = None
            continue
        dtype, shape = np.dtype(spec[0]), tuple(spec[1])
        nbytes = int(np.prod(shape)) * dtype.itemsize
        if rank == src:
            host = np.ascontiguousarray(getattr(flat, k)).view(np.uint8).reshape(-1)
            buf = torch.from_numpy(host).to(device)
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        if nbytes:
            dist.broadcast(buf, src=src)
        out[k] = getattr(flat, k) if rank == src else buf.cpu().numpy().view(dtype).reshape(shape)
        del buf
    if rank == src:
        return flat
    return FlatForest(nodes=out["nodes"], tree_offsets=out["tree_offsets"], node_ids=out["node_ids"],
                      members=out["members"], member_offsets=out["member_offsets"], y_sorted=out["y_sorted"],
                      leaf_values=out["leaf_values"], sorter=out["sorter"],
                      leaf_impurity=out["leaf_impurity"], **meta["scalars"])
