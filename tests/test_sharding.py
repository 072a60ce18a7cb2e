"""Multi-GPU host logic on CPU: world_size-2 ``gloo`` processes broadcast the flattened
forest and split the test rows; the shards' results put back together equal the
single-process result (output independent of the shard count)."""
import os
import socket
import sys

import numpy as np
import pytest

import _golden
import _layout

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import _golden as G
    import _layout as L
    from skgarden_b200.sharding import broadcast_flat, shard_rows
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = G.load(name)
    flat = G.flat_from_golden(g) if rank == 0 else None          # only rank 0 holds the fitted forest
    flat = broadcast_flat(flat, rank, "cpu")
    lo, hi = shard_rows(g.X_test.shape[0], world, rank)
    node_idx, start, count = L.decode_apply(flat, g.X_test[lo:hi])   # walk the broadcast layout
    leaves = L.sklearn_ids(flat, node_idx)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, leaves=leaves,
             members=count.sum(axis=1), nodes_sum=np.uint64(flat.nodes.sum()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partition():
    from skgarden_b200.sharding import shard_rows
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            blocks = [shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


@pytest.mark.parametrize("name", ["et_msl5", "rf_maxleaf50"])
def test_two_rank_gloo_shards_match_single_process(name, tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = _golden.load(name)
    flat = _golden.flat_from_golden(g)
    parts = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert [int(p["lo"]) for p in parts] == [0, int(parts[0]["hi"])]
    leaves = np.concatenate([p["leaves"] for p in parts])
    np.testing.assert_array_equal(leaves, g.apply)                        # == the reference's apply
    for p in parts:
        assert int(p["nodes_sum"]) == int(flat.nodes.sum())               # every rank got the same forest
