"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU paths (row / node partitioning,
unique-id broadcast, node-result gather).  The CUDA/NCCL side is covered by scripts/mgpu_check.py."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_rows_exactly():
    from bayesgp_b200.distributed import node_slice, shard_bounds
    for n in (1, 7, 10, 1_000_003):
        for world in (1, 2, 3, 8):
            segs = [shard_bounds(n, r, world) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == n
            assert all(segs[i][1] == segs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in segs]
            assert max(sizes) - min(sizes) <= 1
    for K, world in ((15, 2), (49, 8), (125, 8), (4, 8)):
        got = sorted(sum((node_slice(K, r, world) for r in range(world)), []))
        assert got == list(range(K))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        from bayesgp_b200.distributed import allgather_nodes, broadcast_unique_id, node_slice, shard_bounds
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        uid = broadcast_unique_id(lambda: bytes(range(128)), rank)
        K, p = 7, 3
        mine = node_slice(K, rank, world)
        vals = np.array([[100.0 * j + c for c in range(p)] for j in mine]).reshape(len(mine), p)
        full = allgather_nodes(vals, K, rank, world)
        # a sharded "allreduce of the gradient": each rank sums its own rows
        import torch
        x = np.arange(11.0)
        lo, hi = shard_bounds(len(x), rank, world)
        t = torch.tensor([x[lo:hi].sum()], dtype=torch.float64)
        dist.all_reduce(t)
        q.put((rank, uid == bytes(range(128)), full.tolist(), float(t[0])))
        dist.destroy_process_group()
    except Exception as e:     # pragma: no cover
        q.put((rank, "ERR", repr(e), 0.0))


def test_gloo_world2_broadcast_and_gather():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    want = [[100.0 * j + c for c in range(3)] for j in range(7)]
    for rank, ok, full, tot in res:
        assert ok is True, (rank, ok, full)
        assert full == want
        assert tot == 55.0
