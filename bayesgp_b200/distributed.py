"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous, NCCL inside libbgp.

The reference has no distributed path; SURVEY.md section 8e defines the three shardings used here:
  * observation shards  — rank g holds rows [lo, hi) of the data; every Newton iteration all-reduces
    [g_lik | ll | sumsq | flag] (p + 4 doubles) and the p x p likelihood Hessian (NCCL, on the
    library's stream); prior / Cholesky / step run replicated, so ranks stay bit-identical;
  * node shards         — the ranks of a node group (bgp_model_set_node_group) hold the same rows and each
    evaluates a contiguous run of the quadrature nodes inside bgp_aghq_fit*; the K values are all-reduced, modes
    and Hessians stay on the device that produced them; bgp_sample* draws each node's block on its owner and
    all-reduces the p x M matrix (one owner per column, zeros elsewhere);
  * grid-row shards     — bgp_fit_predict_* splits the rows of x_new over the node group the same way.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Rows [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def node_slice(K: int, rank: int, world: int):
    """Contiguous, balanced runs of K pieces over the ranks of a node group — how the library deals quadrature nodes
    (runs of the expand.grid order: neighbours along the first coordinate share a rank, which keeps the warm starts
    close) and prediction rows (piece_owner / piece_bounds in csrc/bgp_internal.h).  The owners a fit actually used
    are reported by ``AGHQ.node_owner`` (bgp_fit_node_owner)."""
    lo, hi = shard_bounds(K, rank, world)
    return list(range(lo, hi))


def broadcast_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the 128-byte ncclUniqueId (`make_id()`), everyone receives it through
    torch.distributed (works over gloo or nccl)."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().tolist())


def nccl_unique_id() -> bytes:
    import ctypes as C
    from . import _lib
    lib = _lib.load()
    buf = C.create_string_buffer(128)
    _lib.check(lib.bgp_nccl_unique_id(buf))
    return buf.raw


def allgather_nodes(values: np.ndarray, K: int, rank: int, world: int, group=None) -> np.ndarray:
    """Gather per-node results (first axis = this rank's node_slice order) back into node order."""
    import torch
    import torch.distributed as dist
    mine = node_slice(K, rank, world)
    per = (K + world - 1) // world
    trail = values.shape[1:]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    pad = np.zeros((per,) + trail)
    pad[:len(mine)] = values
    t = torch.from_numpy(pad).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    res = np.empty((K,) + trail)
    for r in range(world):
        idx = node_slice(K, r, world)
        res[idx] = out[r].cpu().numpy()[:len(idx)]
    return res
