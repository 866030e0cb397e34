"""Multi-GPU plumbing of the hot path: samples are independent, so they shard contiguously by rank with no data-path
collective; the only exchange is the sum of the small normal-equation partials (P^2 + P + 1 doubles) after the fused
regressor->Gram kernel.  torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank` out of `world`: [r*n/R, (r+1)*n/R)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n) // world, ((rank + 1) * n) // world


def pack_normal_equations(G, b, tau_sq, out=None):
    """Flatten (G[P,P], b[P], tau_sq[1]) into one buffer of P*P+P+1 doubles so a single all-reduce moves them."""
    import torch
    P = b.shape[0]
    if out is None:
        out = torch.empty(P * P + P + 1, dtype=torch.float64, device=b.device)
    out[:P * P].copy_(G.reshape(-1))
    out[P * P:P * P + P].copy_(b)
    out[P * P + P:].copy_(tau_sq.reshape(-1))
    return out


def unpack_normal_equations(flat, P: int):
    return flat[:P * P].reshape(P, P), flat[P * P:P * P + P], flat[P * P + P:]


def allreduce_normal_equations(G, b, tau_sq, group=None, flat=None):
    """Sum the per-rank partial normal equations over all ranks (one all-reduce of P^2+P+1 doubles)."""
    import torch.distributed as dist
    P = b.shape[0]
    flat = pack_normal_equations(G, b, tau_sq, flat)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return unpack_normal_equations(flat, P)


def sharded_gram_host(chains, q, dq, ddq, tau_meas=None):
    """One process, several handles / GPUs (rdb_regressor_gram_sharded_host): the HOST batch q/dq/ddq [n_act][N] (numpy, ideally pinned)
    is cut into len(chains) contiguous shards, every handle runs its shard through its own host pipeline concurrently, and the partial
    normal equations are summed on the host in rank order.  Returns (G[P,P], b[P], tau_sq) as numpy."""
    import ctypes

    import numpy as np

    from ._lib import CSamples, check, load
    lib = load()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, dq, ddq)]
    n_in, n = arrs[0].shape
    tm = None if tau_meas is None else np.ascontiguousarray(tau_meas, dtype=np.float64)
    P = 10 * chains[0].nJ
    G, b, tt = np.zeros((P, P)), np.zeros(P), np.zeros(1)
    hs = (ctypes.c_void_p * len(chains))(*[c._h for c in chains])
    vp = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    s = CSamples(n, n, vp(arrs[0]), vp(arrs[1]), vp(arrs[2]), None)
    check(lib.rdb_regressor_gram_sharded_host(hs, len(chains), ctypes.byref(s), vp(tm), vp(G), vp(b), vp(tt), 0))
    return G, b, float(tt[0])
