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
