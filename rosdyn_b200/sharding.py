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
    if not chains:
        raise ValueError("no chain handles")
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (q, dq, ddq)]
    tm = None if tau_meas is None else np.ascontiguousarray(tau_meas, dtype=np.float64)
    if arrs[0].ndim != 2 or arrs[0].shape[0] != chains[0].n_in or any(a.shape != arrs[0].shape for a in arrs[1:]) \
            or (tm is not None and tm.shape != arrs[0].shape):
        raise ValueError("Input data dimensions mismatch")   # every array is n_act x N: a shorter one would be read past its end
    n_in, n = arrs[0].shape
    P = 10 * chains[0].nJ
    G, b, tt = np.zeros((P, P)), np.zeros(P), np.zeros(1)
    hs = (ctypes.c_void_p * len(chains))(*[c._h for c in chains])
    vp = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    s = CSamples(n, n, vp(arrs[0]), vp(arrs[1]), vp(arrs[2]), None)
    check(lib.rdb_regressor_gram_sharded_host(hs, len(chains), ctypes.byref(s), vp(tm), vp(G), vp(b), vp(tt), 0))
    return G, b, float(tt[0])


class Group:
    """The C-ABI's NCCL group (rdb_group_create / rdb_group_create_rank, include/rosdyn_b200.h): device-resident sample shards, the fused
    regressor -> normal-equation kernel on every device and ONE ncclAllReduce of the packed partials over NVLink, all below the C-ABI (torch only
    supplies the device buffers here).

    Group(desc, devices=[0, 1, ...])                   one process drives several GPUs (ncclCommInitAll)
    Group.from_torch_distributed(desc, local_device)   one process per GPU under torchrun: the NCCL unique id travels through torch.distributed
    """

    def __init__(self, desc, devices=None, *, _rank_args=None):
        import ctypes

        from ._lib import check, load
        from .descriptor import to_ctypes
        self._lib = load()
        self._h = ctypes.c_void_p()
        self.desc = desc
        cdesc, keep = to_ctypes(desc)
        if _rank_args is None:
            devices = list(devices) if devices is not None else [0]
            ids = (ctypes.c_int32 * len(devices))(*devices)
            check(self._lib.rdb_group_create(ctypes.byref(cdesc), len(devices), ids, ctypes.byref(self._h)))
            self.devices = devices
        else:
            device, nranks, rank, uid = _rank_args
            buf = (ctypes.c_uint8 * 128)(*uid) if uid is not None else None
            check(self._lib.rdb_group_create_rank(ctypes.byref(cdesc), int(device), int(nranks), int(rank), buf, ctypes.byref(self._h)))
            self.devices = [int(device)]
        del keep
        self.P = 10 * desc.n_joints
        self.ranks = int(self._lib.rdb_group_ranks(self._h))

    @classmethod
    def from_torch_distributed(cls, desc, device: int, group=None):
        """One rank per process: rank 0 draws the NCCL unique id (rdb_group_unique_id) and torch.distributed carries its 128 bytes to the others."""
        import ctypes

        import torch
        import torch.distributed as dist

        from ._lib import check, load
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        uid = None
        if world > 1:
            lib = load()
            buf = (ctypes.c_uint8 * 128)()
            if rank == 0:
                check(lib.rdb_group_unique_id(buf))
            t = torch.tensor(list(buf), dtype=torch.uint8)
            if dist.get_backend(group) == "nccl":
                t = t.to(torch.device("cuda", device))
            dist.broadcast(t, src=0, group=group)
            uid = [int(v) for v in t.cpu().tolist()]
        return cls(desc, _rank_args=(device, world, rank, uid))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.rdb_group_destroy(h)
            self._h = None

    def gram(self, shards, tau_meas=None, out=None, streams=None, accumulate=False):
        """shards: one (q, Dq, DDq) triple of torch CUDA tensors [n_act][N_k] per local device (on that device).  Returns a list of
        (G[P,P], b[P], tau_sq[1]) per local device, every one holding the sum over ALL ranks.  Asynchronous on `streams` (default: torch's
        current stream of each device)."""
        import ctypes

        import torch

        from ._lib import CSamples, check
        nd = len(self.devices)
        if len(shards) != nd:
            raise ValueError("one shard per local device")
        smp = (CSamples * nd)()
        keep = []
        tau_p = (ctypes.c_void_p * nd)()
        G_p, b_p, t_p, st_p = ((ctypes.c_void_p * nd)() for _ in range(4))
        res = []
        for k, (q, dq, ddq) in enumerate(shards):
            dev = torch.device("cuda", self.devices[k])
            arrs = [a.to(torch.float64).contiguous() for a in (q, dq, ddq)]
            if any(a.device != dev for a in arrs) or any(a.shape != arrs[0].shape for a in arrs) or arrs[0].dim() != 2:
                raise ValueError("Input data dimensions mismatch")
            n = arrs[0].shape[1]
            smp[k] = CSamples(n, max(n, 1), arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(), None)
            keep.append(arrs)
            if tau_meas is not None and tau_meas[k] is not None:
                tm = tau_meas[k].to(torch.float64).contiguous()
                if tm.shape != arrs[0].shape:
                    raise ValueError("Input data dimensions mismatch")
                keep.append(tm)
                tau_p[k] = tm.data_ptr()
            if out is not None:
                G, b, tt = out[k]
            else:
                G = torch.empty((self.P, self.P), dtype=torch.float64, device=dev)
                b = torch.empty((self.P,), dtype=torch.float64, device=dev)
                tt = torch.empty((1,), dtype=torch.float64, device=dev)
            res.append((G, b, tt))
            G_p[k], b_p[k], t_p[k] = G.data_ptr(), b.data_ptr(), tt.data_ptr()
            st_p[k] = streams[k] if streams is not None else torch.cuda.current_stream(dev).cuda_stream
        check(self._lib.rdb_regressor_gram_sharded(self._h, smp, ctypes.cast(tau_p, ctypes.POINTER(ctypes.c_void_p)),
                                                   ctypes.cast(G_p, ctypes.POINTER(ctypes.c_void_p)), ctypes.cast(b_p, ctypes.POINTER(ctypes.c_void_p)),
                                                   ctypes.cast(t_p, ctypes.POINTER(ctypes.c_void_p)), int(accumulate or out is not None and accumulate),
                                                   ctypes.cast(st_p, ctypes.POINTER(ctypes.c_void_p))))
        self._keep = keep  # inputs must outlive the asynchronous launches
        return res
