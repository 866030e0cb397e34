"""CPU-only, world_size 2 over gloo: the N>1 path of the bench (contiguous sample shards + one all-reduce of the
normal-equation partials) gives the single-rank result.  The per-rank partials come from the oracle here
(no GPU in this tier); the GPU tier checks the fused kernel against the same oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rosdyn_b200 import fixtures
from rosdyn_b200.sharding import allreduce_normal_equations, shard_range


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import OracleChain, fill_uniform
        d = fixtures.by_name("c6")
        oc = OracleChain(d)
        xs = [fill_uniform(6, n, 0x5EED0000 + 4, s) for s in range(3)]
        lo, hi = shard_range(n, rank, world)
        G, b, tt = oc.gram(*[np.ascontiguousarray(x[:, lo:hi]) for x in xs])
        Gt, bt, ttt = allreduce_normal_equations(torch.from_numpy(G), torch.from_numpy(b), torch.tensor([tt], dtype=torch.float64))
        if rank == 0:
            q.put((Gt.numpy().copy(), bt.numpy().copy(), float(ttt[0])))
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_rank():
    from oracle.oracle import OracleChain, fill_uniform
    n, world = 501, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    G, b, tt = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oc = OracleChain(fixtures.by_name("c6"))
    Gr, br, ttr = oc.gram(*[fill_uniform(6, n, 0x5EED0000 + 4, s) for s in range(3)])
    assert np.max(np.abs(G - Gr)) <= 1e-12 * np.max(np.abs(Gr))
    assert np.max(np.abs(b - br)) <= 1e-12 * np.max(np.abs(br))
    assert abs(tt - ttr) <= 1e-12 * ttr
