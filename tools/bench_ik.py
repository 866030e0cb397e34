#!/usr/bin/env python
"""Dev tool: throughput of the batched local IK (rdb_local_ik_batch) for C6: N targets = FK of random joint vectors, seeds 0.3 rad away.
  python tools/bench_ik.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rosdyn_b200 import fixtures  # noqa: E402
from rosdyn_b200.chain import Chain, fill_uniform  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = fixtures.by_name("c6")
ch = Chain(d)
q = fill_uniform(6, N, 7, 0, device="cuda")
T = ch.kinematics(q, want=("T_tool",))["T_tool"]
seed = q + 0.3 * fill_uniform(6, N, 8, 1, device="cuda")
lim = np.full(6, 2 * np.pi)
for _ in range(2):
    sol, ok, it, err = ch.computeLocalIk(T, seed, -lim, lim, toll=1e-8, max_iter=30)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sol, ok, it, err = ch.computeLocalIk(T, seed, -lim, lim, toll=1e-8, max_iter=30)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"local IK, C6, {N} targets: {ms:.2f} ms = {N / ms / 1e3:.2f} M targets/s; converged {float(ok.float().mean()):.4f}, "
      f"mean iterations {float(it.float().mean()):.2f}, {N * float(it.float().mean()) / ms / 1e3:.1f} M Gauss-Newton steps/s")
