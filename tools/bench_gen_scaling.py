#!/usr/bin/env python
"""Dev tool (needs a -DRDB_DEV_SWITCHES library and RDB_GRAM_DEBUG=2 = generation only): time per generated group against the size of the
unrolled generator code, for the C6 chain with its first K joints as inputs (the folded chain then has K moving joints)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rosdyn_b200 import _lib, fixtures
if "--lib" in sys.argv:
    k = sys.argv.index("--lib"); _lib.set_library_path(os.path.abspath(sys.argv[k + 1])); del sys.argv[k:k + 2]
import torch
from rosdyn_b200.chain import Chain, fill_uniform
S = 8_000_000
for K in range(1, 7):
    d = fixtures.by_name("c6"); ch = Chain(d)
    names = ch.getActiveJointsName()
    ch.setInputJointsName(names[:K])
    q, dq, ddq = (fill_uniform(K, S, 3, s, device="cuda") for s in range(3))
    for _ in range(2): ch.regressorGram(q, dq, ddq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ch.regressorGram(q, dq, ddq)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    groups_per_warp = S / 32 / (148 * 4)
    print(f"dbg={os.environ.get('RDB_GRAM_DEBUG','0')} K={K}: {ms:7.3f} ms  cycles per group per generator warp {ms * 1e-3 * 1.965e9 / groups_per_warp:9.0f}", flush=True)
