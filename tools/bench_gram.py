#!/usr/bin/env python
"""Dev tool: times the fused Gram entry (device-resident inputs) for C6 and C7.   python tools/bench_gram.py [S] [steps] [--lib path]
--lib selects an experimental build (tools/build_variant.sh); builds made with -DRDB_DEV_SWITCHES honour RDB_GRAM_DEBUG=1|2
(skip generation | MMA: timing experiments only -- the shipped library has no such switch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from rosdyn_b200 import _lib, fixtures  # noqa: E402

LIB = "default"
if "--lib" in sys.argv:
    k = sys.argv.index("--lib")
    LIB = sys.argv[k + 1]
    _lib.set_library_path(os.path.abspath(LIB))
    del sys.argv[k:k + 2]
from rosdyn_b200.chain import Chain, fill_uniform  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for name, flop in (("c6", 30660), ("c7", 35770)):
    d = fixtures.by_name(name)
    ch = Chain(d)
    n_in = d.n_in if hasattr(d, "n_in") else (6 if name == "c6" else 7)
    q, dq, ddq = (fill_uniform(n_in, S, 0x5EED0000, s, device="cuda") for s in range(3))
    for _ in range(3):
        ch.regressorGram(q, dq, ddq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ch.regressorGram(q, dq, ddq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{LIB:40s} dbg={os.environ.get('RDB_GRAM_DEBUG', '0')} {name}: {ms:8.3f} ms  "
          f"{S / ms / 1e6:7.4f} G samples/s  {S * flop / ms / 1e9:6.2f} TFLOP/s (algorithmic)", flush=True)
