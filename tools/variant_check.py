#!/usr/bin/env python
"""Dev tool: normal equations (rigid C6 / C7, extended C6) of a library build, saved for an A/B comparison.
python tools/variant_check.py out.npz [--lib path];   python tools/variant_check.py --compare a.npz b.npz"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    worst = 0.0
    for k in a.files:
        e = float(np.max(np.abs(a[k] - b[k])) / max(np.max(np.abs(a[k])), 1.0))
        worst = max(worst, e)
        print(f"{k:12s} rel diff {e:.2e}  bit-identical {np.array_equal(a[k], b[k])}")
    sys.exit(0 if worst < 1e-12 else 1)
from rosdyn_b200 import _lib, fixtures
if "--lib" in sys.argv:
    k = sys.argv.index("--lib"); _lib.set_library_path(os.path.abspath(sys.argv[k + 1])); del sys.argv[k:k + 2]
import torch
from rosdyn_b200.chain import Chain, fill_uniform
out = {}
for name in ("c6", "c7"):
    d = fixtures.by_name(name); ch = Chain(d)
    S = 1_000_013
    q, dq, ddq = (fill_uniform(d.n_inputs, S, 11, s, device="cuda") for s in range(3))
    G, b, tt = ch.regressorGram(q, dq, ddq)
    out[name + "_G"], out[name + "_b"], out[name + "_tt"] = G.cpu().numpy(), b.cpu().numpy(), tt.cpu().numpy()
    ch.setComponents([{"type": "friction1", "joint": n, "min_velocity": 0.01, "max_velocity": 2.0} for n in ch.getActiveJointsName()])
    G, b, tt = ch.regressorGram(q, dq, ddq)
    out[name + "x_G"], out[name + "x_b"] = G.cpu().numpy(), b.cpu().numpy()
np.savez(sys.argv[1], **out)
print("saved", sys.argv[1])
