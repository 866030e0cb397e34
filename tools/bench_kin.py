#!/usr/bin/env python
"""Dev tool: config 2 / config 5b kinematics kernels, A/B between library builds.  python tools/bench_kin.py [--lib path]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rosdyn_b200 import _lib, fixtures
LIB = "default"
if "--lib" in sys.argv:
    k = sys.argv.index("--lib"); LIB = sys.argv[k + 1]; _lib.set_library_path(os.path.abspath(LIB)); del sys.argv[k:k + 2]
import torch
from rosdyn_b200._lib import CKinematicsOut, CSamples, check, load
from rosdyn_b200.chain import Chain, fill_uniform
lib = load()
d = fixtures.by_name("c6"); ch = Chain(d)
S, n_in, nL = 16_000_000, 6, 8
q, dq, ddq, dddq = (fill_uniform(n_in, S, 7, s, device="cuda") for s in range(4))
def run(fields, smp, label, bytes_per):
    outs = {k: torch.empty((r, S), dtype=torch.float64, device="cuda") for k, r in fields}
    ko = CKinematicsOut(); ko.ld = S
    for k, v in outs.items(): setattr(ko, k, v.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f = lambda: check(lib.rdb_kinematics_batch(ch._h, ctypes.byref(smp), ctypes.byref(ko), st))
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{LIB:36s} {label}: {ms:7.3f} ms  {S / ms / 1e6:6.3f} G samples/s  {S * bytes_per / ms / 1e6:7.1f} GB/s", flush=True)
run((("T_tool", 12), ("jacobian", 36), ("twist", 48), ("dtwist", 48), ("torque", 6)), CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None), "config2", 1344)
run((("ddtwist", 48),), CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), dddq.data_ptr()), "config5b jerk", 576)
