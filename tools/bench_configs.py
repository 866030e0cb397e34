#!/usr/bin/env python
"""Measures every BASELINE.json config on one B200 (device-resident inputs, CUDA events, >= 3 warm-ups, working sets >> L2)
and prints one JSON object per config.  bench.py stays the headline line; this is the table behind DESIGN.md section 3.

  python tools/bench_configs.py [--steps 5] > gpurun_out/configs.jsonl
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from rosdyn_b200 import fixtures  # noqa: E402
from rosdyn_b200._lib import CKinematicsOut, CSamples, check, load  # noqa: E402
from rosdyn_b200.chain import Chain, fill_uniform, fp64_peak  # noqa: E402

SEED = 0x5EED0000


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    lib = load()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    fp64 = max(fp64_peak("dmma", 3), fp64_peak("dfma", 3))
    dev = torch.device("cuda", 0)

    def stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def inputs(n_in, S, cfg):
        return [fill_uniform(n_in, S, SEED + cfg, s, device=dev) for s in range(4)]

    def report(name, S, ms, bytes_per_sample=None, flop_per_sample=None, note=""):
        r = {"config": name, "samples_per_launch": S, "ms": ms, "samples_per_s": S / (ms * 1e-3), "note": note}
        if bytes_per_sample:
            gbs = S * bytes_per_sample / (ms * 1e-3) / 1e9
            r.update(bytes_per_sample=bytes_per_sample, achieved_GBps=gbs, hbm_peak_GBps=hbm, hbm_frac=gbs / hbm)
        if flop_per_sample:
            tf = S * flop_per_sample / (ms * 1e-3) / 1e12
            r.update(flop_per_sample=flop_per_sample, achieved_TFLOPs=tf, fp64_peak_TFLOPs=fp64, fp64_frac=tf / fp64)
        print(json.dumps(r), flush=True)

    # ---- config 2: C6, pose + Jacobian + twists + acceleration twists (all links) + RNEA torque
    d = fixtures.by_name("c6")
    ch = Chain(d)
    n_in, nL, S = 6, 8, 8_000_000
    q, dq, ddq, dddq = inputs(n_in, S, 2)
    outs = {k: torch.empty((r, S), dtype=torch.float64, device=dev) for k, r in
            (("T_tool", 12), ("jacobian", 6 * n_in), ("twist", 6 * nL), ("dtwist", 6 * nL), ("torque", n_in))}
    ko = CKinematicsOut()
    ko.ld = S
    for k, v in outs.items():
        setattr(ko, k, v.data_ptr())
    smp = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)
    ms = timed(lambda: check(lib.rdb_kinematics_batch(ch._h, ctypes.byref(smp), ctypes.byref(ko), stream())), args.steps)
    report("config2: C6 pose+Jacobian+twist+dtwist(all links)+torque [kin_kernel<7,CFG2>]", S, ms, bytes_per_sample=8 * (18 + 150))
    del outs

    # ---- torque only (FP64-pipe bound)
    tau = torch.empty((n_in, S), dtype=torch.float64, device=dev)
    ms = timed(lambda: check(lib.rdb_torque_batch(ch._h, ctypes.byref(smp), tau.data_ptr(), S, stream())), args.steps)
    report("C6 getJointTorque only [dyn_kernel<7,TORQUE>]", S, ms, bytes_per_sample=8 * (18 + 6), note="FP64-pipe bound, not HBM")

    # ---- config 5: C6 getJointInertia + getDDTwist (two launches)
    M = torch.empty((n_in * n_in, S), dtype=torch.float64, device=dev)
    jerk = torch.empty((6 * nL, S), dtype=torch.float64, device=dev)
    ko5 = CKinematicsOut()
    ko5.ld = S
    ko5.ddtwist = jerk.data_ptr()
    smp5 = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), dddq.data_ptr())
    ms_i = timed(lambda: check(lib.rdb_inertia_batch(ch._h, ctypes.byref(smp5), M.data_ptr(), S, stream())), args.steps)
    ms_j = timed(lambda: check(lib.rdb_kinematics_batch(ch._h, ctypes.byref(smp5), ctypes.byref(ko5), stream())), args.steps)
    report("config5a: C6 getJointInertia [dyn_kernel<7,INERTIA>]", S, ms_i, bytes_per_sample=8 * (6 + 36), note="FP64-pipe bound at this byte count")
    report("config5b: C6 getDDTwist all links [kin_kernel<7,JERK>]", S, ms_j, bytes_per_sample=8 * (24 + 48))
    report("config5: inertia + jerk twists (sum of the two launches)", S, ms_i + ms_j, bytes_per_sample=8 * (6 + 36) + 8 * (24 + 48))
    del M, jerk, tau

    # ---- headline chain C6: materialised regressor + torque, and the fused Gram
    S = 4_000_000
    smp = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)
    phi = torch.empty((70 * n_in + n_in, S), dtype=torch.float64, device=dev)
    ms = timed(lambda: check(lib.rdb_regressor_batch(ch._h, ctypes.byref(smp), phi.data_ptr(), phi[70 * n_in:].data_ptr(), S, stream())), args.steps)
    report("headline (materialised): C6 getRegressor 6x70 + torque [dyn_kernel<7,REGRESSOR|TORQUE>]", S, ms, bytes_per_sample=8 * (18 + 420 + 6))
    del phi
    S = 8_000_000
    smp = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)
    G = torch.empty((70, 70), dtype=torch.float64, device=dev)
    b = torch.empty((70,), dtype=torch.float64, device=dev)
    tt = torch.empty((1,), dtype=torch.float64, device=dev)
    ms = timed(lambda: check(lib.rdb_regressor_gram_batch(ch._h, ctypes.byref(smp), None, G.data_ptr(), b.data_ptr(), tt.data_ptr(), 0, stream())), args.steps)
    report("headline (fused Gram): C6 regressor+torque -> PhiT Phi / PhiT tau [gram_fused_kernel<7>]", S, ms, flop_per_sample=6 * 70 * 71 + 2 * 6 * 70,
           note="flops: BLAS SYRK+GEMV convention; executed (folded chain, 6 moving joints): 13952 DMMA flop/sample + ~5 k generation")
    del q, dq, ddq, dddq

    # ---- config 3 / 4: C7
    d7 = fixtures.by_name("c7")
    ch7 = Chain(d7)
    n_in, S = 7, 4_000_000
    q, dq, ddq, _ = inputs(n_in, S, 3)
    smp = CSamples(S, S, q.data_ptr(), dq.data_ptr(), ddq.data_ptr(), None)
    phi = torch.empty((70 * n_in, S), dtype=torch.float64, device=dev)
    ms = timed(lambda: check(lib.rdb_regressor_batch(ch7._h, ctypes.byref(smp), phi.data_ptr(), None, S, stream())), args.steps)
    report("config3: C7 materialised regressor 7x70 [dyn_kernel<7,REGRESSOR>]", S, ms, bytes_per_sample=8 * (21 + 490))
    del phi
    ms = timed(lambda: check(lib.rdb_regressor_gram_batch(ch7._h, ctypes.byref(smp), None, G.data_ptr(), b.data_ptr(), tt.data_ptr(), 0, stream())), args.steps)
    report("config4 (1 GPU): C7 fused regressor -> Gram [gram_fused_kernel<7>]", S, ms, flop_per_sample=7 * 70 * 71 + 2 * 7 * 70)


if __name__ == "__main__":
    main()
