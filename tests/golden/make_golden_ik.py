"""Generates tests/golden/ref_ik_<chain>.npz by RUNNING THE REFERENCE'S OWN Chain::computeLocalIk / getMultiplicity (oracle/_ref: the
reference's headers compiled where they lie; tick clock instead of the wall clock; box-QP stand-in for the un-vendored solve_quadprog,
see oracle/ref_driver.cpp and DESIGN.md section 3.4).  /root/reference does not exist on the GPU box, so the outputs are committed.
Run from the repo root in the build container:   python tests/golden/make_golden_ik.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from oracle.oracle import OracleChain  # noqa: E402
from rosdyn_b200 import fixtures  # noqa: E402

CHAINS = ["c6", "c6_perturbed", "random_b", "random_d"]     # non-redundant chains: J^T J regular, the QP minimiser is unique
N, TOLL, MAX_ITER = 96, 1e-8, 25


def main():
    assert oracle.build_ref(force=True), "needs /root/reference (run in the build container)"
    for ci, name in enumerate(CHAINS):
        d = fixtures.by_name(name)
        rc = OracleChain(d, fast="ref")
        rng = np.random.RandomState(100 + ci)
        q_goal = rng.uniform(-1.0, 1.0, (d.n_inputs, N))
        target = rc.kinematics(q_goal, want=("T_tool",))["T_tool"]      # the reference's own forward kinematics
        seed = q_goal + rng.uniform(-0.35, 0.35, q_goal.shape)
        q_min, q_max = np.full(d.n_inputs, -1.2), np.full(d.n_inputs, 1.2)
        sol, status, iters, _ = rc.local_ik(target, seed, q_min, q_max, toll=TOLL, max_iter=MAX_ITER)
        # getMultiplicity of the first goal inside +-7 rad
        types = [0] * d.n_inputs
        for j in d.joints:
            if j.input_index >= 0:
                types[j.input_index] = int(j.type)
        dp = ctypes.POINTER(ctypes.c_double)
        f = rc._l.lib.oracle_multiplicity
        f.restype = ctypes.c_int64
        f.argtypes = [ctypes.c_void_p, dp, dp, dp, dp, ctypes.c_int64]
        lim = np.full(d.n_inputs, 7.0)
        q0 = np.ascontiguousarray(q_goal[:, 0])
        buf = np.zeros((1 << 14, d.n_inputs))
        cnt = f(rc._h, q0.ctypes.data_as(dp), (-lim).ctypes.data_as(dp), lim.ctypes.data_as(dp), buf.ctypes.data_as(dp), buf.shape[0])
        assert cnt <= buf.shape[0]
        out = os.path.join(ROOT, "tests", "golden", f"ref_ik_{name}.npz")
        np.savez_compressed(out, target=target, seed=seed, q_min=q_min, q_max=q_max, toll=TOLL, max_iter=MAX_ITER, sol=sol, status=status,
                            iters=iters, joint_types=np.array(types, dtype=np.int32), mult_q=q0, mult_lim=lim, mult=buf[:cnt].copy())
        print(name, "converged", float(status.mean()), "multiplicity", cnt, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
