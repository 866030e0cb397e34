"""Generates tests/golden/ref_<chain>.npz by RUNNING THE REFERENCE ITSELF: oracle/_ref/librosdyn_ref.so is the reference's own
rosdyn::Chain (rosdyn_core/include/rosdyn_core/*.h, compiled where it lies under /root/reference against the stand-in third-party
headers of oracle/shim/, see oracle/ref_driver.cpp).  /root/reference does not exist on the GPU box, so its outputs are committed
here as fixtures.  Run from the repo root in the build container:   python tests/golden/make_golden_ref.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from rosdyn_b200 import fixtures  # noqa: E402

CHAINS = ["c6", "c7", "c6_perturbed", "c7_perturbed", "random_a", "random_b", "random_c", "random_d"]
N = 32
KIN = ("T_links", "jacobian", "twist", "dtwist", "dtwist_lin", "dtwist_nonlin", "ddtwist", "ddtwist_lin", "ddtwist_nonlin", "torque")


def inputs(n_in, seed):
    q, dq, ddq, dddq = (oracle.fill_uniform(n_in, N, seed, s) for s in range(4))
    q *= np.pi                       # the whole circle, not only (-1, 1) rad
    q[:, 0] = 0.0                    # edge cases the domain has: zero pose, +-pi, zero velocity / acceleration / jerk, large angles
    q[:, 1] = np.pi
    q[:, 2] = -np.pi
    dq[:, 3] = 0.0
    ddq[:, 4] = 0.0
    dq[:, 5] = 0.0
    ddq[:, 5] = 0.0
    dddq[:, 5] = 0.0
    q[:, 6] *= 1.0e3
    dq[:, 7] *= 50.0
    ddq[:, 8] *= 500.0
    return q, dq, ddq, dddq


def main():
    assert oracle.build_ref(force=True), "needs /root/reference (run in the build container)"
    for ci, name in enumerate(CHAINS):
        d = fixtures.by_name(name)
        rc = oracle.OracleChain(d, fast="ref")
        q, dq, ddq, dddq = inputs(d.n_inputs, 0x4EF00000 + ci)
        arrays = dict(rc.kinematics(q, dq, ddq, dddq, want=KIN))
        phi, tau = rc.regressor_torque(q, dq, ddq)
        arrays["regressor"] = phi
        np.testing.assert_array_equal(tau, arrays["torque"])
        arrays["torque_nonlin"] = rc.kinematics(q, dq, None, want=("torque",))["torque"]   # getJointTorqueNonLinearPart == DDq = 0
        arrays["inertia"] = rc.inertia(q)
        arrays.update(q=q, dq=dq, ddq=ddq, dddq=dddq, nominal=rc.nominal_parameters())
        path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, arrays["regressor"].shape, os.path.getsize(path))


if __name__ == "__main__":
    main()
