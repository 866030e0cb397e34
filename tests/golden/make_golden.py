"""Generates tests/golden/<chain>.npz from the independent numpy transcription of the reference
(oracle/numpy_transcription.py).  Run from the repo root:  python tests/golden/make_golden.py
The reference holds no golden vectors of its own; these pin the C oracle and the CUDA engine
to a second, separately written restatement of the same reference lines."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import numpy_transcription as nt  # noqa: E402
from rosdyn_b200 import fixtures  # noqa: E402

CHAINS = ["c6", "c7", "c6_perturbed", "c7_perturbed", "random_a", "random_b", "random_c", "random_d"]
N = 12


def inputs(n_in, seed):
    q, dq, ddq, dddq = (nt.fill_uniform(n_in, N, seed, s) for s in range(4))
    # edge cases the domain has: zero pose, +-pi, zero velocity / acceleration
    q[:, 0] = 0.0
    q[:, 1] = np.pi
    q[:, 2] = -np.pi
    dq[:, 3] = 0.0
    ddq[:, 4] = 0.0
    dq[:, 5] = 0.0
    ddq[:, 5] = 0.0
    dddq[:, 5] = 0.0
    return q, dq, ddq, dddq


def main():
    for ci, name in enumerate(CHAINS):
        d = fixtures.by_name(name)
        c = nt.NpChain(d)
        q, dq, ddq, dddq = inputs(d.n_inputs, 0x5EED0000 + ci)
        out = {k: [] for k in ("T_links", "jacobian", "twist", "dtwist", "dtwist_lin", "dtwist_nonlin", "ddtwist", "ddtwist_lin",
                               "ddtwist_nonlin", "torque", "torque_nonlin", "regressor", "inertia")}
        for i in range(N):
            a = (q[:, i], dq[:, i], ddq[:, i], dddq[:, i])
            out["T_links"].append(np.concatenate([t[:3, :].reshape(-1) for t in c.getTransformations(a[0])]))
            out["jacobian"].append(c.getJacobian(a[0]).T.reshape(-1))          # plane col*6+row
            out["twist"].append(np.concatenate(c.getTwist(a[0], a[1])))
            out["dtwist"].append(np.concatenate(c.getDTwist(a[0], a[1], a[2])))
            out["dtwist_lin"].append(np.concatenate(c.getDTwistLinearPart(a[0], a[2])))
            out["dtwist_nonlin"].append(np.concatenate(c.getDTwistNonLinearPart(a[0], a[1])))
            out["ddtwist"].append(np.concatenate(c.getDDTwist(*a)))
            out["ddtwist_lin"].append(np.concatenate(c.getDDTwistLinearPart(a[0], a[3])))
            out["ddtwist_nonlin"].append(np.concatenate(c.getDDTwistNonLinearPart(a[0], a[1], a[2])))
            out["torque"].append(c.getJointTorque(a[0], a[1], a[2]))
            out["torque_nonlin"].append(c.getJointTorqueNonLinearPart(a[0], a[1]))
            out["regressor"].append(c.getRegressor(a[0], a[1], a[2]).T.reshape(-1))  # plane col*n_in+row
            out["inertia"].append(c.getJointInertia(a[0]).T.reshape(-1))             # plane col*n_in+row
        arrays = {k: np.stack(v, axis=1) for k, v in out.items()}                     # [rows][N] planes
        arrays.update(q=q, dq=dq, ddq=ddq, dddq=dddq, nominal=c.getNominalParameters())
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(name, {k: v.shape for k, v in arrays.items() if k in ("regressor", "T_links")}, os.path.getsize(path))


if __name__ == "__main__":
    main()
