"""CPU-only: the C oracle against the committed golden vectors (independent numpy transcription), the
algebraic invariants of SURVEY.md section 4 and the UR10 zero-pose known answer."""
import numpy as np
import pytest

from conftest import CHAINS, assert_close
from oracle import numpy_transcription as nt
from oracle.oracle import OracleChain, fill_uniform
from rosdyn_b200 import fixtures

KIN = ("T_links", "jacobian", "twist", "dtwist", "dtwist_lin", "dtwist_nonlin", "ddtwist", "ddtwist_lin", "ddtwist_nonlin", "torque")


@pytest.mark.parametrize("name", CHAINS)
@pytest.mark.parametrize("fast", [False, True])
def test_oracle_matches_golden(name, fast, golden):
    g = golden(name)
    oc = OracleChain(fixtures.by_name(name), fast=fast)
    K = oc.kinematics(g["q"], g["dq"], g["ddq"], g["dddq"])
    for k in KIN:
        assert_close(K[k], g[k], f"{name}:{k}")
    assert_close(K["T_tool"], g["T_links"][-12:], f"{name}:T_tool")
    phi, tau = oc.regressor_torque(g["q"], g["dq"], g["ddq"])
    assert_close(phi, g["regressor"], f"{name}:regressor")
    assert_close(tau, g["torque"], f"{name}:torque")
    assert_close(oc.kinematics(g["q"], g["dq"], None, want=("torque",))["torque"], g["torque_nonlin"], f"{name}:torque_nonlin")
    assert_close(oc.inertia(g["q"]), g["inertia"], f"{name}:inertia")
    assert_close(oc.nominal_parameters(), g["nominal"], f"{name}:nominal")


@pytest.mark.parametrize("name", CHAINS)
def test_invariants(name):
    d = fixtures.by_name(name)
    oc = OracleChain(d)
    n, n_in, P = 64, d.n_inputs, 10 * d.n_joints
    q, dq, ddq = (fill_uniform(n_in, n, 77, s) for s in range(3))
    phi, tau = oc.regressor_torque(q, dq, ddq)
    Phi = phi.reshape(P, n_in, n)                      # [col][row][sample]
    pi = oc.nominal_parameters()
    assert_close(np.einsum("cri,c->ri", Phi, pi), tau, "Phi * pi_nom == tau", 1e-12)
    M = oc.inertia(q).reshape(n_in, n_in, n)           # [col][row][sample]
    h = oc.kinematics(q, dq, None, want=("torque",))["torque"]
    assert_close(np.einsum("cri,ci->ri", M, ddq) + h, tau, "M ddq + h == tau", 1e-12)
    assert_close(M, np.swapaxes(M, 0, 1), "M == M^T", 1e-12)
    K = oc.kinematics(q, dq, want=("jacobian", "twist"))
    J = K["jacobian"].reshape(n_in, 6, n)              # [col][row][sample]
    assert_close(np.einsum("cri,ci->ri", J, dq), K["twist"][-6:], "J dq == twist_tool", 1e-12)
    # Phi is block upper-triangular: chain joint j has exact zeros in the column blocks of links before it
    for j, jd in enumerate(d.joints):
        if jd.input_index >= 0:
            assert np.all(Phi[:10 * j, jd.input_index, :] == 0.0)


def test_ur10_zero_pose_known_answer():
    oc = OracleChain(fixtures.by_name("c6"))
    T = oc.kinematics(np.zeros((6, 1)), want=("T_tool",))["T_tool"][:, 0].reshape(3, 4)
    np.testing.assert_allclose(T[:, 3], [1.1843, 0.256141, 0.0116], atol=1e-12)


def test_jerk_coefficient_quirk_is_mirrored():
    """The reference's jerk uses 1x (v x s) DDq (primitives_impl.h:1216); the exact derivative needs 2x.  Check the
    oracle keeps the reference's formula: finite differences of dtwist must NOT match ddtwist in general."""
    d = fixtures.by_name("c6")
    oc = OracleChain(d)
    q, dq, ddq, dddq = (fill_uniform(6, 4, 5, s) for s in range(4))
    h = 1e-6
    a0 = oc.kinematics(q, dq, ddq, want=("dtwist",))["dtwist"]
    a1 = oc.kinematics(q + h * dq, dq + h * ddq, ddq + h * dddq, want=("dtwist",))["dtwist"]
    fd = (a1 - a0) / h
    j = oc.kinematics(q, dq, ddq, dddq, want=("ddtwist",))["ddtwist"]
    assert np.max(np.abs(fd[-3:] - j[-3:])) > 1e-3     # angular part of the tool jerk differs (missing 1x (v x s) DDq)


def test_generators_agree():
    a = fill_uniform(6, 100, 0x5EED0001, 2)
    b = nt.fill_uniform(6, 100, 0x5EED0001, 2)
    assert np.array_equal(a, b)
    assert a.min() >= -1.0 and a.max() < 1.0


def test_set_input_joints_permutation_and_subset():
    """setInputJointsName (primitives_impl.h:705-742): permuted / partial input lists; unlisted joints get q = 0."""
    d = fixtures.by_name("c6")
    names = [j.name for j in d.joints if j.type != 0]
    perm = [names[i] for i in (3, 0, 5, 1)]            # subset + permutation
    d2 = fixtures.by_name("c6")
    assert d2.set_input_joints(perm)
    o1, o2 = OracleChain(d), OracleChain(d2)
    n = 10
    q2, dq2, ddq2 = (fill_uniform(4, n, 9, s) for s in range(3))
    q1, dq1, ddq1 = np.zeros((6, n)), np.zeros((6, n)), np.zeros((6, n))
    for k, i in enumerate((3, 0, 5, 1)):
        q1[i], dq1[i], ddq1[i] = q2[k], dq2[k], ddq2[k]
    t1 = o1.kinematics(q1, dq1, ddq1, want=("torque", "T_tool"))
    t2 = o2.kinematics(q2, dq2, ddq2, want=("torque", "T_tool"))
    assert_close(t2["T_tool"], t1["T_tool"], "T_tool")
    assert_close(t2["torque"], t1["torque"][[3, 0, 5, 1]], "torque rows follow the input order")
    nc = nt.NpChain(d2)
    assert_close(t2["torque"][:, 0], nc.getJointTorque(q2[:, 0], dq2[:, 0], ddq2[:, 0]), "vs transcription")
    assert not fixtures.by_name("c6").set_input_joints(["nope"])
