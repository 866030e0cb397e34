"""getWrench / getJointTorque with external wrenches (primitives_impl.h:1225-1274) and getJacobianLink (primitives_impl.h:951-979).
CPU: restatement vs the reference's own methods (oracle/_ref).  GPU: the CUDA entry points vs the restatement."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import oracle
from oracle.oracle import OracleChain, fill_uniform
from rosdyn_b200 import fixtures

CH = ["c6", "c7_perturbed", "random_b", "random_c"]
needs_ref = pytest.mark.skipif(not (oracle.have_ref() or oracle.build_ref()), reason="oracle/_ref not built")


def _ext(d, n, seed=0):
    return np.random.default_rng(seed).normal(size=(6 * (d.n_joints + 1), n)) * 10.0


@needs_ref
@pytest.mark.parametrize("name", CH)
def test_restatement_matches_reference(name):
    d = fixtures.by_name(name)
    oc, rc = OracleChain(d), OracleChain(d, fast="ref")
    n = 60
    q, dq, ddq = (fill_uniform(d.n_inputs, n, 31, s) for s in range(3))
    ext = _ext(d, n)
    (ta, wa), (tb, wb) = oc.wrench(q, dq, ddq, ext), rc.wrench(q, dq, ddq, ext)
    assert_close(ta, tb, "torque with external wrenches", 1e-12)
    assert_close(wa, wb, "wrenches", 1e-12)
    t0, _ = oc.wrench(q, dq, ddq)
    assert np.array_equal(t0, oc.kinematics(q, dq, ddq, want=("torque",))["torque"])
    assert np.max(np.abs(ta - t0)) > 1e-3                     # the external wrenches do act
    for link in range(d.n_joints + 1):
        assert_close(oc.jacobian_link(q, link), rc.jacobian_link(q, link), f"jacobian of link {link}", 1e-12)
    assert np.array_equal(oc.jacobian_link(q, d.n_joints), oc.kinematics(q, want=("jacobian",))["jacobian"])
    assert np.all(oc.jacobian_link(q, 0) == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CH)
def test_gpu_wrench_and_link_jacobian(name):
    import torch
    from rosdyn_b200.chain import Chain
    d = fixtures.by_name(name)
    ch, oc = Chain(d), OracleChain(d)
    n = 3000
    q, dq, ddq = (fill_uniform(d.n_inputs, n, 32, s) for s in range(3))
    ext = _ext(d, n, 1)
    tq, tdq, tddq, text = (torch.tensor(x, device="cuda") for x in (q, dq, ddq, ext))
    tau_ref, w_ref = oc.wrench(q, dq, ddq, ext)
    w, tau = ch.getWrench(tq, tdq, tddq, text.reshape(d.n_joints + 1, 6, n), with_torque=True)
    assert_close(tau.cpu().numpy(), tau_ref, f"{name}: torque with external wrenches")
    assert_close(w.reshape(-1, n).cpu().numpy(), w_ref, f"{name}: wrenches")
    w0, tau0 = ch.getWrench(tq, tdq, tddq, with_torque=True)
    assert_close(tau0.cpu().numpy(), ch.getJointTorque(tq, tdq, tddq).cpu().numpy(), "no external wrench == getJointTorque")
    names = ch.getLinksName()
    for link in (0, 1, d.n_joints // 2, d.n_joints):
        J = ch.getJacobianLink(tq, names[link])
        assert_close(J.transpose(0, 1).reshape(6 * d.n_inputs, n).cpu().numpy(), oc.jacobian_link(q, link), f"{name}: jacobian of link {link}")
    assert_close(ch.getJacobianLink(tq, names[-1]).cpu().numpy(), ch.getJacobian(tq).cpu().numpy(), "tool link == getJacobian")
    with pytest.raises(ValueError):
        ch.getJacobianLink(tq, "no_such_link")
    T = ch.getTransformations(tq)
    assert torch.equal(ch.getTransformationLink(tq, names[2]), T[2])
    assert torch.equal(ch.getTwistLink(tq, tdq, names[-1]), ch.getTwistTool(tq, tdq))
