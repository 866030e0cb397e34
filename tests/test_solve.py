"""Normal-equation solve: host-side, so the numerics are checked on the CPU; the GPU test closes the loop
samples -> fused Gram -> solve -> torque prediction on fresh samples."""
import numpy as np
import pytest

from conftest import assert_close
from oracle.oracle import OracleChain, fill_uniform
from rosdyn_b200 import fixtures
from rosdyn_b200.chain import solveNormalEquations


def _normal_equations(name, n, seed):
    d = fixtures.by_name(name)
    oc = OracleChain(d)
    q, dq, ddq = (fill_uniform(d.n_inputs, n, seed, s) for s in range(3))
    phi, tau = oc.regressor_torque(q, dq, ddq)
    X = phi.reshape(10 * d.n_joints, d.n_inputs, n)
    return oc, X, tau, np.einsum("ari,bri->ab", X, X), np.einsum("ari,ri->a", X, tau), float(np.sum(tau * tau))


@pytest.mark.parametrize("name", ["c6", "c7_perturbed", "random_b"])
def test_minimum_norm_solution(name):
    oc, X, tau, G, b, tt = _normal_equations(name, 400, 21)
    r = solveNormalEquations(G, b, tt)
    P = b.shape[0]
    w, V = np.linalg.eigh(G)
    keep = w > 1e-10 * w.max()
    assert r["rank"] == int(keep.sum()) < P                      # standard parameters are not all identifiable
    ref = V[:, keep] @ ((V[:, keep].T @ b) / w[keep])
    assert_close(r["parameters"], ref, "minimum-norm solution", 1e-8)
    assert_close(np.sort(r["eigenvalues"])[::-1] / w.max(), np.sort(w)[::-1] / w.max(), "eigenvalues", 1e-12)
    # tau came from the rigid-body model, so the fit is exact: residual ~ 0 and Phi pi == tau
    assert abs(r["residual_sq"]) <= 1e-9 * tt
    assert_close(np.einsum("ari,a->ri", X, r["parameters"]), tau, "Phi pi == tau", 1e-8)
    # and it is the projection of the nominal parameters on the identifiable subspace
    pi_nom = oc.nominal_parameters()
    assert_close(r["parameters"], V[:, keep] @ (V[:, keep].T @ pi_nom), "projection of the nominal parameters", 1e-7)


def test_solve_argument_errors():
    with pytest.raises(ValueError):
        solveNormalEquations(np.eye(3), np.zeros(4))
    r = solveNormalEquations(np.zeros((3, 3)), np.zeros(3))
    assert r["rank"] == 0 and np.all(r["parameters"] == 0)


@pytest.mark.gpu
def test_identification_loop_on_gpu():
    import torch
    from rosdyn_b200.chain import Chain, fill_uniform as dev_fill
    d = fixtures.by_name("c6_perturbed")
    ch = Chain(d)
    n = 200_000
    q, dq, ddq = (dev_fill(6, n, 0x1D0, s, device="cuda") for s in range(3))
    G, b, tt = ch.regressorGram(q, dq, ddq)
    sol = solveNormalEquations(G, b, tt)
    assert sol["rank"] < 70 and abs(sol["residual_sq"]) <= 1e-9 * float(tt[0])
    q2, dq2, ddq2 = (dev_fill(6, 1000, 0x1D1, s, device="cuda") for s in range(3))
    phi, tau = ch.getRegressor(q2, dq2, ddq2, with_torque=True)
    pred = torch.einsum("rci,c->ri", phi, torch.tensor(sol["parameters"], device="cuda"))
    assert_close(pred.cpu().numpy(), tau.cpu().numpy(), "torque predicted from the identified parameters", 1e-7)
