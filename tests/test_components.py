"""Additive joint components: friction / spring regressor columns.
CPU: the restatement against the reference's own component classes (oracle/_ref) incl. the constructor quirks.
GPU: the CUDA entry points against the restatement, and the extended normal equations against a numpy contraction."""
import numpy as np
import pytest

from conftest import assert_close
from oracle import oracle
from rosdyn_b200 import fixtures

# (type, input index, min_velocity, max_velocity): poly1, poly2, spring, poly2 with the reference's negative-max quirk
# (friction_polynomial2.h:91-96), poly1 with max_velocity <= 0 -> 1e6 (friction_polynomial1.h:82-87), tiny thresholds -> 1e-6
COMPS = [(1, 0, 0.05, 0.8), (2, 1, 0.1, 0.7), (3, 2, 0.0, 0.0), (2, 3, 1e-9, -1.0), (1, 4, 0.01, -5.0), (2, 5, 1e-12, 2.0)]
PC = 2 + 3 + 2 + 3 + 2 + 3


def _inputs(n, seed=3):
    q, dq, ddq = (oracle.fill_uniform(6, n, seed, s) for s in range(3))
    dq[:, 0] = 0.0            # omega == 0 branch
    dq[1, 1], dq[1, 2] = 0.1, -0.1   # exactly on the threshold
    dq[:, 3] *= 1e-3          # inside the linear zone
    dq[:, 4] *= 5.0           # saturated
    return q, dq, ddq


needs_ref = pytest.mark.skipif(not (oracle.have_ref() or oracle.build_ref()), reason="oracle/_ref not built")


@needs_ref
def test_components_restatement_matches_reference_classes():
    q, dq, _ = _inputs(600)
    a = oracle.components_regressor(COMPS, 6, q, dq)
    b = oracle.components_regressor(COMPS, 6, q, dq, fast="ref")
    assert a.shape == (PC * 6, 600) and not np.isnan(a).any() and not np.isnan(b).any()
    assert np.array_equal(a, b)
    # a component's columns are zero except in the row of its joint
    A = a.reshape(PC, 6, 600)
    col = 0
    for t, j, _, _ in COMPS:
        nc = 3 if t == 2 else 2
        mask = np.ones(6, bool)
        mask[j] = False
        assert np.all(A[col:col + nc, mask] == 0.0)
        col += nc


def test_component_columns_known_values():
    q = np.array([[0.25]]); dq = np.array([[0.03]])
    a = oracle.components_regressor([(1, 0, 0.05, 0.8), (2, 0, 0.05, 0.8), (3, 0, 0, 0)], 1, q, dq)[:, 0]
    np.testing.assert_allclose(a, [0.6, 0.03, 0.6, 0.03, 0.03 ** 2 * 0.6, 0.25, 1.0], rtol=1e-15)


@pytest.mark.gpu
def test_components_gpu_against_oracle():
    import torch
    from rosdyn_b200.chain import Chain
    ch = Chain(fixtures.by_name("c6"))
    assert ch.getComponentColumns() == 0
    names = ch.getActiveJointsName()
    tn = {1: "friction1", 2: "friction2", 3: "spring"}
    comps = [{"type": tn[t], "joint": names[j] if k % 2 else j, "min_velocity": lo, "max_velocity": hi} for k, (t, j, lo, hi) in enumerate(COMPS)]
    assert ch.setComponents(comps) == PC
    n = 5000
    q, dq, ddq = _inputs(n)
    dq_, q_, ddq_ = (torch.tensor(x, device="cuda") for x in (dq, q, ddq))
    ref = oracle.components_regressor(COMPS, 6, q, dq)
    phi_c = ch.getComponentsRegressor(q_, dq_)
    assert np.array_equal(phi_c.transpose(0, 1).reshape(PC * 6, n).cpu().numpy(), ref)      # element-wise fp64: bit exact
    prm = np.linspace(0.5, 2.0, PC)
    tau = ch.getComponentsTorque(q_, dq_, prm).cpu().numpy()
    assert_close(tau, np.einsum("cri,c->ri", ref.reshape(PC, 6, n), prm), "components torque")
    rigid = ch.getJointTorque(q_, dq_, ddq_)
    both = ch.getComponentsTorque(q_, dq_, prm, out=rigid.clone()).cpu().numpy()
    assert_close(both, rigid.cpu().numpy() + tau, "accumulated onto the rigid-body torque")
    with pytest.raises(LookupError):
        ch.setComponents([{"type": "spring", "joint": "no_such_joint"}])
    # extended normal equations [Phi | Phi_c]
    from oracle.oracle import OracleChain
    oc = OracleChain(fixtures.by_name("c6"))
    phi, tau_r = oc.regressor_torque(q, dq, ddq)
    X = np.concatenate([phi.reshape(70, 6, n), ref.reshape(PC, 6, n)], axis=0)              # [col][row][sample]
    G_ref = np.einsum("ari,bri->ab", X, X)
    b_ref = np.einsum("ari,ri->a", X, tau_r)
    G, b, tt = ch.regressorGramExt(q_, dq_, ddq_)
    assert G.shape == (70 + PC, 70 + PC)
    assert_close(G.cpu().numpy() / np.max(np.abs(G_ref)), G_ref / np.max(np.abs(G_ref)), "extended gram", 1e-12)
    assert_close(b.cpu().numpy() / np.max(np.abs(b_ref)), b_ref / np.max(np.abs(b_ref)), "extended rhs", 1e-12)
    assert abs(float(tt[0]) - float(np.sum(tau_r * tau_r))) <= 1e-12 * float(np.sum(tau_r * tau_r))
    tau_meas = torch.tensor(tau_r + both - rigid.cpu().numpy(), device="cuda")                 # rigid + components
    G2, b2, _ = ch.regressorGramExt(q_, dq_, ddq_, tau_meas=tau_meas)
    assert_close(b2.cpu().numpy() / np.max(np.abs(b_ref)), np.einsum("ari,ri->a", X, tau_meas.cpu().numpy()) / np.max(np.abs(b_ref)), "rhs, measured torque", 1e-12)
    # accumulation over two chunks == one pass; the matrix is symmetric; the general pipeline (RDB_GRAM_IMPL=v0 at load time) is the same
    # computation, so its result is what the fused cross mode must reproduce on other chains too (folded chain with a massive tool below)
    h = 2048
    o = ch.regressorGramExt(q_[:, :h].contiguous(), dq_[:, :h].contiguous(), ddq_[:, :h].contiguous())
    o = ch.regressorGramExt(q_[:, h:].contiguous(), dq_[:, h:].contiguous(), ddq_[:, h:].contiguous(), out=o)
    assert float((o[0] - G).abs().max()) <= 1e-12 * float(G.abs().max()) and float((o[1] - b).abs().max()) <= 1e-12 * float(b.abs().max())
    assert bool((G == G.T).all())
    # the rigid-body entry is unchanged by the components
    G0, _, _ = ch.regressorGram(q_, dq_, ddq_)
    assert_close(G0.cpu().numpy() / np.max(np.abs(G_ref)), G_ref[:70, :70] / np.max(np.abs(G_ref)), "rigid gram", 1e-12)
    ch.setComponents([])
    assert ch.getComponentColumns() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c6_perturbed", "c7_perturbed", "random_b"])
def test_extended_gram_on_folded_and_mixed_chains(name):
    """Cross mode of the fused kernel on chains with a massive rigidly attached link, 7 moving joints, prismatic joints: the extended normal
    equations against the oracle's regressor and component columns contracted in numpy."""
    import torch
    from oracle.oracle import OracleChain
    from rosdyn_b200.chain import Chain
    d = fixtures.by_name(name)
    ch, oc = Chain(d), OracleChain(d)
    n_in = d.n_inputs
    comps = [(1 + (k % 3), k % n_in, 0.05, 0.7) for k in range(n_in + 2)]   # two joints carry two components
    tn = {1: "friction1", 2: "friction2", 3: "spring"}
    pc = ch.setComponents([{"type": tn[t], "joint": j, "min_velocity": lo, "max_velocity": hi} for t, j, lo, hi in comps])
    n = 3000
    rng = np.random.RandomState(4)
    q, dq, ddq = (rng.uniform(-1, 1, (n_in, n)) for _ in range(3))
    phi, tau_r = oc.regressor_torque(q, dq, ddq)
    ref_c = oracle.components_regressor(comps, n_in, q, dq)
    P = 10 * d.n_joints
    X = np.concatenate([phi.reshape(P, n_in, n), ref_c.reshape(pc, n_in, n)], axis=0)
    G_ref = np.einsum("ari,bri->ab", X, X)
    b_ref = np.einsum("ari,ri->a", X, tau_r)
    G, b, tt = ch.regressorGramExt(*(torch.tensor(x, device="cuda") for x in (q, dq, ddq)))
    assert_close(G.cpu().numpy() / np.max(np.abs(G_ref)), G_ref / np.max(np.abs(G_ref)), "extended gram", 1e-11)
    assert_close(b.cpu().numpy() / np.max(np.abs(b_ref)), b_ref / np.max(np.abs(b_ref)), "extended rhs", 1e-11)
    assert bool((G == G.T).all())


@pytest.mark.gpu
@pytest.mark.parametrize("name,ctype", [("c6", 1), ("c7", 1), ("c6", 3), ("random_b", 1)])
def test_extended_gram_two_column_components(name, ctype):
    """One two-column component (friction_polynomial1 or ideal spring) on every joint: the narrow side buffer of the one-pass kernel (4 slots
    for 6 joints), ragged batch, against the oracle's regressor and component columns contracted in numpy."""
    import torch
    from oracle.oracle import OracleChain
    from rosdyn_b200.chain import Chain
    d = fixtures.by_name(name)
    ch, oc = Chain(d), OracleChain(d)
    n_in = d.n_inputs
    comps = [(ctype, k, 0.05, 0.7) for k in range(n_in)]
    tn = {1: "friction1", 3: "spring"}
    pc = ch.setComponents([{"type": tn[t], "joint": j, "min_velocity": lo, "max_velocity": hi} for t, j, lo, hi in comps])
    assert pc == 2 * n_in
    n = 4099  # ends inside a group and inside a k-step
    rng = np.random.RandomState(5)
    q, dq, ddq = (rng.uniform(-1, 1, (n_in, n)) for _ in range(3))
    dq[:, :3] = 0.0
    phi, tau_r = oc.regressor_torque(q, dq, ddq)
    ref_c = oracle.components_regressor(comps, n_in, q, dq)
    P = 10 * d.n_joints
    X = np.concatenate([phi.reshape(P, n_in, n), ref_c.reshape(pc, n_in, n)], axis=0)
    G_ref = np.einsum("ari,bri->ab", X, X)
    b_ref = np.einsum("ari,ri->a", X, tau_r)
    G, b, tt = ch.regressorGramExt(*(torch.tensor(x, device="cuda") for x in (q, dq, ddq)))
    assert_close(G.cpu().numpy() / np.max(np.abs(G_ref)), G_ref / np.max(np.abs(G_ref)), "extended gram", 1e-11)
    assert_close(b.cpu().numpy() / np.max(np.abs(b_ref)), b_ref / np.max(np.abs(b_ref)), "extended rhs", 1e-11)
    assert bool((G == G.T).all())
    G2, b2, _ = ch.regressorGramExt(*(torch.tensor(x, device="cuda") for x in (q, dq, ddq)))
    assert torch.equal(G, G2) and torch.equal(b, b2)
