"""GPU parity: the CUDA engine, called through the C-ABI, against the CPU oracle on the same seeded inputs,
against the committed golden vectors, and through size-independent invariants at large N.
Tolerance: fp64, maxabs(x - ref) <= 1e-10 * max(maxabs(ref), 1) per output array (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import CHAINS, RTOL, assert_close, rel_err
from rosdyn_b200 import fixtures

pytestmark = pytest.mark.gpu

KIN = ("T_tool", "T_links", "jacobian", "twist", "dtwist", "dtwist_lin", "dtwist_nonlin", "ddtwist", "ddtwist_lin", "ddtwist_nonlin",
       "torque")


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def chains(torch):
    from oracle.oracle import OracleChain
    from rosdyn_b200.chain import Chain
    cache = {}

    def get(name):
        if name not in cache:
            d = fixtures.by_name(name)
            cache[name] = (d, Chain(d), OracleChain(d))
        return cache[name]
    return get


def _inputs(torch, n_in, n, seed):
    from rosdyn_b200.chain import fill_uniform
    return [fill_uniform(n_in, n, seed, s, device="cuda") for s in range(4)]


def _np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("source", ["numpy_transcription", "reference"])
@pytest.mark.parametrize("name", CHAINS)
def test_golden_vectors(name, source, chains, golden, torch):
    """source = "reference": outputs of the reference's own rosdyn::Chain (tests/golden/ref_*.npz, made by
    tests/golden/make_golden_ref.py from oracle/_ref); "numpy_transcription": the independent transcription (make_golden.py)."""
    d, ch, _ = chains(name)
    g = golden(name if source == "numpy_transcription" else "ref_" + name)
    dev = [torch.tensor(g[k], device="cuda") for k in ("q", "dq", "ddq", "dddq")]
    K = ch.kinematics(*dev, want=KIN)
    for k in KIN:
        ref = g["T_links"][-12:] if k == "T_tool" else g[k]
        assert_close(_np(K[k]), ref, f"{name}:{k}")
    phi, tau = ch.getRegressor(dev[0], dev[1], dev[2], with_torque=True)
    P = 10 * d.n_joints
    assert_close(_np(phi.transpose(0, 1).reshape(P * d.n_inputs, -1)), g["regressor"], f"{name}:regressor")
    assert_close(_np(tau), g["torque"], f"{name}:torque(regressor pass)")
    assert_close(_np(ch.getJointTorque(dev[0], dev[1], dev[2])), g["torque"], f"{name}:torque")
    assert_close(_np(ch.getJointTorqueNonLinearPart(dev[0], dev[1])), g["torque_nonlin"], f"{name}:torque_nonlin")
    M = ch.getJointInertia(dev[0])
    assert_close(_np(M.transpose(0, 1).reshape(d.n_inputs ** 2, -1)), g["inertia"], f"{name}:inertia")
    assert_close(ch.getNominalParameters(), g["nominal"], f"{name}:nominal")


@pytest.mark.parametrize("name", CHAINS)
def test_against_oracle_seeded(name, chains, torch):
    d, ch, oc = chains(name)
    n = 20000 if d.n_joints <= 8 else 4000
    q, dq, ddq, dddq = _inputs(torch, d.n_inputs, n, 0x5EED0000 + 17)
    hq, hdq, hddq, hdddq = (_np(x) for x in (q, dq, ddq, dddq))
    from oracle.oracle import fill_uniform
    assert np.array_equal(hq, fill_uniform(d.n_inputs, n, 0x5EED0000 + 17, 0)), "device and host generators must agree bit for bit"
    K = ch.kinematics(q, dq, ddq, dddq, want=KIN)
    R = oc.kinematics(hq, hdq, hddq, hdddq, nthreads=0)
    for k in KIN:
        assert_close(_np(K[k]), R[k], f"{name}:{k}")
    rphi, rtau = oc.regressor_torque(hq, hdq, hddq, nthreads=0)
    phi, tau = ch.getRegressor(q, dq, ddq, with_torque=True)
    P = 10 * d.n_joints
    planes = _np(phi.transpose(0, 1).reshape(P * d.n_inputs, n))
    assert_close(planes, rphi, f"{name}:regressor")
    Phi = planes.reshape(P, d.n_inputs, n)
    for j, jd in enumerate(d.joints):  # block upper-triangular: exact zeros in the column blocks of the links before joint j
        if jd.input_index >= 0:
            assert np.all(Phi[:10 * j, jd.input_index, :] == 0.0), "structural zeros must be exact"
    assert_close(_np(tau), rtau, f"{name}:torque via regressor pass")
    assert_close(_np(ch.getRegressor(q, dq, ddq)).reshape(-1), _np(phi).reshape(-1), "regressor without torque", 0.0)
    assert_close(_np(ch.getJointTorque(q, dq, ddq)), rtau, f"{name}:torque")
    assert_close(_np(ch.getJointInertia(q).transpose(0, 1).reshape(d.n_inputs ** 2, n)), oc.inertia(hq, nthreads=0), f"{name}:inertia")


@pytest.mark.parametrize("name", ["c6", "c7_perturbed", "random_b", "random_c"])
def test_gram_against_long_double_oracle(name, chains, torch):
    d, ch, oc = chains(name)
    n = 3001  # ragged: not a multiple of 4 / 32 / 128
    q, dq, ddq, _ = _inputs(torch, d.n_inputs, n, 0x5EED0000 + 4)
    Gr, br, ttr = oc.gram(_np(q), _np(dq), _np(ddq))
    G, b, tt = ch.regressorGram(q, dq, ddq)
    G, b, tt = _np(G), _np(b), _np(tt)
    scale = np.max(np.abs(Gr))
    assert np.max(np.abs(G - Gr)) <= 1e-10 * scale, np.max(np.abs(G - Gr)) / scale
    assert np.array_equal(G, G.T)
    assert np.max(np.abs(b - br)) <= 1e-10 * np.max(np.abs(br))
    assert abs(tt[0] - ttr) <= 1e-10 * ttr
    # tau from RNEA == Phi pi_nom  =>  G pi_nom == b
    pi = ch.getNominalParameters()
    assert np.max(np.abs(G @ pi - b)) <= 1e-9 * np.max(np.abs(b))
    # external tau_meas + accumulation over two chunks == one pass
    tau = ch.getJointTorque(q, dq, ddq) * 1.5
    G1, b1, t1 = ch.regressorGram(q, dq, ddq, tau_meas=tau)
    h = 1024
    out = ch.regressorGram(q[:, :h].contiguous(), dq[:, :h].contiguous(), ddq[:, :h].contiguous(), tau_meas=tau[:, :h].contiguous())
    out = ch.regressorGram(q[:, h:], dq[:, h:], ddq[:, h:], tau_meas=tau[:, h:], out=out)
    assert np.max(np.abs(_np(out[0]) - _np(G1))) <= 1e-12 * scale
    assert np.max(np.abs(_np(out[1]) - _np(b1))) <= 1e-12 * np.max(np.abs(_np(b1)))
    assert np.max(np.abs(_np(b1) - 1.5 * br)) <= 1e-10 * np.max(np.abs(br)) * 1.5


def _fold_case(case):
    """Chains whose never-moving joints exercise every branch of the folded-chain Gram path (fold.cpp: fold_chain)."""
    from rosdyn_b200.descriptor import FIXED
    if case == "leading+double_interior":  # fixed joint at the base, two consecutive fixed joints inside, massive links everywhere
        d = fixtures.random_chain(909, 9, p_prismatic=0.3, p_fixed=0.0)
        for j in (0, 3, 4):
            d.joints[j].type = FIXED
        d.set_default_inputs()
    elif case == "trailing_massive":  # C6 with a tool that has mass, cog and a full rotated inertia
        d = fixtures.by_name("c6_perturbed")
    elif case == "unlisted_and_permuted":  # moving-type joints that are not inputs behave as fixed at q = 0 (PI.h:865)
        d = fixtures.by_name("c7_perturbed")
        names = [j.name for j in d.joints]
        d.set_input_joints([names[5], names[0], names[3], names[2]])
    else:  # 11 joints, 7 of them moving: only the folded chain fits the fused kernel
        d = fixtures.random_chain(777, 11, p_prismatic=0.2, p_fixed=0.0)
        for j in (1, 4, 7, 10):
            d.joints[j].type = FIXED
        d.set_default_inputs()
    return d


@pytest.mark.parametrize("case", ["leading+double_interior", "trailing_massive", "unlisted_and_permuted", "long_chain"])
def test_gram_folded_chain(case, torch):
    from oracle.oracle import OracleChain
    from rosdyn_b200.chain import Chain
    d = _fold_case(case)
    ch, oc = Chain(d), OracleChain(d)
    n = 2500
    q, dq, ddq, _ = _inputs(torch, d.n_inputs, n, 0x5EED0000 + 9)
    Gr, br, ttr = oc.gram(_np(q), _np(dq), _np(ddq))
    G, b, tt = (_np(x) for x in ch.regressorGram(q, dq, ddq))
    scale = np.max(np.abs(Gr))
    assert np.max(np.abs(G - Gr)) <= 1e-10 * scale, np.max(np.abs(G - Gr)) / scale
    assert np.array_equal(G, G.T)
    assert np.max(np.abs(b - br)) <= 1e-10 * np.max(np.abs(br))
    assert abs(tt[0] - ttr) <= 1e-10 * ttr
    # the materialised regressor (no folding there) gives the same normal equations
    phi, tau = ch.getRegressor(q, dq, ddq, with_torque=True)
    F = _np(phi).transpose(2, 0, 1).reshape(n * d.n_inputs, -1)  # (sample, row) x column
    assert np.max(np.abs(F.T @ F - G)) <= 1e-10 * scale
    assert np.max(np.abs(F.T @ _np(tau).T.reshape(-1) - b)) <= 1e-10 * np.max(np.abs(br))


def test_edge_cases(chains, torch):
    d, ch, oc = chains("c6")
    # empty batch
    e = torch.empty((6, 0), dtype=torch.float64, device="cuda")
    assert ch.getJointTorque(e, e, e).shape == (6, 0)
    assert ch.getRegressor(e, e, e).shape == (6, 70, 0)
    G, b, tt = ch.regressorGram(e, e, e)
    assert float(G.abs().sum()) == 0.0 and float(b.abs().sum()) == 0.0
    # one sample, 1-D API like the reference
    q = torch.zeros(6, dtype=torch.float64, device="cuda")
    T = _np(ch.getTransformation(q))
    np.testing.assert_allclose(T[:3, 3], [1.1843, 0.256141, 0.0116], atol=1e-12)  # UR10 zero pose known answer
    assert T.shape == (4, 4) and ch.getJacobian(q).shape == (6, 6) and ch.getRegressor(q, q, q).shape == (6, 70)
    # ragged N and a padded leading dimension (ld > n)
    n = 1000 + 37
    buf = [torch.full((6, n + 11), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3)]
    src = _inputs(torch, 6, n, 99)
    for bb, s in zip(buf, src):
        bb[:, :n] = s
    tau = ch.getJointTorque(buf[0][:, :n], buf[1][:, :n], buf[2][:, :n])
    assert_close(_np(tau), _np(ch.getJointTorque(*src[:3])), "ld > n", 0.0)
    # dimension mismatch -> the reference's std::invalid_argument
    with pytest.raises(ValueError, match="dimensions mismatch"):
        ch.getRegressor(src[0], src[1][:5], src[2])
    with pytest.raises(ValueError, match="dimensions mismatch"):
        ch.getRegressor(src[0], src[1], None)
    # large |q| (range reduction of sincos), zero gravity, q = +-pi
    big = src[0] * 1.0e4
    assert_close(_np(ch.getJointTorque(big, src[1], src[2])), oc.kinematics(_np(big), _np(src[1]), _np(src[2]), want=("torque",))["torque"],
                 "large |q|", 1e-9)


def test_zero_gravity_and_input_selection(torch):
    from oracle.oracle import OracleChain
    from rosdyn_b200.chain import Chain
    d = fixtures.by_name("c6_perturbed")
    d.gravity = (0.0, 0.0, 0.0)  # the Chain ctor default (primitives.h:346)
    names = [j.name for j in d.joints if j.type != 0]
    sel = [names[i] for i in (4, 0, 2)]
    ch = Chain(d)
    assert ch.setInputJointsName(sel) and ch.getActiveJointsNumber() == 3
    d2 = fixtures.by_name("c6_perturbed")
    d2.gravity = (0.0, 0.0, 0.0)
    d2.set_input_joints(sel)
    oc = OracleChain(d2)
    q, dq, ddq, dddq = _inputs(torch, 3, 500, 5)
    K = ch.kinematics(q, dq, ddq, dddq, want=KIN)
    R = oc.kinematics(*(_np(x) for x in (q, dq, ddq, dddq)))
    for k in KIN:
        assert_close(_np(K[k]), R[k], k)
    rphi, rtau = oc.regressor_torque(_np(q), _np(dq), _np(ddq))
    phi, tau = ch.getRegressor(q, dq, ddq, with_torque=True)
    assert_close(_np(phi.transpose(0, 1).reshape(70 * 3, -1)), rphi, "regressor (3 selected inputs)")
    assert_close(_np(tau), rtau, "torque")
    assert_close(_np(ch.getJointInertia(q).transpose(0, 1).reshape(9, -1)), oc.inertia(_np(q)), "inertia")
    assert not ch.setInputJointsName(["no_such_joint"])


def test_host_buffer_entry_points(chains, torch):
    d, ch, oc = chains("c7")
    from oracle.oracle import fill_uniform
    n = 300001  # > one pipeline chunk for the regressor path, ragged
    q, dq, ddq, dddq = (fill_uniform(7, n, 21, s) for s in range(4))
    tau = ch.getJointTorque(q, dq, ddq)
    assert isinstance(tau, np.ndarray)
    sub = slice(0, n, 997)
    ref = oc.kinematics(q[:, sub], dq[:, sub], ddq[:, sub], dddq[:, sub])
    assert_close(tau[:, sub], ref["torque"], "host torque")
    K = ch.kinematics(q, dq, ddq, dddq, want=("T_tool", "ddtwist", "jacobian"))
    for k in ("T_tool", "ddtwist", "jacobian"):
        assert_close(K[k][:, sub], ref[k], f"host {k}")
    phi, t2 = ch.getRegressor(q, dq, ddq, with_torque=True)
    rphi, _ = oc.regressor_torque(q[:, sub], dq[:, sub], ddq[:, sub])
    assert_close(np.swapaxes(phi, 0, 1).reshape(490, n)[:, sub], rphi, "host regressor")
    assert_close(t2, tau, "host torque from the regressor pass", 1e-12)
    G, b, tt = ch.regressorGram(q, dq, ddq)
    Gd, bd, ttd = ch.regressorGram(*(torch.tensor(x, device="cuda") for x in (q, dq, ddq)))
    assert np.max(np.abs(G - _np(Gd))) <= 1e-12 * np.max(np.abs(G))
    assert np.max(np.abs(b - _np(bd))) <= 1e-12 * np.max(np.abs(b))


def test_large_batch_invariants(chains, torch):
    """BASELINE-scale property checks that need no CPU oracle: Phi pi == tau, M ddq + h == tau, J dq == v_tool."""
    d, ch, _ = chains("c6")
    n = 4_000_000
    q, dq, ddq, _ = _inputs(torch, 6, n, 0x5EED0001)
    pi = torch.tensor(ch.getNominalParameters(), device="cuda")
    tau = ch.getJointTorque(q, dq, ddq)
    step = 1_000_000
    for s in range(0, n, step):
        sl = slice(s, s + step)
        phi = ch.getRegressor(q[:, sl], dq[:, sl], ddq[:, sl])          # [6,70,step]
        err = (torch.einsum("rcs,c->rs", phi, pi) - tau[:, sl]).abs().max() / tau.abs().max()
        assert float(err) <= RTOL
        del phi
    M = ch.getJointInertia(q)
    h = ch.getJointTorqueNonLinearPart(q, dq)
    err = (torch.einsum("rcs,cs->rs", M, ddq) + h - tau).abs().max() / tau.abs().max()
    assert float(err) <= RTOL
    assert float((M - M.transpose(0, 1)).abs().max()) == 0.0
    del M
    J = ch.getJacobian(q)
    v = ch.getTwist(q, dq)[-1]
    assert float((torch.einsum("rcs,cs->rs", J, dq) - v).abs().max()) <= RTOL * max(1.0, float(v.abs().max()))


def test_gram_large_batch_invariants(chains, torch):
    """Fused normal equations at BASELINE scale without a CPU oracle: G pi_nom == b (tau = Phi pi_nom), tau_sq == pi^T G pi, additivity over
    shards (what the multi-GPU all-reduce relies on), and agreement with Phi^T Phi of the MATERIALISED regressor contracted by cuBLAS."""
    for name in ("c6", "c7"):
        d, ch, _ = chains(name)
        n = 6_000_000
        q, dq, ddq, _ = _inputs(torch, d.n_inputs, n, 0x5EED0002)
        G, b, tt = ch.regressorGram(q, dq, ddq)
        pi = torch.tensor(ch.getNominalParameters(), device="cuda")
        assert float((G @ pi - b).abs().max()) <= 1e-10 * float(b.abs().max())
        assert abs(float(pi @ G @ pi) - float(tt[0])) <= 1e-10 * float(tt[0])
        assert bool((G == G.T).all())
        h = 2_500_032
        o = ch.regressorGram(q[:, :h].contiguous(), dq[:, :h].contiguous(), ddq[:, :h].contiguous())
        o = ch.regressorGram(q[:, h:].contiguous(), dq[:, h:].contiguous(), ddq[:, h:].contiguous(), out=o)
        scale = float(G.abs().max())
        assert float((o[0] - G).abs().max()) <= 1e-11 * scale and float((o[1] - b).abs().max()) <= 1e-11 * float(b.abs().max())
        m = 1_000_000
        phi, tau = ch.getRegressor(q[:, :m].contiguous(), dq[:, :m].contiguous(), ddq[:, :m].contiguous(), with_torque=True)
        F = phi.permute(2, 0, 1).reshape(m * d.n_inputs, -1)
        Gm, bm, _ = ch.regressorGram(q[:, :m].contiguous(), dq[:, :m].contiguous(), ddq[:, :m].contiguous())
        Fg = F.T @ F
        assert float((Fg - Gm).abs().max()) <= 1e-10 * float(Fg.abs().max())
        assert float((F.T @ tau.T.reshape(-1) - bm).abs().max()) <= 1e-10 * float(bm.abs().max())
        del phi, F


def test_sharded_gram_single_process(chains, torch):
    """rdb_regressor_gram_sharded_host: several handles (one per GPU when the box has more than one, else all on cuda:0) each take a
    contiguous shard of a host batch; the rank-ordered host sum equals the single-handle result and the oracle."""
    from rosdyn_b200.chain import Chain
    from rosdyn_b200.sharding import sharded_gram_host
    d, ch, oc = chains("c6")
    ndev = torch.cuda.device_count()
    handles = [Chain(d, device=r % ndev) for r in range(3)]
    assert [h._lib.rdb_chain_device(h._h) for h in handles] == [r % ndev for r in range(3)]
    n = 300_001
    q, dq, ddq, _ = (_np(x) for x in _inputs(torch, 6, n, 0x5EED0003))
    G, b, tt = sharded_gram_host(handles, q, dq, ddq)
    G1, b1, t1 = (_np(x) for x in ch.regressorGram(*(torch.tensor(x, device="cuda") for x in (q, dq, ddq))))
    scale = np.max(np.abs(G1))
    assert np.max(np.abs(G - G1)) <= 1e-12 * scale and np.max(np.abs(b - b1)) <= 1e-12 * np.max(np.abs(b1))
    assert abs(tt - t1[0]) <= 1e-12 * t1[0]
    G2, b2, _ = sharded_gram_host(handles, q, dq, ddq)
    assert np.array_equal(G, G2) and np.array_equal(b, b2)          # rank-ordered host sum: reproducible
    m = 20_000
    Gr, br, _ = oc.gram(q[:, :m], dq[:, :m], ddq[:, :m])
    Gs, bs, _ = sharded_gram_host(handles[:2], q[:, :m].copy(), dq[:, :m].copy(), ddq[:, :m].copy())
    assert np.max(np.abs(Gs - Gr)) <= 1e-10 * np.max(np.abs(Gr)) and np.max(np.abs(bs - br)) <= 1e-10 * np.max(np.abs(br))


def test_chain_from_urdf(chains, torch):
    """createChain(urdf, base, tool, g) through the library's URDF loader == the hand-written fixture chain."""
    import os
    from conftest import GOLDEN_DIR
    from rosdyn_b200.chain import Chain, createChain
    urdf = open(os.path.join(GOLDEN_DIR, "ur10_like.urdf")).read()
    ch = createChain(urdf, "base_link", "tool0", (0.0, 0.0, -9.806))
    assert ch is not None and ch.getActiveJointsNumber() == 6 and ch.getLinksName()[-1] == "tool0"
    assert createChain(urdf, "base_link", "no_such_link", (0, 0, 0)) is None      # reference: createChain returns null
    with pytest.raises(LookupError, match="Tool link not found"):
        Chain.from_urdf(urdf, "base_link", "no_such_link")
    _, ref, _ = chains("c6")
    q, dq, ddq, _ = _inputs(torch, 6, 1000, 31)
    assert_close(_np(ch.getJointTorque(q, dq, ddq)), _np(ref.getJointTorque(q, dq, ddq)), "torque", 1e-13)
    assert_close(_np(ch.getRegressor(q, dq, ddq)), _np(ref.getRegressor(q, dq, ddq)), "regressor", 1e-13)


def test_cpp_facade(torch, tmp_path):
    """The C++ rosdyn::Chain facade over the C-ABI (include/rosdyn_b200/chain.hpp): build the example and run its self-checks."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "build", "facade_check")
    if not os.path.exists(exe):
        pytest.skip("build/facade_check not built")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "rosdyn_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):   # per-call latency of the N = 1 getters (kept as evidence: profiles/r02_latency.txt)
        open(os.path.join(out, "r02_latency.txt"), "w").write(r.stdout)


def test_cpp_eigen_overloads(torch, tmp_path):
    """The Eigen-typed overloads of chain.hpp (the reference's signatures), compiled against the Eigen subset of oracle/shim, run on the GPU:
    Phi pi = tau, M ddq + h = tau, J dq = v_tool on the Eigen objects, and the batched Eigen-record sibling."""
    import subprocess
    from test_cpp_headers import build_eigen_facade
    exe = build_eigen_facade(str(tmp_path / "eigen_facade"))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "eigen facade ok" in r.stdout, r.stdout + r.stderr
