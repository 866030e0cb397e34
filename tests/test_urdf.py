"""CPU-only: the library's URDF loader (csrc/urdf.cpp, rdb_urdf_parse) against an independent xml.etree restatement of the
reference's model build and against the hand-written UR10-like fixture."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from rosdyn_b200 import fixtures
from rosdyn_b200.urdf import chain_from_urdf
from urdf_checker import chain_from_urdf as check_from_urdf

URDF = open(os.path.join(GOLDEN_DIR, "ur10_like.urdf")).read()


def _same(a, b, tol=0.0):
    assert [j.name for j in a.joints] == [j.name for j in b.joints]
    assert [l.name for l in a.links] == [l.name for l in b.links]
    assert a.n_inputs == b.n_inputs and tuple(a.gravity) == tuple(b.gravity)
    for x, y in zip(a.joints, b.joints):
        assert (x.type, x.input_index) == (y.type, y.input_index)
        for f in ("xyz", "rot") + (("axis",) if x.type != 0 else ()):   # the axis of a fixed joint is never used
            np.testing.assert_allclose(getattr(x, f), getattr(y, f), rtol=0, atol=tol)
    for x, y in zip(a.links, b.links):
        assert abs(x.mass - y.mass) <= tol
        for f in ("cog", "inertial_rot", "inertia"):
            np.testing.assert_allclose(getattr(x, f), getattr(y, f), rtol=0, atol=tol)


def test_ur10_like_chain_matches_checker_and_fixture():
    g = (0.0, 0.0, -9.806)
    d = chain_from_urdf(URDF, "base_link", "tool0", g)
    _same(d, check_from_urdf(URDF, "base_link", "tool0", g))            # same arithmetic (urdfdom quaternion path): bit for bit
    ref = fixtures.ur10_like_6r_fixed()
    ref.links[0].inertial_rot = d.links[0].inertial_rot
    _same(d, ref, tol=1e-15)
    assert (d.n_joints, d.n_inputs) == (7, 6)
    assert d.limits["elbow_joint"]["q_max"] == pytest.approx(np.pi) and d.limits["elbow_joint"]["DDq_max"] == pytest.approx(31.5)
    assert d.limits["wrist_3_link-tool0_fixed_joint"]["q_max"] == 0.0


def test_sub_chains_and_world_root():
    d = chain_from_urdf(URDF, "world", "ee_link")                        # through the fixed world joint and the ee side branch
    assert [j.name for j in d.joints][0] == "world_joint" and d.joints[-1].name == "ee_fixed_joint"
    assert (d.n_joints, d.n_inputs) == (8, 6) and tuple(d.gravity) == (0.0, 0.0, 0.0)    # ctor default gravity is zero
    _same(d, check_from_urdf(URDF, "world", "ee_link"))
    d2 = chain_from_urdf(URDF, "upper_arm_link", "wrist_2_link")
    assert [j.name for j in d2.joints] == ["elbow_joint", "wrist_1_joint", "wrist_2_joint"] and d2.n_inputs == 3


def test_errors_follow_the_reference():
    with pytest.raises(LookupError, match="Base link not found"):
        chain_from_urdf(URDF, "nope", "tool0")
    with pytest.raises(LookupError, match="Tool link not found"):
        chain_from_urdf(URDF, "base_link", "nope")
    with pytest.raises(LookupError, match="Tool link not found"):
        chain_from_urdf(URDF, "forearm_link", "shoulder_link")          # tool is not below base
    with pytest.raises(ValueError):
        chain_from_urdf("<robot><link name='a'></robot>", "a", "a")
    with pytest.raises(ValueError):
        chain_from_urdf("<model/>", "a", "a")


def test_types_defaults_and_malformed_limits():
    text = """<robot name="t"><!-- c --><link name="a"/><link name="b"/><link name="c"/><link name="d"/><link name="e"/>
      <joint name="j1" type="continuous"><parent link="a"/><child link="b"/><limit effort="5" velocity="2"/></joint>
      <joint name="j2" type="prismatic"><parent link="b"/><child link="c"/><axis xyz="0 0 2"/><limit lower="1" upper="-1" velocity="-3" effort="7"/></joint>
      <joint name="j3" type="revolute"><parent link="c"/><child link="d"/><origin xyz="1 2 3"/></joint>
      <joint name="j4" type="floating"><parent link="d"/><child link="e"/></joint></robot>"""
    d = chain_from_urdf(text, "a", "e")
    assert [j.type for j in d.joints] == [1, 2, 1, 0] and [j.input_index for j in d.joints] == [0, 1, 2, -1]
    assert tuple(d.joints[0].axis) == (1.0, 0.0, 0.0) and tuple(d.joints[1].axis) == (0.0, 0.0, 2.0)   # normalised later, on create
    L = d.limits
    assert L["j1"]["q_max"] == 1e10 and L["j1"]["DDq_max"] == 20.0 and L["j1"]["tau_max"] == 5.0
    assert L["j2"]["q_max"] == pytest.approx(2 * np.pi) and L["j2"]["q_min"] == pytest.approx(-2 * np.pi)           # upper <= lower
    assert L["j2"]["Dq_max"] == pytest.approx(2 * np.pi) and L["j2"]["DDq_max"] == pytest.approx(20 * np.pi)       # velocity <= 0
    assert L["j3"]["q_max"] == 1e10 and L["j3"]["tau_max"] == 1e10                                                    # no <limit>
    _same(d, check_from_urdf(text, "a", "e"))
    assert d.links[1].mass == 0.0 and tuple(d.links[1].inertia) == (0.0,) * 6                                         # no <inertial>


def test_urdf_chain_through_the_oracle_equals_fixture_chain():
    """End to end on the CPU: dynamics of the URDF-loaded chain == dynamics of the hand-written fixture."""
    from oracle.oracle import OracleChain, fill_uniform
    g = (0.0, 0.0, -9.806)
    a = OracleChain(chain_from_urdf(URDF, "base_link", "tool0", g))
    b = OracleChain(fixtures.ur10_like_6r_fixed())
    q, dq, ddq = (fill_uniform(6, 50, 3, s) for s in range(3))
    pa, ta = a.regressor_torque(q, dq, ddq)
    pb, tb = b.regressor_torque(q, dq, ddq)
    assert np.max(np.abs(pa - pb)) <= 1e-12 and np.max(np.abs(ta - tb)) <= 1e-12
