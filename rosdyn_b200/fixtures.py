"""Benchmark / test chains.

The reference's tests load the UR10 from the external ROS package `ur_description`
(rosdyn_core/test/test.launch:15), which is not vendored; these descriptors are synthetic fixtures with the
public UR10 geometry, shared by the CUDA engine, the oracle and the golden-vector generator.
"""
from __future__ import annotations

import copy
import math
from typing import List

import numpy as np

from .descriptor import FIXED, PRISMATIC, REVOLUTE, ChainDesc, JointDesc, LinkDesc, rpy_to_rot

GRAVITY = (0.0, 0.0, -9.806)  # rosdyn_speed_test.cpp:61-62, test.cpp:64-65


def _cyl(m: float, r: float, l: float):
    ixx = m * (3 * r * r + l * l) / 12.0
    return (ixx, 0.0, 0.0, ixx, 0.0, m * r * r / 2.0)


def ur10_like_6r_fixed() -> ChainDesc:
    """C6: UR10-like base_link -> tool0, 6 revolute + 1 fixed joint: nJ=7, nL=8, n_act=6, regressor 6 x 70."""
    hp = math.pi / 2
    J = [
        JointDesc("shoulder_pan_joint", REVOLUTE, (0, 0, 0.1273), rpy_to_rot(0, 0, 0), (0, 0, 1)),
        JointDesc("shoulder_lift_joint", REVOLUTE, (0, 0.220941, 0), rpy_to_rot(0, hp, 0), (0, 1, 0)),
        JointDesc("elbow_joint", REVOLUTE, (0, -0.1719, 0.612), rpy_to_rot(0, 0, 0), (0, 1, 0)),
        JointDesc("wrist_1_joint", REVOLUTE, (0, 0, 0.5723), rpy_to_rot(0, hp, 0), (0, 1, 0)),
        JointDesc("wrist_2_joint", REVOLUTE, (0, 0.1149, 0), rpy_to_rot(0, 0, 0), (0, 0, 1)),
        JointDesc("wrist_3_joint", REVOLUTE, (0, 0, 0.1157), rpy_to_rot(0, 0, 0), (0, 1, 0)),
        JointDesc("wrist_3_link-tool0_fixed_joint", FIXED, (0, 0.0922, 0), rpy_to_rot(-hp, 0, 0), (0, 0, 0)),
    ]
    r = 0.075
    L = [
        LinkDesc("base_link", 4.0, (0, 0, 0), inertia=(0.0061063308908, 0, 0, 0.0061063308908, 0, 0.01125)),
        LinkDesc("shoulder_link", 7.778, (0, 0, 0), inertia=_cyl(7.778, r, 0.178)),
        LinkDesc("upper_arm_link", 12.93, (0, 0, 0.306), inertia=_cyl(12.93, r, 0.612)),
        LinkDesc("forearm_link", 3.87, (0, 0, 0.28615), inertia=_cyl(3.87, r, 0.5723)),
        LinkDesc("wrist_1_link", 1.96, (0, 0, 0), inertia=_cyl(1.96, r, 0.12)),
        LinkDesc("wrist_2_link", 1.96, (0, 0, 0), inertia=_cyl(1.96, r, 0.12)),
        LinkDesc("wrist_3_link", 0.202, (0, 0, 0), inertia=_cyl(0.202, r, 0.12)),
        LinkDesc("tool0"),  # no <inertial>: mass 0, zero inertia (PI.h:320-328)
    ]
    return ChainDesc(J, L, GRAVITY, name="ur10_like_6r_fixed")


def chain_7r() -> ChainDesc:
    """C7: C6 with the fixed tool joint replaced by a 7th revolute joint: nJ=7, n_act=7, regressor 7 x 70."""
    c = ur10_like_6r_fixed()
    j = c.joints[6]
    c.joints[6] = JointDesc("wrist_4_joint", REVOLUTE, j.xyz, j.rot, (0, 0, 1))
    c.links[7] = LinkDesc("tool0", 0.5, (0, 0, 0.05), inertia=_cyl(0.5, 0.04, 0.1))
    c.name = "chain_7r"
    c.set_default_inputs()
    return c


def perturbed(chain: ChainDesc, seed: int = 1) -> ChainDesc:
    """Same geometry, but off-axis cogs, full SPD inertias and rotated inertial frames, so that no term of the
    dynamics is hidden by the UR10's axis-aligned zeros."""
    rng = np.random.RandomState(seed)
    c = copy.deepcopy(chain)
    for l in c.links[1:]:
        if l.mass == 0.0:
            l.mass = float(rng.uniform(0.1, 0.5))
        l.cog = tuple(float(v) for v in (np.asarray(l.cog) + rng.normal(0, 0.05, 3)))
        a = rng.normal(0, 1, (3, 3))
        spd = a @ a.T * 0.01 + np.eye(3) * 0.005
        l.inertia = (spd[0, 0], spd[0, 1], spd[0, 2], spd[1, 1], spd[1, 2], spd[2, 2])
        l.inertial_rot = tuple(rpy_to_rot(*rng.uniform(-1, 1, 3)))
    c.name = chain.name + "_perturbed"
    return c


def random_chain(seed: int, n_joints: int, p_prismatic: float = 0.25, p_fixed: float = 0.2,
                 gravity=GRAVITY) -> ChainDesc:
    """Seeded random serial chain: revolute / prismatic / (interior) fixed joints, random rpy origins and
    non-axis-aligned axes, full inertias."""
    rng = np.random.RandomState(seed)
    J: List[JointDesc] = []
    L: List[LinkDesc] = [LinkDesc("link0", 1.0)]
    for i in range(n_joints):
        u = rng.uniform()
        t = PRISMATIC if u < p_prismatic else (FIXED if u < p_prismatic + p_fixed else REVOLUTE)
        axis = rng.normal(0, 1, 3)
        if t == FIXED:
            axis = np.zeros(3)
        J.append(JointDesc(f"joint{i}", t, tuple(rng.uniform(-0.4, 0.4, 3)), tuple(rpy_to_rot(*rng.uniform(-math.pi, math.pi, 3))),
                           tuple(float(v) for v in axis)))
        a = rng.normal(0, 1, (3, 3))
        spd = a @ a.T * 0.02 + np.eye(3) * 0.01
        L.append(LinkDesc(f"link{i + 1}", float(rng.uniform(0.2, 8.0)), tuple(rng.normal(0, 0.1, 3)),
                          tuple(rpy_to_rot(*rng.uniform(-1, 1, 3))),
                          (spd[0, 0], spd[0, 1], spd[0, 2], spd[1, 1], spd[1, 2], spd[2, 2])))
    return ChainDesc(J, L, gravity, name=f"random_{seed}_{n_joints}")


def by_name(name: str) -> ChainDesc:
    table = {
        "ur10_like_6r_fixed": ur10_like_6r_fixed, "c6": ur10_like_6r_fixed,
        "chain_7r": chain_7r, "c7": chain_7r,
        "c6_perturbed": lambda: perturbed(ur10_like_6r_fixed(), 11),
        "c7_perturbed": lambda: perturbed(chain_7r(), 12),
        "random_a": lambda: random_chain(101, 7),
        "random_b": lambda: random_chain(202, 5, p_prismatic=0.5, p_fixed=0.2),
        "random_c": lambda: random_chain(303, 11, p_prismatic=0.2, p_fixed=0.3),
        "random_d": lambda: random_chain(404, 3, p_prismatic=0.34, p_fixed=0.0),
    }
    return table[name.lower()]()
