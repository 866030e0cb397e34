"""Flat chain descriptor: the Python image of `rdb_chain_desc` (include/rosdyn_b200.h).

A descriptor holds exactly what `Joint::fromUrdf` / `Link::fromUrdf` read from the URDF
(reference: rosdyn_core/include/rosdyn_core/internal/primitives_impl.h:50-83, 288-319) for the joints and
links of one serial chain base->tool, fixed joints included (primitives_impl.h:615-639), plus the
input-joint selection of `Chain::setInputJointsName` (primitives_impl.h:705-742) and gravity.

Pure data + ctypes marshalling; no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from typing import List, Sequence

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2
MAX_JOINTS = 64


def rpy_to_rot(roll: float, pitch: float, yaw: float) -> List[float]:
    """URDF rpy -> 3x3 rotation, row-major, through the quaternion exactly as the reference does:
    urdfdom `Rotation::setFromRPY` (un-vendored third party, published formula) followed by
    `Eigen::Quaterniond -> Affine3d` (urdf_parser.h:44-50).  R = Rz(yaw) Ry(pitch) Rx(roll)."""
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / s, y / s, z / s, w / s
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return [1 - (tyy + tzz), txy - twz, txz + twy,
            txy + twz, 1 - (txx + tzz), tyz - twx,
            txz - twy, tyz + twx, 1 - (txx + tyy)]


IDENTITY3 = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]


@dataclass
class JointDesc:
    name: str
    type: int
    xyz: Sequence[float] = (0.0, 0.0, 0.0)
    rot: Sequence[float] = tuple(IDENTITY3)  # row-major R_pj
    axis: Sequence[float] = (1.0, 0.0, 0.0)  # URDF default axis
    input_index: int = -1


@dataclass
class LinkDesc:
    name: str
    mass: float = 0.0
    cog: Sequence[float] = (0.0, 0.0, 0.0)
    inertial_rot: Sequence[float] = tuple(IDENTITY3)
    inertia: Sequence[float] = (0.0, 0.0, 0.0, 0.0, 0.0, 0.0)  # ixx ixy ixz iyy iyz izz


@dataclass
class ChainDesc:
    joints: List[JointDesc]
    links: List[LinkDesc]  # len(joints) + 1, links[0] = base
    gravity: Sequence[float] = (0.0, 0.0, 0.0)  # ctor default is zero (primitives.h:346)
    name: str = "chain"
    n_inputs: int = field(default=-1)

    def __post_init__(self):
        if len(self.links) != len(self.joints) + 1:
            raise ValueError("a chain of nJ joints has nJ+1 links")
        if len(self.joints) > MAX_JOINTS:
            raise ValueError("too many joints")
        if self.n_inputs < 0:
            self.set_default_inputs()

    # Chain::init -> setInputJointsName(m_moveable_joints_name): non-fixed joints, base->tool (PI.h:631-636,700)
    def set_default_inputs(self):
        k = 0
        for j in self.joints:
            if j.type != FIXED:
                j.input_index = k
                k += 1
            else:
                j.input_index = -1
        self.n_inputs = k

    # Chain::setInputJointsName (PI.h:705-742): unknown names are skipped (they get no column of S)
    def set_input_joints(self, names: Sequence[str]) -> bool:
        ok = True
        for j in self.joints:
            j.input_index = -1
        by_name = {j.name: j for j in self.joints}
        for idx, nm in enumerate(names):
            if nm in by_name:
                by_name[nm].input_index = idx
            else:
                ok = False
        self.n_inputs = len(names)
        return ok

    @property
    def n_joints(self) -> int:
        return len(self.joints)

    @property
    def n_links(self) -> int:
        return len(self.links)


# ---------------------------------------------------------------------------------------------- ctypes image
class CJointDesc(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("input_index", ctypes.c_int32), ("xyz", ctypes.c_double * 3),
                ("rot", ctypes.c_double * 9), ("axis", ctypes.c_double * 3)]


class CLinkDesc(ctypes.Structure):
    _fields_ = [("mass", ctypes.c_double), ("cog", ctypes.c_double * 3), ("inertial_rot", ctypes.c_double * 9),
                ("inertia", ctypes.c_double * 6)]


class CChainDesc(ctypes.Structure):
    _fields_ = [("n_joints", ctypes.c_int32), ("n_inputs", ctypes.c_int32), ("gravity", ctypes.c_double * 3),
                ("joints", ctypes.POINTER(CJointDesc)), ("links", ctypes.POINTER(CLinkDesc))]


def to_ctypes(desc: ChainDesc):
    """Returns (CChainDesc, keepalive) - keep `keepalive` referenced while the struct is in use."""
    nj = desc.n_joints
    joints = (CJointDesc * max(nj, 1))()
    links = (CLinkDesc * (nj + 1))()
    for i, j in enumerate(desc.joints):
        joints[i].type = int(j.type)
        joints[i].input_index = int(j.input_index)
        joints[i].xyz[:] = [float(v) for v in j.xyz]
        joints[i].rot[:] = [float(v) for v in j.rot]
        joints[i].axis[:] = [float(v) for v in j.axis]
    for i, l in enumerate(desc.links):
        links[i].mass = float(l.mass)
        links[i].cog[:] = [float(v) for v in l.cog]
        links[i].inertial_rot[:] = [float(v) for v in l.inertial_rot]
        links[i].inertia[:] = [float(v) for v in l.inertia]
    c = CChainDesc()
    c.n_joints = nj
    c.n_inputs = int(desc.n_inputs)
    c.gravity[:] = [float(v) for v in desc.gravity]
    c.joints = ctypes.cast(joints, ctypes.POINTER(CJointDesc))
    c.links = ctypes.cast(links, ctypes.POINTER(CLinkDesc))
    return c, (joints, links)
