"""URDF string -> ChainDesc through the library's own loader (rdb_urdf_parse, csrc/urdf.cpp): the drop-in for
`rosdyn::createChain(urdf_model, base_frame, tool_frame, gravity)` (primitives_impl.h:1518-1527) without urdfdom.
Host only - needs no GPU."""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

from . import _lib
from .descriptor import ChainDesc, JointDesc, LinkDesc


def chain_from_urdf(urdf_xml: str, base_link: str, tool_link: str, gravity: Optional[Sequence[float]] = None) -> ChainDesc:
    """Descriptor of the chain base_link -> tool_link.  Raises LookupError("Base link not found" / "Tool link not found")
    like the reference's constructor, ValueError on malformed XML.  `desc.limits` holds per-joint q/Dq/DDq/tau limits."""
    lib = _lib.load()
    out = ctypes.POINTER(_lib.CUrdfChain)()
    g = (ctypes.c_double * 3)(*[float(v) for v in gravity]) if gravity is not None else None
    st = lib.rdb_urdf_parse(urdf_xml.encode(), base_link.encode(), tool_link.encode(), g, ctypes.byref(out))
    if st == _lib.RDB_ERR_INVALID_ARG:
        raise ValueError(lib.rdb_last_error().decode())
    _lib.check(st)
    try:
        u = out.contents
        nj = u.desc.n_joints
        joints, links = [], []
        limits: Dict[str, Dict[str, float]] = {}
        for k in range(nj):
            j = u.desc.joints[k]
            name = u.joint_names[k].decode()
            joints.append(JointDesc(name, int(j.type), tuple(j.xyz), tuple(j.rot), tuple(j.axis), int(j.input_index)))
            limits[name] = {"q_max": u.q_max[k], "q_min": u.q_min[k], "Dq_max": u.dq_max[k], "DDq_max": u.ddq_max[k], "tau_max": u.tau_max[k]}
        for k in range(nj + 1):
            l = u.desc.links[k]
            links.append(LinkDesc(u.link_names[k].decode(), float(l.mass), tuple(l.cog), tuple(l.inertial_rot), tuple(l.inertia)))
        desc = ChainDesc(joints, links, tuple(u.desc.gravity), name=f"{base_link}->{tool_link}", n_inputs=int(u.desc.n_inputs))
        desc.limits = limits
        return desc
    finally:
        lib.rdb_urdf_chain_free(out)
