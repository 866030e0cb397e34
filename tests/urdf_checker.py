"""TEST INFRASTRUCTURE: an independent restatement of the reference's URDF -> chain extraction with Python's xml.etree
(Link::fromUrdf / Joint::fromUrdf / Chain::init, primitives_impl.h:50-149, 276-331, 580-703), used only to check csrc/urdf.cpp."""
import math
import xml.etree.ElementTree as ET

from rosdyn_b200.descriptor import FIXED, PRISMATIC, REVOLUTE, ChainDesc, JointDesc, LinkDesc, rpy_to_rot


def _vec(s, default):
    return tuple(float(x) for x in s.split()) if s else default


def chain_from_urdf(text, base, tool, gravity=(0.0, 0.0, 0.0)):
    root = ET.fromstring(text)
    links, joint_of_child = {}, {}
    for l in root.findall("link"):
        d = LinkDesc(l.get("name"))
        inr = l.find("inertial")
        if inr is not None:
            o = inr.find("origin")
            d.cog = _vec(o.get("xyz") if o is not None else None, (0.0, 0.0, 0.0))
            d.inertial_rot = tuple(rpy_to_rot(*_vec(o.get("rpy") if o is not None else None, (0.0, 0.0, 0.0))))
            d.mass = float(inr.find("mass").get("value"))
            I = inr.find("inertia")
            d.inertia = tuple(float(I.get(k, 0.0)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz"))
        links[d.name] = d
    for j in root.findall("joint"):
        t = j.get("type")
        typ = REVOLUTE if t in ("revolute", "continuous") else (PRISMATIC if t == "prismatic" else FIXED)
        o = j.find("origin")
        ax = j.find("axis")
        jd = JointDesc(j.get("name"), typ, _vec(o.get("xyz") if o is not None else None, (0.0, 0.0, 0.0)),
                       tuple(rpy_to_rot(*_vec(o.get("rpy") if o is not None else None, (0.0, 0.0, 0.0)))),
                       _vec(ax.get("xyz") if ax is not None else None, (1.0, 0.0, 0.0)))
        lim = j.find("limit")
        lims = dict(q_max=0.0, q_min=0.0, Dq_max=0.0, DDq_max=0.0, tau_max=0.0)
        if t in ("revolute", "prismatic"):
            if lim is None:
                lims.update(q_max=1e10, q_min=-1e10, tau_max=1e10)
            else:
                hi, lo = float(lim.get("upper", 0)), float(lim.get("lower", 0))
                if hi <= lo:
                    hi, lo = 2 * math.pi, -2 * math.pi
                vel = float(lim.get("velocity", 0))
                if vel <= 0:
                    vel = 2 * math.pi
                lims.update(q_max=hi, q_min=lo, Dq_max=vel, DDq_max=10 * vel, tau_max=float(lim.get("effort", 0)))
        elif t == "continuous":
            lims.update(q_max=1e10, q_min=-1e10)
            if lim is None:
                lims.update(tau_max=1e10)
            else:
                vel = float(lim.get("velocity", 0))
                lims.update(Dq_max=vel, DDq_max=10 * vel, tau_max=float(lim.get("effort", 0)))
        joint_of_child[j.find("child").get("link")] = (jd, j.find("parent").get("link"), lims)
    if base not in links:
        raise LookupError("Base link not found")
    if tool not in links:
        raise LookupError("Tool link not found")
    chain, act = [], tool
    while act != base:
        if act not in joint_of_child:
            raise LookupError("Tool link not found")
        jd, parent, lims = joint_of_child[act]
        chain.append((jd, act, lims))
        act = parent
    chain.reverse()
    d = ChainDesc([c[0] for c in chain], [links[base]] + [links[c[1]] for c in chain], tuple(gravity))
    d.limits = {c[0].name: c[2] for c in chain}
    return d
