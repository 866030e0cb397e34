"""CPU-only: the C++ headers a reference-side maintainer would use compile, and the bridge is numerically right.

* include/rosdyn_b200/chain.hpp: the Eigen-typed overloads (the reference's signatures, primitives.h:452-548) compile and link against the
  Eigen subset of oracle/shim (Eigen3 itself is not installed here) -- tests/cpp/eigen_facade_compile.cpp; the -m gpu suite runs the binary.
* include/rosdyn_b200/rosdyn_core_bridge.h: rosdyn::toB200Desc, written against accessors the reference really has, is compiled into
  oracle/_ref with the reference's own headers and RUN on the reference's Chain objects: the descriptor it produces, fed to the CPU
  restatement, must reproduce the reference's outputs on the original chain."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import CHAINS, ROOT, assert_close
from rosdyn_b200 import fixtures
from rosdyn_b200.descriptor import CJointDesc, CLinkDesc, ChainDesc, JointDesc, LinkDesc


def build_eigen_facade(out):
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "eigen_facade_compile.cpp"), "-o", out, "-L", os.path.join(ROOT, "rosdyn_b200"), "-lrosdyn_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "rosdyn_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return out


def test_eigen_overloads_compile_and_link(tmp_path):
    exe = build_eigen_facade(str(tmp_path / "eigen_facade"))
    from rosdyn_b200 import _lib
    if _lib.load().rdb_device_count() == 0:   # without a device the binary only reports that it linked
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 0 and "compiled and linked only" in r.stdout


def test_drop_in_alias_compiles(tmp_path):
    src = tmp_path / "alias.cpp"
    src.write_text('#define ROSDYN_B200_DROP_IN\n#include "rosdyn_b200/chain.hpp"\n'
                   "double f(rosdyn::Chain& c, const Eigen::VectorXd& q, const Eigen::VectorXd& dq, const Eigen::VectorXd& ddq)\n"
                   "{ return c.getRegressor(q, dq, ddq)(0, 0) + c.getJointTorque(q, dq, ddq)(0) + c.getJointInertia(q)(0, 0); }\n")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "oracle", "shim"),
                        "-I", os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


@pytest.mark.parametrize("name", CHAINS)
def test_reference_side_bridge_round_trip(name):
    """reference Chain (oracle/_ref, the reference's own headers) -> rosdyn::toB200Desc -> descriptor -> restatement == reference outputs."""
    from oracle import oracle
    from oracle.oracle import OracleChain, fill_uniform
    if not oracle.have_ref() and not oracle.build_ref():
        pytest.skip("oracle/_ref is not built and /root/reference is not mounted")
    d = fixtures.by_name(name)
    ref = OracleChain(d, fast="ref")
    L = ref._l.lib
    if not hasattr(L, "oracle_bridge_descriptor"):
        pytest.skip("oracle/_ref predates the bridge entry")
    J = (CJointDesc * d.n_joints)()
    K = (CLinkDesc * d.n_links)()
    g = (ctypes.c_double * 3)()
    L.oracle_bridge_descriptor.restype = ctypes.c_int
    L.oracle_bridge_descriptor.argtypes = [ctypes.c_void_p, ctypes.POINTER(CJointDesc), ctypes.POINTER(CLinkDesc), ctypes.POINTER(ctypes.c_double)]
    n_in = L.oracle_bridge_descriptor(ref._h, J, K, g)
    assert n_in == d.n_inputs
    d2 = ChainDesc(joints=[JointDesc(name=f"j{k}", type=J[k].type, xyz=tuple(J[k].xyz), rot=tuple(J[k].rot), axis=tuple(J[k].axis),
                                     input_index=J[k].input_index) for k in range(d.n_joints)],
                   links=[LinkDesc(name=f"l{k}", mass=K[k].mass, cog=tuple(K[k].cog), inertial_rot=tuple(K[k].inertial_rot), inertia=tuple(K[k].inertia))
                          for k in range(d.n_links)],
                   gravity=tuple(g), name=name + "_bridged", n_inputs=n_in)
    assert [j.input_index for j in d2.joints] == [j.input_index if j.type != 0 else -1 for j in d.joints]
    assert [j.type for j in d2.joints] == [j.type for j in d.joints]
    oc = OracleChain(d2)
    n = 64
    q, dq, ddq, dddq = (fill_uniform(d.n_inputs, n, 0x5EED0000 + 77, s) for s in range(4))
    A, B = oc.kinematics(q, dq, ddq, dddq), ref.kinematics(q, dq, ddq, dddq)
    for k in A:
        assert_close(A[k], B[k], f"{name}:{k} through the bridge", 1e-12)
    pa, ta = oc.regressor_torque(q, dq, ddq)
    pb, tb = ref.regressor_torque(q, dq, ddq)
    assert_close(pa, pb, f"{name}: regressor through the bridge", 1e-12)
    assert_close(ta, tb, f"{name}: torque through the bridge", 1e-12)
    assert_close(oc.inertia(q), ref.inertia(q), f"{name}: inertia through the bridge", 1e-12)
    assert_close(oc.nominal_parameters(), ref.nominal_parameters(), f"{name}: nominal parameters through the bridge", 1e-12)
