"""CPU-only: the C-ABI library loads, exports every symbol include/rosdyn_b200.h declares, and refuses to
compute without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from rosdyn_b200 import _lib, fixtures
from rosdyn_b200.descriptor import to_ctypes


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "rosdyn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rdb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rosdyn_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    assert lib.rdb_abi_version() == 2


def test_struct_layouts_match_header():
    # sizes the C side was compiled with: rdb_joint_desc 2*4 + 15*8, rdb_link_desc 19*8, rdb_samples 6*8, rdb_kinematics_out 12*8 + layout,
    # rdb_dynamics_out 4*8 + layout
    from rosdyn_b200.descriptor import CChainDesc, CJointDesc, CLinkDesc
    assert ctypes.sizeof(CJointDesc) == 128 and ctypes.sizeof(CLinkDesc) == 152
    assert ctypes.sizeof(CChainDesc) == 8 + 24 + 16
    assert ctypes.sizeof(_lib.CSamples) == 48 and ctypes.sizeof(_lib.CKinematicsOut) == 104 and ctypes.sizeof(_lib.CDynamicsOut) == 40


def test_host_generator_matches_oracle():
    from oracle.oracle import fill_uniform
    lib = _lib.load()
    x = np.empty((7, 50))
    lib.rdb_fill_uniform_host(ctypes.c_void_p(x.ctypes.data), 7, 50, 50, 0x5EED0003, 1)
    assert np.array_equal(x, fill_uniform(7, 50, 0x5EED0003, 1))


def test_documented_generator_formula():
    """include/rosdyn_b200.h states the generator in closed form; a reference-side harness that implements exactly that text gets the same bits."""
    M = (1 << 64) - 1

    def splitmix64(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
        return x ^ (x >> 31)
    lib = _lib.load()
    seed, sid, planes, n = 0x5EED0004, 2, 5, 40
    x = np.empty((planes, n))
    lib.rdb_fill_uniform_host(ctypes.c_void_p(x.ctypes.data), planes, n, n, seed, sid)
    for p in range(planes):
        for i in range(n):
            z = splitmix64((seed + 256 * i + 64 * sid + p) & M)
            assert x[p, i] == 2.0 * ((z >> 11) * 2.0 ** -53) - 1.0
    hdr = open(os.path.join(ROOT, "include", "rosdyn_b200.h")).read()
    assert "seed + 256*i + 64*stream_id + plane" in hdr


def test_group_entries_without_a_device():
    """The NCCL group entries validate their arguments and refuse to run without a CUDA device; NCCL itself is bound at run time."""
    lib = _lib.load()
    g = ctypes.c_void_p()
    c, keep = to_ctypes(fixtures.by_name("c6"))
    assert lib.rdb_group_create(None, 1, None, ctypes.byref(g)) == _lib.RDB_ERR_INVALID_ARG
    assert lib.rdb_group_create(ctypes.byref(c), 0, None, ctypes.byref(g)) == _lib.RDB_ERR_INVALID_ARG
    assert lib.rdb_group_create_rank(ctypes.byref(c), 0, 2, 2, None, ctypes.byref(g)) == _lib.RDB_ERR_INVALID_ARG
    assert lib.rdb_group_size(None) == -1 and lib.rdb_group_ranks(None) == -1 and lib.rdb_group_chain(None, 0) is None
    if lib.rdb_device_count() == 0:
        assert lib.rdb_group_create(ctypes.byref(c), 2, None, ctypes.byref(g)) == _lib.RDB_ERR_NO_DEVICE
        assert lib.rdb_group_create_rank(ctypes.byref(c), 0, 1, 0, None, ctypes.byref(g)) == _lib.RDB_ERR_NO_DEVICE
    lib.rdb_group_destroy(None)


def test_hostile_inputs_are_refused():
    """a deeply nested XML document, multi-turn enumeration far outside the limits and non-finite joint values."""
    lib = _lib.load()
    deep = ("<a>" * 300000).encode()
    out = ctypes.POINTER(_lib.CUrdfChain)()
    assert lib.rdb_urdf_parse(deep, b"b", b"t", None, ctypes.byref(out)) == _lib.RDB_ERR_INVALID_ARG
    assert b"nested" in lib.rdb_last_error()
    from rosdyn_b200.chain import multiplicity
    assert multiplicity([1], [0.5], [-1.0], [1.0]).shape == (1, 1)
    with pytest.raises(_lib.RosdynB200Error):
        multiplicity([1], [-3.0e7], [-1.0], [1.0])
    with pytest.raises(_lib.RosdynB200Error):
        multiplicity([1], [float("nan")], [-1.0], [1.0])
    with pytest.raises(_lib.RosdynB200Error):   # 9 joints x 9 images = 3.9e8 vectors
        multiplicity([1] * 9, [0.0] * 9, [-26.0] * 9, [26.0] * 9)


def test_argument_errors_do_not_need_a_device():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.rdb_chain_create(None, ctypes.byref(h)) == _lib.RDB_ERR_INVALID_ARG
    d = fixtures.by_name("c6")
    d.joints[0].input_index = 99
    c, keep = to_ctypes(d)
    assert lib.rdb_chain_create(ctypes.byref(c), ctypes.byref(h)) == _lib.RDB_ERR_INVALID_ARG
    assert b"input_index" in lib.rdb_last_error()


def test_no_cpu_fallback_without_device():
    lib = _lib.load()
    if lib.rdb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    c, keep = to_ctypes(fixtures.by_name("c6"))
    h = ctypes.c_void_p()
    assert lib.rdb_chain_create(ctypes.byref(c), ctypes.byref(h)) == _lib.RDB_ERR_NO_DEVICE
    from rosdyn_b200.chain import Chain
    with pytest.raises(_lib.RosdynB200Error):
        Chain(fixtures.by_name("c6"))
    v = ctypes.c_double()
    assert lib.rdb_fp64_peak(1, 1, ctypes.byref(v)) == _lib.RDB_ERR_NO_DEVICE


def test_descriptor_host_logic():
    from rosdyn_b200.descriptor import rpy_to_rot
    R = np.array(rpy_to_rot(0.3, -0.7, 1.9)).reshape(3, 3)
    cr, sr, cp, sp, cy, sy = np.cos(0.3), np.sin(0.3), np.cos(-0.7), np.sin(-0.7), np.cos(1.9), np.sin(1.9)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    np.testing.assert_allclose(R, Rz @ Ry @ Rx, atol=1e-15)
    d = fixtures.by_name("c6")
    assert (d.n_joints, d.n_links, d.n_inputs) == (7, 8, 6)
    assert [j.input_index for j in d.joints] == [0, 1, 2, 3, 4, 5, -1]
    assert fixtures.by_name("c7").n_inputs == 7
