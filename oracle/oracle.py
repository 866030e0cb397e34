"""TEST INFRASTRUCTURE: ctypes loader of the CPU restatement oracle (oracle/rosdyn_oracle.c).

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from rosdyn_b200.descriptor import CChainDesc, ChainDesc, to_ctypes  # interface structs only

_HERE = os.path.dirname(os.path.abspath(__file__))
_dbl_p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> None:
    """Compile the C restatement (gcc); a no-op when the .so is newer than the source."""
    so = os.path.join(_HERE, "librosdyn_oracle.so")
    src = os.path.join(_HERE, "rosdyn_oracle.c")
    srcs = [src, os.path.join(_HERE, "box_qp.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dbl_p)


REF_LIB = os.path.join(_HERE, "_ref", "librosdyn_ref.so")
REFERENCE_ROOT = "/root/reference/rosdyn_core/include"


def build_ref(force: bool = False) -> bool:
    """Compile the REFERENCE's own headers (where they lie under /root/reference) against the stand-in third-party headers of
    oracle/shim/ into oracle/_ref/librosdyn_ref.so.  Returns False when the reference tree is absent (the GPU box only uses
    the prebuilt file)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return os.path.exists(REF_LIB)
    deps = [os.path.join(_HERE, "ref_driver.cpp"), os.path.join(_HERE, "shim", "mini_eigen.h"), os.path.join(_HERE, "box_qp.h"),
            os.path.join(_HERE, "shim", "ros", "ros.h"), os.path.join(_HERE, "shim", "eigen_matrix_utils", "eiquadprog.hpp"),
            os.path.join(_HERE, "..", "include", "rosdyn_b200", "rosdyn_core_bridge.h")]
    if force or not os.path.exists(REF_LIB) or any(os.path.getmtime(REF_LIB) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return True


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


class _Lib:
    def __init__(self, fast=False):
        """fast: False = restatement (-O3), True = restatement with the reference's -Ofast test flags, "ref" = the reference's own
        headers compiled against oracle/shim (oracle/_ref/librosdyn_ref.so)."""
        if fast == "ref":
            path = REF_LIB
            if not os.path.exists(path) and not build_ref():
                raise FileNotFoundError(path)
        else:
            name = "librosdyn_oracle_fast.so" if fast else "librosdyn_oracle.so"
            path = os.path.join(_HERE, name)
            if not os.path.exists(path):
                build()
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.oracle_chain_create.restype = ctypes.c_void_p
        L.oracle_chain_create.argtypes = [ctypes.POINTER(CChainDesc)]
        L.oracle_chain_destroy.argtypes = [ctypes.c_void_p]
        L.oracle_nominal_parameters.argtypes = [ctypes.c_void_p, _dbl_p]
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_kind.restype = ctypes.c_char_p
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        L.oracle_kinematics_batch.argtypes = [vp, i64, i64] + [_dbl_p] * 4 + [i64] + [_dbl_p] * 11 + [ci]
        L.oracle_regressor_torque_batch.argtypes = [vp, i64, i64, _dbl_p, _dbl_p, _dbl_p, i64, _dbl_p, _dbl_p, ci]
        L.oracle_inertia_batch.argtypes = [vp, i64, i64, _dbl_p, i64, _dbl_p, ci]
        L.oracle_regressor_gram.argtypes = [vp, i64, i64, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p, _dbl_p]
        if fast != "ref":
            L.oracle_fill_uniform.argtypes = [_dbl_p, ci, i64, i64, ctypes.c_uint64, ci]


_libs = {}


def lib(fast=False) -> _Lib:
    if fast not in _libs:
        _libs[fast] = _Lib(fast)
    return _libs[fast]


class CComponentDesc(ctypes.Structure):  # image of rdb_component_desc (include/rosdyn_b200.h)
    _fields_ = [("type", ctypes.c_int32), ("input_index", ctypes.c_int32), ("min_velocity", ctypes.c_double), ("max_velocity", ctypes.c_double)]


def components_regressor(components, n_in: int, q, dq, fast=False) -> np.ndarray:
    """components: [(type 1|2|3, input_index, min_velocity, max_velocity), ...] -> phi_c[Pc*n_in][N] (plane col*n_in+row).
    fast="ref" runs the reference's own component classes (friction_polynomial1.h, friction_polynomial2.h, ideal_spring.h)."""
    L = lib(fast).lib
    arr = (CComponentDesc * max(len(components), 1))()
    for k, (t, j, lo, hi) in enumerate(components):
        arr[k].type, arr[k].input_index, arr[k].min_velocity, arr[k].max_velocity = int(t), int(j), float(lo), float(hi)
    Pc = sum(3 if int(c[0]) == 2 else 2 for c in components)
    q, dq = _c(q, n_in), _c(dq, n_in)
    n = q.shape[1]
    out = np.full((Pc * n_in, n), np.nan)
    L.oracle_components_regressor_batch.argtypes = [ctypes.c_int, ctypes.POINTER(CComponentDesc), ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                    _dbl_p, _dbl_p, ctypes.c_int64, _dbl_p]
    L.oracle_components_regressor_batch(len(components), arr, n_in, n, n, _ptr(q), _ptr(dq), n, _ptr(out))
    return out


def fill_uniform(n_planes: int, n: int, seed: int, stream_id: int) -> np.ndarray:
    x = np.empty((n_planes, n))
    lib().lib.oracle_fill_uniform(_ptr(x), n_planes, n, n, seed, stream_id)
    return x


def _c(a, rows=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and (rows is None or a.shape[0] == rows), (a.shape, rows)
    return a


class OracleChain:
    """CPU restatement of rosdyn::Chain, batched over SoA arrays x[component][N] (same layout as the C-ABI).
    `OracleChain(desc, fast="ref")` drives the reference's own code instead (oracle/ref_driver.cpp), same methods."""

    KIN_FIELDS = ("T_tool", "T_links", "jacobian", "twist", "dtwist", "dtwist_lin", "dtwist_nonlin", "ddtwist",
                  "ddtwist_lin", "ddtwist_nonlin", "torque")

    def __init__(self, desc: ChainDesc, fast=False):
        self._l = lib(fast)
        cdesc, keep = to_ctypes(desc)
        self._h = self._l.lib.oracle_chain_create(ctypes.byref(cdesc))
        if not self._h:
            raise ValueError("oracle_chain_create failed")
        del keep
        self.nJ, self.nL, self.n_in = desc.n_joints, desc.n_joints + 1, desc.n_inputs

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.lib.oracle_chain_destroy(self._h)
            self._h = None

    @staticmethod
    def max_threads() -> int:
        return int(lib().lib.oracle_max_threads())

    def nominal_parameters(self) -> np.ndarray:
        out = np.zeros(10 * self.nJ)
        self._l.lib.oracle_nominal_parameters(self._h, _ptr(out))
        return out

    def kinematics(self, q, dq=None, ddq=None, dddq=None, want=KIN_FIELDS, nthreads=1):
        q, dq, ddq, dddq = (_c(x, self.n_in) for x in (q, dq, ddq, dddq))
        n = q.shape[1]
        rows = {"T_tool": 12, "T_links": 12 * self.nL, "jacobian": 6 * self.n_in, "torque": self.n_in}
        out = {k: (np.zeros((rows.get(k, 6 * self.nL), n)) if k in want else None) for k in self.KIN_FIELDS}
        self._l.lib.oracle_kinematics_batch(self._h, n, n, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(dddq), n,
                                            *[_ptr(out[k]) for k in self.KIN_FIELDS], nthreads)
        return {k: v for k, v in out.items() if v is not None}

    def regressor_torque(self, q, dq, ddq, nthreads=1, store=True):
        q, dq, ddq = (_c(x, self.n_in) for x in (q, dq, ddq))
        n = q.shape[1]
        phi = np.zeros((10 * self.nJ * self.n_in, n)) if store else None
        tau = np.zeros((self.n_in, n)) if store else None
        self._l.lib.oracle_regressor_torque_batch(self._h, n, n, _ptr(q), _ptr(dq), _ptr(ddq), n, _ptr(phi), _ptr(tau), nthreads)
        return phi, tau

    def inertia(self, q, nthreads=1):
        q = _c(q, self.n_in)
        n = q.shape[1]
        M = np.zeros((self.n_in * self.n_in, n))
        self._l.lib.oracle_inertia_batch(self._h, n, n, _ptr(q), n, _ptr(M), nthreads)
        return M

    def wrench(self, q, dq, ddq, ext=None):
        """getJointTorque(q,Dq,DDq,ext_wrenches_in_link_frame) and getWrench: (torque[n_in][N], wrenches[6 nL][N])."""
        q, dq, ddq = (_c(x, self.n_in) for x in (q, dq, ddq))
        ext = _c(ext, 6 * self.nL)
        n = q.shape[1]
        tau, w = np.zeros((self.n_in, n)), np.zeros((6 * self.nL, n))
        f = self._l.lib.oracle_wrench_batch
        f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _dbl_p, _dbl_p, _dbl_p, _dbl_p, ctypes.c_int64, ctypes.c_int64, _dbl_p, _dbl_p]
        f(self._h, n, n, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(ext), n, n, _ptr(tau), _ptr(w))
        return tau, w

    def jacobian_link(self, q, link: int):
        q = _c(q, self.n_in)
        n = q.shape[1]
        J = np.zeros((6 * self.n_in, n))
        f = self._l.lib.oracle_jacobian_link_batch
        f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _dbl_p, ctypes.c_int, ctypes.c_int64, _dbl_p]
        f(self._h, n, n, _ptr(q), int(link), n, _ptr(J))
        return J

    def local_ik(self, target, seed, q_min=None, q_max=None, weight=None, toll=1e-6, max_iter=50):
        """computeLocalIk / computeWeigthedLocalIk (PI.h:1398-1468) with an iteration budget: (sol[n_in][N], status[N], iters[N], err[N])."""
        target, seed = _c(target, 12), _c(seed, self.n_in)
        n = target.shape[1]
        qmin = None if q_min is None else np.ascontiguousarray(q_min, dtype=np.float64)
        qmax = None if q_max is None else np.ascontiguousarray(q_max, dtype=np.float64)
        w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        sol, err = np.zeros((self.n_in, n)), np.zeros(n)
        status, iters = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
        f = self._l.lib.oracle_local_ik_batch
        i32p = ctypes.POINTER(ctypes.c_int32)
        f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64] + [_dbl_p] * 5 + [ctypes.c_double, ctypes.c_int, _dbl_p, i32p, i32p, _dbl_p]
        f(self._h, n, n, _ptr(target), _ptr(seed), _ptr(qmin), _ptr(qmax), _ptr(w), float(toll), int(max_iter), _ptr(sol),
          status.ctypes.data_as(i32p), iters.ctypes.data_as(i32p), _ptr(err))
        return sol, status, iters, err

    def gram(self, q, dq, ddq, tau_meas=None):
        q, dq, ddq, tau_meas = (_c(x, self.n_in) for x in (q, dq, ddq, tau_meas))
        n = q.shape[1]
        P = 10 * self.nJ
        G, b, tt = np.zeros((P, P)), np.zeros(P), np.zeros(1)
        self._l.lib.oracle_regressor_gram(self._h, n, n, _ptr(q), _ptr(dq), _ptr(ddq), _ptr(tau_meas), _ptr(G), _ptr(b), _ptr(tt))
        return G, b, float(tt[0])  # G is symmetric, so column-major == row-major
