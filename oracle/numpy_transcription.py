"""TEST INFRASTRUCTURE.  Independent numpy transcription of the reference's Chain hot path.

Second, separately written restatement of the same reference lines as oracle/rosdyn_oracle.c, expressed with
numpy matrices the way the reference expresses them with Eigen (4x4 Affine products, 6x6 spatial inertia
matrices, `Matrix610d` wrench regressors).  It exists to pin the C oracle: tests/golden/*.npz are produced
by THIS file (tests/golden/make_golden.py) and the C oracle, then the CUDA engine, are compared with them.
The reference itself cannot be imported or built here (C++/Eigen/ROS, see oracle/rosdyn_oracle.c header), so
parity stays "unpinned by the reference's own vectors"; the two restatements + invariants are the pin.

Citations: SA.h = rosdyn_core/include/rosdyn_core/spacevect_algebra.h,
           PI.h = rosdyn_core/include/rosdyn_core/internal/primitives_impl.h.
One sample at a time, pure numpy; slow on purpose (small cases only).
"""
from __future__ import annotations

import numpy as np

FIXED, REVOLUTE, PRISMATIC = 0, 1, 2


def skew(v):  # SA.h:69-76
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def spatialCrossProduct(a, b):  # SA.h:88-93
    r = np.zeros(6)
    r[3:] = np.cross(a[3:], b[3:])
    r[:3] = np.cross(a[3:], b[:3]) + np.cross(a[:3], b[3:])
    return r


def spatialDualCrossProduct(a, w):  # SA.h:108-113
    r = np.zeros(6)
    r[3:] = np.cross(a[3:], w[3:]) + np.cross(a[:3], w[:3])
    r[:3] = np.cross(a[3:], w[:3])
    return r


def spatialTranslation(t, d):  # SA.h:129-133
    r = t.copy()
    r[:3] = t[:3] + np.cross(t[3:], d)
    return r


def spatialDualTranslation(w, d):  # SA.h:150-154
    r = w.copy()
    r[3:] = w[3:] + np.cross(w[:3], d)
    return r


def spatialRotation(x, R):  # SA.h:172-175
    return np.concatenate([R @ x[:3], R @ x[3:]])


def spatialTranformation(x, T):  # SA.h:193-197
    R, p = T[:3, :3], T[:3, 3]
    return np.concatenate([R @ x[:3] + np.cross(R @ x[3:], p), R @ x[3:]])


def computeSpatialInertiaMatrix(inertia, cog, mass):  # SA.h:232-239
    cs = skew(cog)
    S = np.zeros((6, 6))
    S[:3, :3] = mass * np.eye(3)
    S[:3, 3:] = mass * cs.T
    S[3:, :3] = mass * cs
    S[3:, 3:] = inertia + mass * (cs @ cs.T)
    return S


class NpJoint:
    def __init__(self, jd):  # Joint::fromUrdf PI.h:50-83
        self.type = jd.type
        self.input_index = jd.input_index
        self.T_pj = np.eye(4)
        self.T_pj[:3, :3] = np.asarray(jd.rot, dtype=np.float64).reshape(3, 3)
        self.T_pj[:3, 3] = np.asarray(jd.xyz, dtype=np.float64)
        ax = np.asarray(jd.axis, dtype=np.float64)
        if np.linalg.norm(ax) > 0:
            ax = ax / np.linalg.norm(ax)
        self.axis_in_j = ax
        self.skew_axis_in_j = skew(ax)
        self.square_skew_axis_in_j = self.skew_axis_in_j @ self.skew_axis_in_j
        self.R_pj = self.T_pj[:3, :3].copy()
        self.axis_in_p = self.R_pj @ self.axis_in_j
        self.screw_of_c_in_p = np.zeros(6)  # computeJacobian PI.h:25-35
        if self.type == REVOLUTE:
            self.screw_of_c_in_p[3:] = self.axis_in_p
        elif self.type == PRISMATIC:
            self.screw_of_c_in_p[:3] = self.axis_in_p

    def getTransformation(self, q):  # computedTpc PI.h:38-47
        T = self.T_pj.copy()
        if self.type == REVOLUTE:
            R_jc = np.eye(3) + np.sin(q) * self.skew_axis_in_j + (1 - np.cos(q)) * self.square_skew_axis_in_j
            T[:3, :3] = self.R_pj @ R_jc
        elif self.type == PRISMATIC:
            T[:3, 3] = self.T_pj[:3, 3] + self.axis_in_p * q
        return T


class NpLink:
    def __init__(self, ld):  # Link::fromUrdf PI.h:288-396
        self.mass = float(ld.mass)
        self.cog = np.asarray(ld.cog, dtype=np.float64)
        ixx, ixy, ixz, iyy, iyz, izz = [float(v) for v in ld.inertia]
        inertia = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
        Rin = np.asarray(ld.inertial_rot, dtype=np.float64).reshape(3, 3)
        inertia = Rin @ inertia @ Rin.T
        self.Inertia_cc = computeSpatialInertiaMatrix(inertia, self.cog, self.mass)
        E = [np.zeros((6, 6)) for _ in range(10)]
        E[0][:3, :3] = np.eye(3)
        for k in range(3):
            e = np.zeros(3)
            e[k] = 1
            E[1 + k][:3, 3:] = skew(e).T
            E[1 + k][3:, :3] = skew(e)
        E[4][3, 3] = 1
        E[5][3, 4] = E[5][4, 3] = 1
        E[6][3, 5] = E[6][5, 3] = 1
        E[7][4, 4] = 1
        E[8][4, 5] = E[8][5, 4] = 1
        E[9][5, 5] = 1
        self.Inertia_cc_single_term = E

    def getNominalParameters(self):  # PI.h:399-417
        I0 = self.Inertia_cc[3:, 3:]
        return np.concatenate([[self.mass], self.cog * self.mass, [I0[0, 0], I0[0, 1], I0[0, 2], I0[1, 1], I0[1, 2], I0[2, 2]]])


class NpChain:
    """Stateless image of rosdyn::Chain: every call is a fresh evaluation through the direct paths."""

    def __init__(self, desc):
        self.joints = [NpJoint(j) for j in desc.joints]
        self.links = [NpLink(l) for l in desc.links]
        self.nJ = len(self.joints)
        self.nL = self.nJ + 1
        self.n_in = desc.n_inputs
        self.gravity = np.asarray(desc.gravity, dtype=np.float64)
        # m_input_to_chain_joint PI.h:708-731
        self.S = np.zeros((self.nJ, self.n_in))
        for nj, j in enumerate(self.joints):
            if j.input_index >= 0:
                self.S[nj, j.input_index] = 1

    # --- kinematics
    def _frames(self, q):  # computeFrames PI.h:863-872, computeScrews PI.h:874-882
        sq = self.S @ q
        T = [np.eye(4)]
        for nl in range(1, self.nL):
            T.append(T[nl - 1] @ self.joints[nl - 1].getTransformation(sq[nl - 1]))
        s = [np.zeros(6)]
        for nl in range(1, self.nL):
            s.append(spatialRotation(self.joints[nl - 1].screw_of_c_in_p, T[nl - 1][:3, :3]))
        return T, s

    def getTransformations(self, q):
        return self._frames(q)[0]

    def getJacobian(self, q):  # PI.h:939-945
        T, s = self._frames(q)
        jac = np.zeros((6, self.n_in))
        for nj, j in enumerate(self.joints):
            if j.input_index >= 0 and j.type != FIXED:
                jac[:, j.input_index] = spatialTranslation(s[nj + 1], T[-1][:3, 3] - T[nj + 1][:3, 3])
        return jac

    def getTwist(self, q, Dq, _fs=None):  # PI.h:1004-1009
        T, s = _fs or self._frames(q)
        sDq = self.S @ Dq
        v = [np.zeros(6)]
        for nl in range(1, self.nL):
            v.append(spatialTranslation(v[nl - 1], T[nl][:3, 3] - T[nl - 1][:3, 3]) + s[nl] * sDq[nl - 1])
        return v

    def getDTwistLinearPart(self, q, DDq):  # PI.h:1052-1057
        T, s = self._frames(q)
        sDDq = self.S @ DDq
        a = [np.zeros(6)]
        for nl in range(1, self.nL):
            a.append(spatialTranslation(a[nl - 1], T[nl][:3, 3] - T[nl - 1][:3, 3]) + s[nl] * sDDq[nl - 1])
        return a

    def getDTwistNonLinearPart(self, q, Dq):  # PI.h:1071-1076
        T, s = self._frames(q)
        v = self.getTwist(q, Dq, (T, s))
        sDq = self.S @ Dq
        a = [np.zeros(6)]
        for nl in range(1, self.nL):
            a.append(spatialTranslation(a[nl - 1], T[nl][:3, 3] - T[nl - 1][:3, 3]) + spatialCrossProduct(v[nl], s[nl]) * sDq[nl - 1])
        return a

    def getDTwist(self, q, Dq, DDq, _fs=None):  # direct path PI.h:1113-1118
        T, s = _fs or self._frames(q)
        v = self.getTwist(q, Dq, (T, s))
        sDq, sDDq = self.S @ Dq, self.S @ DDq
        a = [np.zeros(6)]
        for nl in range(1, self.nL):
            a.append(spatialTranslation(a[nl - 1], T[nl][:3, 3] - T[nl - 1][:3, 3]) +
                     spatialCrossProduct(v[nl], s[nl]) * sDq[nl - 1] + s[nl] * sDDq[nl - 1])
        return a

    def _jerk(self, q, Dq, DDq, DDDq, lin, nonlin):  # PI.h:1145-1150, 1171-1179, 1210-1219
        T, s = self._frames(q)
        v = self.getTwist(q, Dq, (T, s))
        a = self.getDTwist(q, Dq, DDq, (T, s))
        sDq, sDDq, sDDDq = self.S @ Dq, self.S @ DDq, self.S @ DDDq
        j = [np.zeros(6)]
        for nl in range(1, self.nL):
            nj = nl - 1
            v_cross_s = spatialCrossProduct(v[nl], s[nl])
            x = spatialTranslation(j[nl - 1], T[nl][:3, 3] - T[nl - 1][:3, 3])
            if lin:
                x = x + s[nl] * sDDDq[nj]
            if nonlin:
                x = x + v_cross_s * sDDq[nj] + (spatialCrossProduct(a[nl], s[nl]) + spatialCrossProduct(v[nl], v_cross_s)) * sDq[nj]
            j.append(x)
        return j

    def getDDTwist(self, q, Dq, DDq, DDDq):
        return self._jerk(q, Dq, DDq, DDDq, True, True)

    def getDDTwistLinearPart(self, q, DDDq):
        z = np.zeros(self.n_in)
        return self._jerk(q, z, z, DDDq, True, False)

    def getDDTwistNonLinearPart(self, q, Dq, DDq):
        return self._jerk(q, Dq, DDq, np.zeros(self.n_in), False, True)

    # --- dynamics
    def getWrench(self, q, Dq, DDq, ext=None):  # PI.h:1231-1258
        T, s = self._frames(q)
        v = self.getTwist(q, Dq, (T, s))
        a = self.getDTwist(q, Dq, DDq, (T, s))
        if ext is None:
            ext = [np.zeros(6) for _ in range(self.nL)]
        w = [None] * self.nL
        for nl in range(self.nL - 1, -1, -1):
            if nl == 0:
                inertial = np.zeros(6)
                grav = np.zeros(6)
            else:
                R = T[nl][:3, :3]
                I = self.links[nl].Inertia_cc
                inertial = spatialRotation(I @ spatialRotation(a[nl], R.T) +
                                           spatialDualCrossProduct(spatialRotation(v[nl], R.T), I @ spatialRotation(v[nl], R.T)), R)
                grav = np.zeros(6)
                grav[:3] = -self.links[nl].mass * self.gravity
                grav[3:] = -np.cross(R @ self.links[nl].cog, self.links[nl].mass * self.gravity)
            if nl < self.nL - 1:
                w[nl] = spatialTranformation(-ext[nl], T[nl]) + inertial + grav + \
                    spatialDualTranslation(w[nl + 1], T[nl][:3, 3] - T[nl + 1][:3, 3])
            else:
                w[nl] = spatialTranformation(-ext[nl], T[nl]) + inertial + grav
        return w, s

    def getJointTorque(self, q, Dq, DDq, ext=None):  # PI.h:1267-1273
        w, s = self.getWrench(q, Dq, DDq, ext)
        tau = np.array([w[nj + 1] @ s[nj + 1] for nj in range(self.nJ)])
        return self.S.T @ tau

    def getJointTorqueNonLinearPart(self, q, Dq):  # PI.h:1285-1293
        return self.getJointTorque(q, Dq, np.zeros(self.n_in))

    def getRegressor(self, q, Dq, DDq):  # PI.h:1321-1352
        T, s = self._frames(q)
        v = self.getTwist(q, Dq, (T, s))
        a = self.getDTwist(q, Dq, DDq, (T, s))
        W = [np.zeros((6, 10)) for _ in range(self.nL)]
        Rext = np.zeros((self.nJ, 10 * self.nJ))
        for nl in range(self.nL - 1, 0, -1):
            R = T[nl][:3, :3]
            for ip in range(10):
                E = self.links[nl].Inertia_cc_single_term[ip]
                W[nl][:, ip] = spatialRotation(E @ spatialRotation(a[nl], R.T) +
                                               spatialDualCrossProduct(spatialRotation(v[nl], R.T), E @ spatialRotation(v[nl], R.T)), R)
            W[nl][:3, 0] -= self.gravity
            W[nl][3:, 1] -= np.cross(R @ np.array([1.0, 0, 0]), self.gravity)
            W[nl][3:, 2] -= np.cross(R @ np.array([0, 1.0, 0]), self.gravity)
            W[nl][3:, 3] -= np.cross(R @ np.array([0, 0, 1.0]), self.gravity)
            Rext[nl - 1, (nl - 1) * 10:(nl - 1) * 10 + 10] = s[nl] @ W[nl]
            for nlf in range(nl + 1, self.nL):
                for ip in range(10):
                    Rext[nl - 1, (nlf - 1) * 10 + ip] = s[nl] @ spatialDualTranslation(W[nlf][:, ip], T[nl][:3, 3] - T[nlf][:3, 3])
        return (Rext.T @ self.S).T

    def getJointInertia(self, q):  # PI.h:1361-1377
        T, s = self._frames(q)
        Mext = np.zeros((self.nJ, self.nJ))
        for nj in range(self.nJ):
            Jn = np.zeros((6, self.nJ))
            for ij in range(nj + 1):
                il = ij + 1
                if self.joints[ij].type != FIXED:
                    c = spatialTranslation(s[il], T[nj + 1][:3, 3] - T[il][:3, 3])
                    Jn[:, ij] = spatialRotation(c, T[nj + 1][:3, :3].T)
            Mext += Jn.T @ self.links[nj + 1].Inertia_cc @ Jn
        return self.S.T @ Mext @ self.S

    def getNominalParameters(self):  # PI.h:1382-1391
        return np.concatenate([self.links[nl].getNominalParameters() for nl in range(1, self.nL)])


def splitmix64(x: int) -> int:
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & m
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & m
    return x ^ (x >> 31)


def fill_uniform(n_planes: int, n: int, seed: int, stream_id: int) -> np.ndarray:
    """Same generator as rdb_fill_uniform_host (include/rosdyn_b200.h): x[j][i] in U(-1,1)."""
    x = np.empty((n_planes, n))
    for j in range(n_planes):
        for i in range(n):
            z = splitmix64((seed + (i << 8) + (stream_id << 6) + j) & ((1 << 64) - 1))
            x[j, i] = 2.0 * ((z >> 11) * (1.0 / 9007199254740992.0)) - 1.0
    return x
