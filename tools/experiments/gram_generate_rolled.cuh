// experiment (round 2), see README.md: generator of the fused Gram kernel with the link loop rolled -- measured 2x slower
// (drop into rosdyn_b200/csrc/gram_common.cuh; call sites in gram_fused.cu guarded by GF_ROLLED)
// The same rows from a walk whose LINK loop is rolled (rigid-body mode).  The fully unrolled gram_generate is 60 KB of straight-line code for 6
// joints; a generator warp streams all of it once per group and runs 1.2 - 1.45x slower than its own static schedule from 40 KB on -- the
// instruction caches stop covering the stream (measured: 3 joints / 24 KB run at the schedule's speed, profiles/r02_gen_scaling.txt).  Here the
// body of one link is compiled once: the loop over the joints j already passed stays unrolled (their unit twists U[j], S[j] and torques live
// in registers, compile-time indices) and is guarded by warp-uniform tests on the running link index l; model constants come from the
// __grid_constant__ parameter through indexed constant-bank loads; the inputs of link l + 1 are requested while link l is computed.
// ~1.1 k instructions (18 KB) for 6 joints instead of 3.8 k.
template <int NJ, bool REV, int Z>
__device__ __forceinline__ void gram_generate_rolled(const ChainDev<NJ>& C, const SamplesDev& in, const double* __restrict__ tau_meas,
                                                     double* __restrict__ slot, int64_t i, int lane, double q_first, double dq_first,
                                                     double ddq_first)
{
  static_assert(!Z || REV, "the zero mass column of a link on its own joint exists for revolute joints only");
  using G = GramGeom<NJ, 0, Z>;
  V3 U[NJ], S[NJ];
  double tau[NJ];
#pragma unroll
  for (int j = 0; j < NJ; j++)
  {
    U[j] = v3(0, 0, 0);
    S[j] = v3(0, 0, 0);
    tau[j] = 0.0;
  }
  V3 v = v3(0, 0, 0), w = v3(0, 0, 0), a = v3(0, 0, 0), al = v3(0, 0, 0);
  V3 g = v3(C.g);
  double q_nx = q_first, dq_nx = dq_first, ddq_nx = ddq_first;
#pragma unroll 1
  for (int l = 0; l < NJ; l++)
  {
    const JointDev& J = C.joint[l];
    const double ql = q_nx, dql = dq_nx, ddql = ddq_nx;
    if (l + 1 < NJ)
    {
      const int inn = C.joint[l + 1].in;
      q_nx = ld_in(in.q, inn, in.ld, i);
      dq_nx = ld_in(in.dq, inn, in.ld, i);
      ddq_nx = ld_in(in.ddq, inn, in.ld, i);
    }
    double sv, cv;
    sincos_fast(ql, sv, cv);
    if (!(fabs(ql) <= 1.0e5)) sincos(ql, &sv, &cv);  // huge or non-finite angle: the library's Payne-Hanek reduction (rare)
    double R[9];
    V3 t = v3(J.t);
    if (REV || J.type == RDB_JOINT_REVOLUTE)
    {
      const double c1 = 1.0 - cv;
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = fma(c1, J.C[k], fma(sv, J.B[k], J.A[k]));
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = J.A[k];
      if (J.type == RDB_JOINT_PRISMATIC) t = axpy(t, v3(J.axp), ql);
    }
    const V3 axj = v3(J.ax);
    const V3 su = (!REV && J.type == RDB_JOINT_PRISMATIC) ? axj : v3(0, 0, 0);
    const V3 ss = (REV || J.type == RDB_JOINT_REVOLUTE) ? axj : v3(0, 0, 0);
    v = rotT(R, cross_add(v, w, t));
    w = rotT(R, w);
    a = rotT(R, cross_add(a, al, t));
    al = rotT(R, al);
    g = rotT(R, g);
    if (REV)
    {
      w = axpy(w, ss, dql);
      a = axpy(a, cross(v, ss), dql);
      al = axpy(axpy(al, cross(w, ss), dql), ss, ddql);
    }
    else
    {
      v = axpy(v, su, dql);
      w = axpy(w, ss, dql);
      const V3 xl = cross_add(cross(w, su), v, ss);
      const V3 xa = cross(w, ss);
      a = axpy(axpy(a, xl, dql), su, ddql);
      al = axpy(axpy(al, xa, dql), ss, ddql);
    }
    const double* Pl = C.link[l].pi;
    const V3 fm = cross_add(a - g, w, v);
    // positions of this link's block: posbase + p' ; the swizzle class of a position is (posbase + (p' & 3)) & 3
    const int posbase = 1 + 10 * (NJ - 1 - l);
    int lsw[4];
#pragma unroll
    for (int c = 0; c < 4; c++) lsw[c] = lane ^ (4 * ((posbase + c) & 3));
    double* const lb = slot + posbase * 32;
    // wrench regressor of link l projected on the unit twist (u, s) of joint j at this link; OWN: u == 0 exactly
    auto emit = [&](auto jc, auto ownc, V3 u, V3 s_) {
      constexpr int j = decltype(jc)::value;
      constexpr bool own = decltype(ownc)::value;
      const double e0 = own ? 0.0 : dot(u, fm);
      const V3 h = own ? cross(fm, s_) : cross_add(cross_add(cross(u, al), w, cross(w, u)), fm, s_);
      const V3 rho = cross(s_, w);
      double e[10];
      e[0] = e0;
      e[1] = h.x;
      e[2] = h.y;
      e[3] = h.z;
      e[4] = fma(s_.x, al.x, rho.x * w.x);
      e[5] = fma(s_.x, al.y, fma(s_.y, al.x, fma(rho.x, w.y, rho.y * w.x)));
      e[6] = fma(s_.x, al.z, fma(s_.z, al.x, fma(rho.x, w.z, rho.z * w.x)));
      e[7] = fma(s_.y, al.y, rho.y * w.y);
      e[8] = fma(s_.y, al.z, fma(s_.z, al.y, fma(rho.y, w.z, rho.z * w.y)));
      e[9] = fma(s_.z, al.z, rho.z * w.z);
      double t0 = tau[j], t1 = e[1] * Pl[1];
#pragma unroll
      for (int p = 0; p < 10; p += 2) t0 = fma(e[p], Pl[p], t0);
#pragma unroll
      for (int p = 3; p < 10; p += 2) t1 = fma(e[p], Pl[p], t1);
      tau[j] = t0 + t1;
      double* o = lb + G::rowbase(j);
#pragma unroll
      for (int p = 0; p < 10; p++)
      {
        const int pq = Z ? (p == 0 ? 9 : p - 1) : p;  // place of the column inside its link block (GramGeom::pos)
        if (!(Z && own && p == 0)) o[pq * 32 + lsw[pq & 3]] = e[p];
      }
    };
    static_for<0, NJ>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      if (j < l)
      {
        U[j] = rotT(R, cross_add(U[j], S[j], t));
        S[j] = rotT(R, S[j]);
        emit(jc, std::false_type{}, U[j], S[j]);
      }
      else if (j == l)
      {
        U[j] = su;
        S[j] = ss;
        if (REV) emit(jc, std::true_type{}, U[j], S[j]);
        else emit(jc, std::false_type{}, U[j], S[j]);
      }
    });
  }
#pragma unroll
  for (int j = 0; j < NJ; j++)
  {
    const double tv = tau_meas ? __ldcs(tau_meas + (int64_t)C.joint[j].in * in.ld + i) : tau[j];
    slot[G::rowbase(j) + lane] = tv;  // position 0
  }
}

