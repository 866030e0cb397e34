// spatial.cuh -- fp64 3-vector helpers and the per-joint transform for the per-sample chain walkers.
// Conventions follow the reference's spacevect_algebra.h (6-vectors = [linear; angular]); the device code
// keeps the two halves in separate V3 registers.
#pragma once
#include "chain_dev.h"

namespace rdb
{

struct V3
{
  double x, y, z;
};

__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 v3(const double* p) { return V3{p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return fma(a.x, b.x, fma(a.y, b.y, a.z * b.z)); }
// a + b*s
__device__ __forceinline__ V3 axpy(V3 a, V3 b, double s) { return V3{fma(b.x, s, a.x), fma(b.y, s, a.y), fma(b.z, s, a.z)}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b)
{
  return V3{fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x))};
}
// c + a x b
__device__ __forceinline__ V3 cross_add(V3 c, V3 a, V3 b)
{
  return V3{fma(a.y, b.z, fma(-a.z, b.y, c.x)), fma(a.z, b.x, fma(-a.x, b.z, c.y)), fma(a.x, b.y, fma(-a.y, b.x, c.z))};
}
// R x   (R row-major)
__device__ __forceinline__ V3 rot(const double* R, V3 a)
{
  return V3{fma(R[0], a.x, fma(R[1], a.y, R[2] * a.z)), fma(R[3], a.x, fma(R[4], a.y, R[5] * a.z)),
            fma(R[6], a.x, fma(R[7], a.y, R[8] * a.z))};
}
// R^T x
__device__ __forceinline__ V3 rotT(const double* R, V3 a)
{
  return V3{fma(R[0], a.x, fma(R[3], a.y, R[6] * a.z)), fma(R[1], a.x, fma(R[4], a.y, R[7] * a.z)),
            fma(R[2], a.x, fma(R[5], a.y, R[8] * a.z))};
}
// C = A B (3x3 row-major)
__device__ __forceinline__ void mul33(const double* A, const double* B, double* C)
{
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = fma(A[3 * i], B[j], fma(A[3 * i + 1], B[3 + j], A[3 * i + 2] * B[6 + j]));
}

// sin and cos, branch free (|x| <= 1e5: 3-term Cody-Waite reduction by pi/2 with exact FMA products, then the fdlibm kernel polynomials on
// |r| <= pi/4; <= 1 ulp like the library).  The walkers evaluate it for all joints of a sample up front, so that the whole walk is ONE
// basic block and the scheduler can overlap the serial transform chain of link l+1 with the projections of link l.  Huge or non-finite
// angles take the library sincos (Payne-Hanek) in a single, rarely executed branch (trig_all).
__device__ __forceinline__ void sincos_fast(double x, double& s, double& c)
{
  const double j = rint(x * 6.36619772367581382433e-01);
  double r = fma(-j, 1.57079632679489655800e+00, x);
  r = fma(-j, 6.12323399573676603587e-17, r);
  r = fma(-j, -1.49738490485916983294e-33, r);  // pi/2 = hi + mid + lo (lo is negative)
  const int q = (int)j;
  const double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double sr = fma(z * r, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
  const double ss = (q & 1) ? cr : sr, cc = (q & 1) ? sr : cr;
  s = (q & 2) ? -ss : ss;
  c = ((q + 1) & 2) ? -cc : cc;
}

template <int N>
__device__ __forceinline__ void trig_all(const double (&q)[N], double (&sv)[N], double (&cv)[N])
{
  bool big = false;
#pragma unroll
  for (int l = 0; l < N; l++)
  {
    sincos_fast(q[l], sv[l], cv[l]);
    big |= !(fabs(q[l]) <= 1.0e5);
  }
  if (big)
  {
#pragma unroll
    for (int l = 0; l < N; l++) sincos(q[l], &sv[l], &cv[l]);
  }
}

// Joint::computedTpc (primitives_impl.h:38-47): parent<-child rotation R (row-major) and translation t (parent frame).
__device__ __forceinline__ void joint_transform(const JointDev& J, double q, double* R, V3& t)
{
  t = v3(J.t);
  if (J.type == RDB_JOINT_REVOLUTE)
  {
    double s, c;
    sincos(q, &s, &c);
    const double c1 = 1.0 - c;
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = fma(c1, J.C[k], fma(s, J.B[k], J.A[k]));
  }
  else
  {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = J.A[k];
    if (J.type == RDB_JOINT_PRISMATIC) t = axpy(t, v3(J.axp), q);
  }
}

// same with sin q / cos q already known
__device__ __forceinline__ void joint_transform_sc(const JointDev& J, double q, double s, double c, double* R, V3& t)
{
  t = v3(J.t);
  if (J.type == RDB_JOINT_REVOLUTE)
  {
    const double c1 = 1.0 - c;
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = fma(c1, J.C[k], fma(s, J.B[k], J.A[k]));
  }
  else
  {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = J.A[k];
    if (J.type == RDB_JOINT_PRISMATIC) t = axpy(t, v3(J.axp), q);
  }
}

// streaming (evict-first) plane accesses: every input is read once and every output written once
__device__ __forceinline__ double ld_in(const double* p, int in, int64_t ld, int64_t i)
{
  return (p != nullptr && in >= 0) ? __ldcs(p + (int64_t)in * ld + i) : 0.0;
}
// L2 prefetch of an input that is needed later in the walk: the real load then finds the line on chip
__device__ __forceinline__ void prefetch_in(const double* p, int in, int64_t ld, int64_t i)
{
  if (p != nullptr && in >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (int64_t)in * ld + i));
}
__device__ __forceinline__ void st_out(double* p, int64_t plane, int64_t ld, int64_t i, double v) { __stcs(p + plane * ld + i, v); }
__device__ __forceinline__ void st3(double* p, int64_t plane, int64_t ld, int64_t i, V3 v)
{
  __stcs(p + plane * ld + i, v.x);
  __stcs(p + (plane + 1) * ld + i, v.y);
  __stcs(p + (plane + 2) * ld + i, v.z);
}

}  // namespace rdb
