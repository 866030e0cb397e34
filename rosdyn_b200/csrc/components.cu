// components.cu -- additive joint components as extra regressor columns.
// Element-wise in q_j / Dq_j: FirstOrderPolynomialFriction::computeRegressor (friction_polynomial1.h:45-52),
// SecondOrderPolynomialFriction::computeRegressor (friction_polynomial2.h:42-58), IdealSpring::getRegressor (ideal_spring.h:64-70).
// One thread per sample, coalesced plane loads / streaming plane stores like the chain walkers; HBM-write bound
// (8 * Pc * n_inputs bytes per sample, all but Pc of them zeros that the reference's dense m_regressor also holds).
#include <cuda_runtime.h>

#include "launch.h"
#include "spatial.cuh"

namespace rdb
{

__device__ __forceinline__ void component_values(const ComponentDev& c, const SamplesDev& in, int64_t i, double* v)
{
  if (c.type == RDB_COMPONENT_IDEAL_SPRING)
  {
    v[0] = ld_in(in.q, c.in, in.ld, i);
    v[1] = 1.0;
    v[2] = 0.0;
    return;
  }
  const double dq = ld_in(in.dq, c.in, in.ld, i);
  const double omega = fmin(fmax(dq, -c.vmax), c.vmax);  // std::min(std::max(Dq, -m_Dq_max), m_Dq_max)
  if (c.type == RDB_COMPONENT_FRICTION_POLY1)
  {
    v[0] = fmin(fmax(omega / c.thr, -1.0), 1.0);
    v[1] = omega;
    v[2] = 0.0;
  }
  else
  {
    double sg;
    if (omega == 0) sg = 0;
    else if (omega > c.thr) sg = 1.0;
    else if (omega < -c.thr) sg = -1.0;
    else sg = omega / c.thr;
    v[0] = sg;
    v[1] = omega;
    v[2] = omega * omega * sg;  // pow(omega, 2.0) * sign_Dq
  }
}

__global__ void __launch_bounds__(256) components_regressor_kernel(const __grid_constant__ ComponentsDev C, const SamplesDev in, const int n_in,
                                                                   double* __restrict__ phi_c, const int64_t ld_out)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  for (int k = 0; k < C.n; k++)
  {
    const ComponentDev& c = C.c[k];
    double v[3];
    component_values(c, in, i, v);
    for (int p = 0; p < c.ncols; p++)
      for (int r = 0; r < n_in; r++) __stcs(phi_c + ((int64_t)(c.col + p) * n_in + r) * ld_out + i, r == c.in ? v[p] : 0.0);
  }
}

__global__ void __launch_bounds__(256) components_torque_kernel(const __grid_constant__ ComponentsDev C, const __grid_constant__ ComponentParams prm,
                                                                const SamplesDev in, const int n_in, double* __restrict__ torque,
                                                                const int64_t ld_out, const int accumulate)
{
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in.n) return;
  if (!accumulate)
    for (int r = 0; r < n_in; r++) torque[(int64_t)r * ld_out + i] = 0.0;
  for (int k = 0; k < C.n; k++)
  {
    const ComponentDev& c = C.c[k];
    double v[3];
    component_values(c, in, i, v);
    double t = 0.0;
    for (int p = 0; p < c.ncols; p++) t = fma(v[p], prm.p[c.col + p], t);  // m_regressor.row(j) * m_nominal_parameters
    torque[(int64_t)c.in * ld_out + i] += t;
  }
}

cudaError_t launch_components_regressor(const ChainHost& ch, const SamplesDev& in, double* phi_c, int64_t ld_out, cudaStream_t st)
{
  if (in.n <= 0 || ch.comps.n == 0) return cudaSuccess;
  components_regressor_kernel<<<(unsigned)((in.n + 255) / 256), 256, 0, st>>>(ch.comps, in, ch.host.n_in, phi_c, ld_out);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_components_torque(const ChainHost& ch, const SamplesDev& in, const ComponentParams& prm, double* torque, int64_t ld_out,
                                     int accumulate, cudaStream_t st)
{
  if (in.n <= 0) return cudaSuccess;
  components_torque_kernel<<<(unsigned)((in.n + 255) / 256), 256, 0, st>>>(ch.comps, prm, in, ch.host.n_in, torque, ld_out, accumulate);
  count_launch();
  return cudaGetLastError();
}

}  // namespace rdb
