// gram_fused.cu -- fused regressor -> normal equations, one persistent warp-specialised kernel per GPU.
//
//   G (+)= sum_s Phi_s^T Phi_s ,  b (+)= sum_s Phi_s^T tau_s ,  tau_sq (+)= sum_s tau_s^T tau_s
//
// Phi never touches HBM.  Per CTA (1 per SM, 11 warps):
//   * 3 generator warps: one thread walks the chain of one sample (same link-frame recursion as dyn_kernel,
//     kernels.cu) and writes the augmented regressor rows [Phi_row | tau_row] of its 32 samples into one of
//     three shared-memory slots (only the structurally non-zero columns, XOR-swizzled, conflict free);
//   * 8 MMA warps: consume a slot as soon as it is full.  The contraction index k = (sample, joint row);
//     a k-step is 4 samples of one joint row, so the zero pattern of Phi (row of chain joint j is zero left
//     of column 10 j) is known at compile time and whole 8x8 tiles are skipped.  The upper-triangular tiles of the
//     (P+1)x(P+1) augmented Gram matrix (45 for P = 70) stay in registers for the whole kernel: the two MMA warps of
//     an SM sub-partition split them by tile-row parity (25 + 20 tiles), the four sub-partitions split the k-steps.
//     tcgen05 has no f64 kind: the FP64 tensor path of sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA).
//   * slots cycle through named barriers (full/empty), generation and DMMA overlap on the same FP64 pipes.
// Per-CTA partials are summed in a fixed order by gram_fused_reduce_kernel (bit-reproducible for a given n).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "spatial.cuh"

namespace rdb
{

#ifndef GF_TSPLIT
#define GF_TSPLIT 2
#endif
constexpr int GF_KSPLIT = 4;                       // MMA warps that share the k-steps of a slot (one per SM sub-partition)
constexpr int GF_TS = GF_TSPLIT;                   // 1: each MMA warp owns all tiles; 2: two warps per sub-partition split the tile rows by parity
constexpr int GF_MMA_WARPS = GF_KSPLIT * GF_TS;
constexpr int GF_GEN_WARPS = 3;  // == number of slots
constexpr int GF_THREADS = 32 * (GF_MMA_WARPS + GF_GEN_WARPS);
constexpr int GF_BAR_FULL = 1;   // named barriers 1..3: slot full ; 4..6: slot empty (0 is __syncthreads)
constexpr int GF_BAR_EMPTY = 1 + GF_GEN_WARPS;
constexpr int GF_BAR_COUNT = 32 + 32 * GF_MMA_WARPS;  // one generator warp + all MMA warps

struct GramRows
{
  int32_t base[8];      // offset (doubles) of the row of chain joint j inside a slot, -1 when the joint is not an input
  int32_t slot_doubles;  // doubles per slot
};

__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void dmma884f(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------- generator
// One sample per lane: getRegressor (+ getJointTorque) of sample i written to the slot (zeros when !active).
template <int NJ>
__device__ __forceinline__ void gram_generate(const ChainDev<NJ>& C, const GramRows& rows, const SamplesDev& in, const double* __restrict__ tau_meas,
                                              double* __restrict__ slot, int64_t i, bool active, int lane)
{
  constexpr int P = 10 * NJ;
  const double keep = active ? 1.0 : 0.0;
  V3 U[NJ], S[NJ];
  double tau[NJ];
  // all loads first: they depend on nothing but the sample index, so the DRAM latency is paid once, not once per joint
  double qv[NJ], dqv[NJ], ddqv[NJ];
#pragma unroll
  for (int l = 0; l < NJ; l++)
  {
    qv[l] = ld_in(in.q, C.joint[l].in, in.ld, i);
    dqv[l] = ld_in(in.dq, C.joint[l].in, in.ld, i);
    ddqv[l] = ld_in(in.ddq, C.joint[l].in, in.ld, i);
  }
  V3 v = v3(0, 0, 0), w = v3(0, 0, 0), a = v3(0, 0, 0), al = v3(0, 0, 0);
  V3 g = v3(C.g);
#pragma unroll
  for (int l = 0; l < NJ; l++)
  {
    const JointDev& J = C.joint[l];
    const double dql = dqv[l], ddql = ddqv[l];
    double R[9];
    V3 t = v3(J.t);
    if (J.type == RDB_JOINT_REVOLUTE)
    {
      // sincos stays inside the walk on purpose: hoisting all of them to the top measured 12 % slower end to end (the
      // branchy sincos bodies then run back to back instead of interleaving with the previous link's projections)
      double sv, cv;
      sincos(qv[l], &sv, &cv);
      const double c1 = 1.0 - cv;
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = fma(c1, J.C[k], fma(sv, J.B[k], J.A[k]));
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = J.A[k];
      if (J.type == RDB_JOINT_PRISMATIC) t = axpy(t, v3(J.axp), qv[l]);
    }
    const V3 axj = v3(J.ax);
    const V3 su = (J.type == RDB_JOINT_PRISMATIC) ? axj : v3(0, 0, 0);
    const V3 ss = (J.type == RDB_JOINT_REVOLUTE) ? axj : v3(0, 0, 0);
    v = rotT(R, cross_add(v, w, t));
    w = rotT(R, w);
    a = rotT(R, cross_add(a, al, t));
    al = rotT(R, al);
    g = rotT(R, g);
    v = axpy(v, su, dql);
    w = axpy(w, ss, dql);
    const V3 xl = cross_add(cross(w, su), v, ss);
    const V3 xa = cross(w, ss);
    a = axpy(axpy(a, xl, dql), su, ddql);
    al = axpy(axpy(al, xa, dql), ss, ddql);
#pragma unroll
    for (int j = 0; j < l; j++)
    {
      U[j] = rotT(R, cross_add(U[j], S[j], t));
      S[j] = rotT(R, S[j]);
    }
    U[l] = su;
    S[l] = ss;
    tau[l] = 0.0;
    const double* Pl = C.link[l].pi;
    const V3 fm = cross_add(a - g, w, v);
#pragma unroll
    for (int j = 0; j <= l; j++)
    {
      const V3 u = U[j], s = S[j];
      const double e0 = dot(u, fm);
      const V3 wu = cross(w, u);
      const V3 h = cross_add(cross_add(cross(u, al), w, wu), fm, s);
      const V3 rho = cross(s, w);
      double e[10];
      e[0] = e0;
      e[1] = h.x;
      e[2] = h.y;
      e[3] = h.z;
      e[4] = fma(s.x, al.x, rho.x * w.x);
      e[5] = fma(s.x, al.y, fma(s.y, al.x, fma(rho.x, w.y, rho.y * w.x)));
      e[6] = fma(s.x, al.z, fma(s.z, al.x, fma(rho.x, w.z, rho.z * w.x)));
      e[7] = fma(s.y, al.y, rho.y * w.y);
      e[8] = fma(s.y, al.z, fma(s.z, al.y, fma(rho.y, w.z, rho.z * w.y)));
      e[9] = fma(s.z, al.z, rho.z * w.z);
      double tj = tau[j];
#pragma unroll
      for (int p = 0; p < 10; p++) tj = fma(e[p], Pl[p], tj);
      tau[j] = tj;
      const int rb = rows.base[j];
      if (rb >= 0)
      {
        double* o = slot + rb + (10 * (l - j)) * 32;
#pragma unroll
        for (int p = 0; p < 10; p++) o[p * 32 + (lane ^ (4 * ((10 * l + p) & 3)))] = e[p] * keep;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; j++)
  {
    const int rb = rows.base[j];
    if (rb >= 0)
    {
      const double tv = tau_meas ? __ldcs(tau_meas + (int64_t)C.joint[j].in * in.ld + i) : tau[j];
      slot[rb + (P - 10 * j) * 32 + (lane ^ (4 * (P & 3)))] = tv * keep;
    }
  }
}

// ---------------------------------------------------------------------------------------------- MMA side
template <int NJ>
struct GramGeom
{
  static constexpr int P = 10 * NJ;
  static constexpr int T = (P + 1 + 7) / 8;       // tile columns of the augmented matrix
  static constexpr int NT = T * (T + 1) / 2;      // upper-triangular tiles
  __host__ __device__ static constexpr int tile(int I, int J) { return I * T - I * (I - 1) / 2 + (J - I); }
  // tile rows owned by an MMA warp: all (TS == 1) or the rows of parity `par` (TS == 2)
  __host__ __device__ static constexpr bool owns(int I, int ts, int par) { return ts == 1 || (I & 1) == par; }
  __host__ __device__ static constexpr int ntiles(int ts, int par)
  {
    int n = 0;
    for (int I = 0; I < T; I++)
      if (owns(I, ts, par)) n += T - I;
    return n;
  }
  __host__ __device__ static constexpr int local(int I, int J, int ts, int par)
  {
    int n = 0;
    for (int K = 0; K < I; K++)
      if (owns(K, ts, par)) n += T - K;
    return n + (J - I);
  }
};

// the k-steps of one slot that belong to k-split index `ks`, for the tile rows this warp owns
template <int NJ, int PAR>
__device__ __forceinline__ void gram_consume(const GramRows& rows, const double* __restrict__ slot, int ks, int lane,
                                             double (&acc)[GramGeom<NJ>::ntiles(GF_TS, PAR)][2])
{
  using G = GramGeom<NJ>;
  constexpr int P = G::P, T = G::T;
  const int g = lane >> 2, t = lane & 3;
  const int swz = 4 * (g & 3);
#pragma unroll
  for (int j = 0; j < NJ; j++)
  {
    const int rb = rows.base[j];
    if (rb < 0) continue;
    const int c0 = 10 * j;
    const int I0 = c0 / 8;
    const double* rowp = slot + rb;
#pragma unroll
    for (int kk = 0; kk < 8 / GF_KSPLIT; kk++)
    {
      const int s = 4 * (ks + GF_KSPLIT * kk) + t;
      double b[T];
#pragma unroll
      for (int J = 0; J < T; J++)
      {
        if (J < I0) continue;
        const int col = 8 * J + g;
        const bool all_valid = (8 * J >= c0) && (8 * J + 7 <= P);
        if (all_valid || (col >= c0 && col <= P)) b[J] = rowp[(col - c0) * 32 + (s ^ swz)];
        else b[J] = 0.0;
      }
#pragma unroll
      for (int I = 0; I < T; I++)
      {
        if (I < I0 || !G::owns(I, GF_TS, PAR)) continue;
#pragma unroll
        for (int J = I; J < T; J++) dmma884f(acc[G::local(I, J, GF_TS, PAR)][0], acc[G::local(I, J, GF_TS, PAR)][1], b[I], b[J]);
      }
    }
  }
}

template <int NJ, int PAR>
__device__ __forceinline__ void gram_mma_role(const GramRows& rows, const SamplesDev& in, double* smem, int ks, int lane, int dbg)
{
  using G = GramGeom<NJ>;
  constexpr int NTP = G::ntiles(GF_TS, PAR);
  double acc[NTP][2];
#pragma unroll
  for (int k = 0; k < NTP; k++) acc[k][0] = acc[k][1] = 0.0;
  const int64_t ngroups = (in.n + 31) / 32;
  const int64_t stride = (int64_t)gridDim.x * GF_GEN_WARPS;
  for (int64_t base = (int64_t)blockIdx.x * GF_GEN_WARPS; base < ngroups; base += stride)
  {
#pragma unroll 1
    for (int s = 0; s < GF_GEN_WARPS; s++)
    {
      if (base + s >= ngroups) break;
      bar_sync(GF_BAR_FULL + s, GF_BAR_COUNT);
      if (!(dbg & 2)) gram_consume<NJ, PAR>(rows, smem + (size_t)s * rows.slot_doubles, ks, lane, acc);
      if (base + s + stride < ngroups) bar_arrive(GF_BAR_EMPTY + s, GF_BAR_COUNT);  // the generator will come back
    }
  }
  // fixed-order reduction over the k-split warps that own the same tiles, into shared memory (the slots are dead by now)
  bar_sync(7, 32 * GF_MMA_WARPS);
  const int g = lane >> 2, t = lane & 3;
  for (int w = 0; w < GF_KSPLIT; w++)
  {
    if (ks == w)
    {
#pragma unroll
      for (int I = 0; I < G::T; I++)
      {
        if (!G::owns(I, GF_TS, PAR)) continue;
#pragma unroll
        for (int J = I; J < G::T; J++)
        {
          double* o = smem + G::tile(I, J) * 64 + g * 8 + 2 * t;
          const int k = G::local(I, J, GF_TS, PAR);
          if (w == 0)
          {
            o[0] = acc[k][0];
            o[1] = acc[k][1];
          }
          else
          {
            o[0] += acc[k][0];
            o[1] += acc[k][1];
          }
        }
      }
    }
    bar_sync(7, 32 * GF_MMA_WARPS);
  }
}

template <int NJ>
__global__ void __launch_bounds__(GF_THREADS, 1)
    gram_fused_kernel(const __grid_constant__ ChainDev<NJ> C, const __grid_constant__ GramRows rows, const SamplesDev in,
                      const double* __restrict__ tau_meas, double* __restrict__ partial, const int dbg)
{
  using G = GramGeom<NJ>;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ngroups = (in.n + 31) / 32;
  // group of (iteration it, CTA, slot s): g = (it*gridDim.x + blockIdx.x)*GF_GEN_WARPS + s
  const int64_t stride = (int64_t)gridDim.x * GF_GEN_WARPS;

  const bool is_gen = warp >= GF_MMA_WARPS;  // (putting the generators first instead changes nothing: measured)
  const int gen_id = warp - GF_MMA_WARPS, mma_id = warp;
  if (is_gen)
  {
    // ------------------------------------------------ generator warp of slot s
    const int s = gen_id;
    double* slot = smem + (size_t)s * rows.slot_doubles;
    bool first = true;
    for (int64_t grp = (int64_t)blockIdx.x * GF_GEN_WARPS + s; grp < ngroups; grp += stride)
    {
      if (!first) bar_sync(GF_BAR_EMPTY + s, GF_BAR_COUNT);  // consumers released the slot
      first = false;
      const int64_t i = grp * 32 + lane;
      const bool active = i < in.n;
      if (!(dbg & 1)) gram_generate<NJ>(C, rows, in, tau_meas, slot, active ? i : in.n - 1, active, lane);
      __threadfence_block();
      bar_arrive(GF_BAR_FULL + s, GF_BAR_COUNT);
    }
    return;
  }
  // ------------------------------------------------ MMA warps: k-split index = warp % 4 (its SM sub-partition), tile-row parity = warp / 4
  const int ks = mma_id % GF_KSPLIT;
  if (GF_TS == 1 || mma_id < GF_KSPLIT) gram_mma_role<NJ, 0>(rows, in, smem, ks, lane, dbg);
  else gram_mma_role<NJ, 1>(rows, in, smem, ks, lane, dbg);
  double* out = partial + (size_t)blockIdx.x * G::NT * 64;
  for (int k = mma_id * 32 + lane; k < G::NT * 64; k += 32 * GF_MMA_WARPS) out[k] = smem[k];
}

// fixed-order sum of the per-CTA partials -> gram (full symmetric, column-major), rhs, tau_sq
__global__ void gram_fused_reduce_kernel(const double* __restrict__ partial, int nparts, int T, int P, double* __restrict__ gram,
                                         double* __restrict__ rhs, double* __restrict__ tau_sq, int accumulate)
{
  const int NT = T * (T + 1) / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= NT * 64) return;
  double s = 0.0;
  for (int p = 0; p < nparts; p++) s += partial[(size_t)p * NT * 64 + e];
  int k = e >> 6, I = 0;
  while (k >= T - I)
  {
    k -= T - I;
    I++;
  }
  const int J = I + k;
  const int row = 8 * I + ((e >> 3) & 7), col = 8 * J + (e & 7);
  if (row > P || col > P) return;
  if (I == J && row > col) return;
  if (col < P)
  {
    const double v = accumulate ? gram[(size_t)col * P + row] + s : s;
    gram[(size_t)col * P + row] = v;
    if (row != col) gram[(size_t)row * P + col] = v;
  }
  else if (row < P)
    rhs[row] = accumulate ? rhs[row] + s : s;
  else if (tau_sq)
    *tau_sq = accumulate ? *tau_sq + s : s;
}

template <int NJ>
static ChainDev<NJ> narrow_g(const ChainDev<RDB_MAX_JOINTS>& h)
{
  ChainDev<NJ> c;
  c.nj = h.nj;
  c.n_in = h.n_in;
  for (int k = 0; k < 3; k++) c.g[k] = h.g[k];
  for (int j = 0; j < NJ; j++)
  {
    c.joint[j] = h.joint[j];
    c.link[j] = h.link[j];
  }
  return c;
}

template <int NJ>
static cudaError_t launch_fused_nj(ChainHost& ch, const GramRows& rows, const SamplesDev& in, const double* tau_meas, double* gram, double* rhs,
                                   double* tau_sq, int accumulate, cudaStream_t st)
{
  using G = GramGeom<NJ>;
  const size_t smem = sizeof(double) * (size_t)std::max(rows.slot_doubles * GF_GEN_WARPS, G::NT * 64);
  {
    cudaError_t e = cudaFuncSetAttribute(gram_fused_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  static const int dbg = [] { const char* e = getenv("RDB_GRAM_DEBUG"); return e ? atoi(e) : 0; }();  // 1: skip generation, 2: skip MMA (timing experiments only)
  const int64_t ngroups = (in.n + 31) / 32;
  const int grid = (int)std::min<int64_t>(ch.sm_count, (ngroups + GF_GEN_WARPS - 1) / GF_GEN_WARPS);
  const size_t need = sizeof(double) * (size_t)ch.sm_count * G::NT * 64;
  if (ch.gram.fused_bytes < need)
  {
    if (ch.gram.fused_partials) cudaFree(ch.gram.fused_partials);
    ch.gram.fused_partials = nullptr;
    ch.gram.fused_bytes = 0;
    cudaError_t e = cudaMalloc(&ch.gram.fused_partials, need);
    if (e != cudaSuccess) return e;
    ch.gram.fused_bytes = need;
  }
  gram_fused_kernel<NJ><<<grid, GF_THREADS, smem, st>>>(narrow_g<NJ>(ch.host), rows, in, tau_meas, ch.gram.fused_partials, dbg);
  count_launch();
  gram_fused_reduce_kernel<<<(G::NT * 64 + 255) / 256, 256, 0, st>>>(ch.gram.fused_partials, grid, G::T, G::P, gram, rhs, tau_sq, accumulate);
  count_launch();
  return cudaGetLastError();
}

// returns cudaErrorNotSupported when the chain does not fit the fused kernel (caller falls back to the v0 pipeline)
cudaError_t launch_gram_fused(ChainHost& ch, const SamplesDev& in, const double* tau_meas, double* gram, double* rhs, double* tau_sq,
                              int accumulate, cudaStream_t st)
{
  const int nj = ch.host.nj;
  if (nj < 1 || nj > 7 || in.n <= 0) return cudaErrorNotSupported;
  GramRows rows;
  int off = 0;
  const int P = 10 * nj;
  for (int j = 0; j < 8; j++)
  {
    rows.base[j] = -1;
    if (j < nj && ch.host.joint[j].in >= 0)
    {
      rows.base[j] = off;
      off += (P + 1 - 10 * j) * 32;
    }
  }
  rows.slot_doubles = off;
  if (off == 0 || sizeof(double) * (size_t)off * GF_GEN_WARPS > 227 * 1024) return cudaErrorNotSupported;
  switch (nj)
  {
#define X(N) \
  case N: return launch_fused_nj<N>(ch, rows, in, tau_meas, gram, rhs, tau_sq, accumulate, st);
    X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#undef X
  }
  return cudaErrorNotSupported;
}

}  // namespace rdb
