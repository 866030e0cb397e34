// group_check.cpp -- the C-ABI's NCCL group from plain C++ (the consumer's language: rosdyn_identification is C++, reference README.md:15).
// One process, every GPU of the box: device-resident shards -> rdb_regressor_gram_sharded (fused kernel per device + ONE ncclAllReduce of the
// packed partials) -> compared on every device with the host-sum entry rdb_regressor_gram_sharded_host on the same samples.
// Build: tools/build_facade.py (g++, links librosdyn_b200.so and libcudart).  Usage: group_check [n_devices] [samples]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rosdyn_b200.h"

#define CK(x)                                                                  \
  do                                                                           \
  {                                                                            \
    if ((x) != RDB_OK)                                                         \
    {                                                                          \
      std::printf("FAIL %s: %s\n", #x, rdb_last_error());                      \
      return 1;                                                                \
    }                                                                          \
  } while (0)

static void rpy(double r, double p, double y, double R[9])
{
  const double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
  R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
  R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
  R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

int main(int argc, char** argv)
{
  int have = rdb_device_count();
  if (have <= 0)
  {
    std::printf("no CUDA device: rosdyn_b200 has no CPU fallback\n");
    return 2;
  }
  const int R = argc > 1 ? std::atoi(argv[1]) : have;
  const int64_t n = argc > 2 ? std::atoll(argv[2]) : 2000003;
  if (R > have)
  {
    std::printf("asked for %d devices, the box has %d\n", R, have);
    return 2;
  }
  // a 7-revolute chain with full inertias (made up; both entries see the same descriptor)
  const int NJ = 7;
  std::vector<rdb_joint_desc> J(NJ);
  std::vector<rdb_link_desc> L(NJ + 1);
  for (int j = 0; j < NJ; j++)
  {
    J[j] = rdb_joint_desc{};
    J[j].type = RDB_JOINT_REVOLUTE;
    J[j].input_index = j;
    J[j].xyz[0] = 0.05 * j; J[j].xyz[1] = 0.1 - 0.03 * j; J[j].xyz[2] = 0.2 + 0.01 * j;
    rpy(0.3 * j, 1.1 - 0.2 * j, 0.05 * j * j, J[j].rot);
    J[j].axis[j % 3] = 1.0; J[j].axis[(j + 1) % 3] = 0.3;
  }
  for (int l = 0; l <= NJ; l++)
  {
    L[l] = rdb_link_desc{};
    L[l].mass = 1.0 + 0.7 * l;
    L[l].cog[0] = 0.01 * l; L[l].cog[1] = -0.02; L[l].cog[2] = 0.05;
    rpy(0.1 * l, 0.2, -0.1 * l, L[l].inertial_rot);
    L[l].inertia[0] = 0.02 + 0.01 * l; L[l].inertia[1] = 0.001; L[l].inertia[2] = -0.002;
    L[l].inertia[3] = 0.03; L[l].inertia[4] = 0.0015; L[l].inertia[5] = 0.025 + 0.005 * l;
  }
  rdb_chain_desc d{};
  d.n_joints = NJ;
  d.n_inputs = NJ;
  d.gravity[2] = -9.806;
  d.joints = J.data();
  d.links = L.data();
  const int P = 10 * NJ;

  rdb_group* g = nullptr;
  CK(rdb_group_create(&d, R, nullptr, &g));
  std::printf("group: %d device(s), %d rank(s)\n", rdb_group_size(g), rdb_group_ranks(g));

  // shard r = samples [r n / R, (r + 1) n / R) of one seeded batch; the host copy has the same numbers (same generator, same indices)
  std::vector<double> hq((size_t)NJ * n), hdq((size_t)NJ * n), hddq((size_t)NJ * n);
  rdb_fill_uniform_host(hq.data(), NJ, n, n, 0x5EED0042, 0);
  rdb_fill_uniform_host(hdq.data(), NJ, n, n, 0x5EED0042, 1);
  rdb_fill_uniform_host(hddq.data(), NJ, n, n, 0x5EED0042, 2);
  std::vector<rdb_samples> shards(R);
  std::vector<double*> dG(R), db(R), dt(R), dq(R), ddq(R), dddq(R);
  for (int r = 0; r < R; r++)
  {
    const int64_t lo = (int64_t)r * n / R, hi = (int64_t)(r + 1) * n / R, m = hi - lo;
    cudaSetDevice(r);
    cudaMalloc(&dq[r], sizeof(double) * NJ * m);
    cudaMalloc(&ddq[r], sizeof(double) * NJ * m);
    cudaMalloc(&dddq[r], sizeof(double) * NJ * m);
    cudaMalloc(&dG[r], sizeof(double) * P * P);
    cudaMalloc(&db[r], sizeof(double) * P);
    cudaMalloc(&dt[r], sizeof(double));
    cudaMemcpy2D(dq[r], m * 8, hq.data() + lo, n * 8, m * 8, NJ, cudaMemcpyHostToDevice);
    cudaMemcpy2D(ddq[r], m * 8, hdq.data() + lo, n * 8, m * 8, NJ, cudaMemcpyHostToDevice);
    cudaMemcpy2D(dddq[r], m * 8, hddq.data() + lo, n * 8, m * 8, NJ, cudaMemcpyHostToDevice);
    shards[r] = rdb_samples{m, m, dq[r], ddq[r], dddq[r], nullptr};
  }
  CK(rdb_regressor_gram_sharded(g, shards.data(), nullptr, dG.data(), db.data(), dt.data(), 0, nullptr));
  CK(rdb_group_synchronize(g));

  // host-sum entry on the same samples
  std::vector<rdb_chain*> hs(R);
  for (int r = 0; r < R; r++) hs[r] = rdb_group_chain(g, r);
  std::vector<double> G0((size_t)P * P), b0(P);
  double t0 = 0;
  rdb_samples all{n, n, hq.data(), hdq.data(), hddq.data(), nullptr};
  CK(rdb_regressor_gram_sharded_host(hs.data(), R, &all, nullptr, G0.data(), b0.data(), &t0, 0));

  int fail = 0;
  double scale = 0;
  for (double v : G0) scale = std::fmax(scale, std::fabs(v));
  for (int r = 0; r < R; r++)
  {
    std::vector<double> G((size_t)P * P), b(P);
    double t = 0;
    cudaSetDevice(r);
    cudaMemcpy(G.data(), dG[r], sizeof(double) * P * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), db[r], sizeof(double) * P, cudaMemcpyDeviceToHost);
    cudaMemcpy(&t, dt[r], sizeof(double), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (size_t k = 0; k < G.size(); k++) worst = std::fmax(worst, std::fabs(G[k] - G0[k]));
    const bool ok = worst <= 1e-12 * scale && std::fabs(t - t0) <= 1e-12 * t0;
    std::printf("device %d: max |G - G_host_sum| / max|G| = %.2e, tau_sq rel %.2e  %s\n", r, worst / scale, std::fabs(t - t0) / t0, ok ? "ok" : "FAIL");
    fail += !ok;
  }
  rdb_group_destroy(g);
  std::printf("group_check: %s (%lld samples over %d GPUs)\n", fail ? "FAILED" : "ok", (long long)n, R);
  return fail ? 1 : 0;
}
