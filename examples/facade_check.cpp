// facade_check.cpp -- self-checking example of the C++ rosdyn::Chain-style facade (include/rosdyn_b200/chain.hpp).
// Mirrors the shape of the reference's smoke loops (rosdyn_core/test/test.cpp:108-187: every getter on random
// inputs) but, unlike them, asserts: the algebraic invariants of SURVEY.md section 4 and the UR10 zero-pose answer.
// Build: tools/build_facade.py (g++ only, links librosdyn_b200.so).  Needs a CUDA device to run.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rosdyn_b200/chain.hpp"

using rosdyn_b200::Chain;
using rosdyn_b200::VectorXd;

static int g_fail = 0;
#define CHECK(cond, ...)                  \
  do                                      \
  {                                       \
    if (!(cond))                          \
    {                                     \
      std::printf("FAIL %s: ", #cond);    \
      std::printf(__VA_ARGS__);           \
      std::printf("\n");                  \
      g_fail++;                           \
    }                                     \
  } while (0)

static rdb_joint_desc joint(int type, double x, double y, double z, double r, double p, double yw, double ax, double ay, double az, int in)
{
  rdb_joint_desc j{};
  j.type = type;
  j.input_index = in;
  j.xyz[0] = x; j.xyz[1] = y; j.xyz[2] = z;
  rosdyn_b200::rpyToRot(r, p, yw, j.rot);
  j.axis[0] = ax; j.axis[1] = ay; j.axis[2] = az;
  return j;
}
static rdb_link_desc link(double m, double cz, double r, double l)
{
  rdb_link_desc k{};
  k.mass = m;
  k.cog[2] = cz;
  k.inertial_rot[0] = k.inertial_rot[4] = k.inertial_rot[8] = 1.0;
  k.inertia[0] = k.inertia[3] = m * (3 * r * r + l * l) / 12.0;
  k.inertia[5] = m * r * r / 2.0;
  return k;
}

int main()
{
  if (rdb_device_count() <= 0)
  {
    std::printf("no CUDA device: rosdyn_b200 has no CPU fallback\n");
    return 2;
  }
  const double hp = M_PI / 2;
  // UR10-like base_link -> tool0, gravity of rosdyn_speed_test.cpp:61-62
  std::vector<rdb_joint_desc> J = {
      joint(RDB_JOINT_REVOLUTE, 0, 0, 0.1273, 0, 0, 0, 0, 0, 1, 0),       joint(RDB_JOINT_REVOLUTE, 0, 0.220941, 0, 0, hp, 0, 0, 1, 0, 1),
      joint(RDB_JOINT_REVOLUTE, 0, -0.1719, 0.612, 0, 0, 0, 0, 1, 0, 2),  joint(RDB_JOINT_REVOLUTE, 0, 0, 0.5723, 0, hp, 0, 0, 1, 0, 3),
      joint(RDB_JOINT_REVOLUTE, 0, 0.1149, 0, 0, 0, 0, 0, 0, 1, 4),       joint(RDB_JOINT_REVOLUTE, 0, 0, 0.1157, 0, 0, 0, 0, 1, 0, 5),
      joint(RDB_JOINT_FIXED, 0, 0.0922, 0, -hp, 0, 0, 0, 0, 0, -1)};
  std::vector<rdb_link_desc> L = {link(4, 0, 0.075, 0.038),       link(7.778, 0, 0.075, 0.178), link(12.93, 0.306, 0.075, 0.612),
                                  link(3.87, 0.28615, 0.075, 0.5723), link(1.96, 0, 0.075, 0.12),   link(1.96, 0, 0.075, 0.12),
                                  link(0.202, 0, 0.075, 0.12),    link(0, 0, 0, 0)};
  rdb_chain_desc d{};
  d.n_joints = 7;
  d.n_inputs = 6;
  d.gravity[2] = -9.806;
  d.joints = J.data();
  d.links = L.data();
  Chain chain(d);
  CHECK(chain.getJointsNumber() == 7 && chain.getLinksNumber() == 8 && chain.getActiveJointsNumber() == 6, "sizes");

  // zero pose known answer
  const VectorXd z(6, 0.0);
  const auto T0 = chain.getTransformation(z);
  CHECK(std::fabs(T0[12] - 1.1843) < 1e-12 && std::fabs(T0[13] - 0.256141) < 1e-12 && std::fabs(T0[14] - 0.0116) < 1e-12, "tool0 at q=0: %g %g %g",
        T0[12], T0[13], T0[14]);

  const VectorXd pi = chain.getNominalParameters();
  std::srand(7);
  auto rnd = [] { return 2.0 * std::rand() / RAND_MAX - 1.0; };  // Eigen::setRandom range, test.cpp:113-116
  for (int trial = 0; trial < 20; trial++)
  {
    VectorXd q(6), Dq(6), DDq(6), DDDq(6);
    for (int k = 0; k < 6; k++) q[k] = rnd(), Dq[k] = rnd(), DDq[k] = rnd(), DDDq[k] = rnd();
    const VectorXd tau = chain.getJointTorque(q, Dq, DDq);
    const VectorXd phi = chain.getRegressor(q, Dq, DDq);  // 6 x 70 column-major
    const VectorXd M = chain.getJointInertia(q);
    const VectorXd h = chain.getJointTorqueNonLinearPart(q, Dq);
    const VectorXd Jac = chain.getJacobian(q);
    const auto v = chain.getTwistTool(q, Dq);
    const auto a = chain.getDTwist(q, Dq, DDq), al = chain.getDTwistLinearPart(q, DDq), an = chain.getDTwistNonLinearPart(q, Dq);
    const auto jk = chain.getDDTwist(q, Dq, DDq, DDDq), jl = chain.getDDTwistLinearPart(q, DDDq), jn = chain.getDDTwistNonLinearPart(q, Dq, DDq);
    for (int r = 0; r < 6; r++)
    {
      double t1 = 0, t2 = h[r], vv = 0;
      for (int c = 0; c < 70; c++) t1 += phi[c * 6 + r] * pi[c];
      for (int c = 0; c < 6; c++) t2 += M[c * 6 + r] * DDq[c], vv += Jac[c * 6 + r] * Dq[c];
      CHECK(std::fabs(t1 - tau[r]) < 1e-10 * (1 + std::fabs(tau[r])), "Phi pi == tau row %d: %g vs %g", r, t1, tau[r]);
      CHECK(std::fabs(t2 - tau[r]) < 1e-10 * (1 + std::fabs(tau[r])), "M ddq + h == tau row %d", r);
      CHECK(std::fabs(vv - v[r]) < 1e-10, "J dq == twist_tool row %d", r);
      for (int c = 0; c < r; c++) CHECK(M[c * 6 + r] == M[r * 6 + c], "M symmetric");
    }
    for (int l = 0; l < 8; l++)
      for (int k = 0; k < 6; k++)
      {
        CHECK(std::fabs(a[l][k] - al[l][k] - an[l][k]) < 1e-10, "dtwist == lin + nonlin");
        CHECK(std::fabs(jk[l][k] - jl[l][k] - jn[l][k]) < 1e-10, "ddtwist == lin + nonlin");
      }
  }
  // error behaviour of the reference: std::invalid_argument on size mismatch (primitives_impl.h:1299-1309)
  bool threw = false;
  try
  {
    chain.getRegressor(VectorXd(6, 0.0), VectorXd(5, 0.0), VectorXd(6, 0.0));
  }
  catch (const std::invalid_argument&)
  {
    threw = true;
  }
  CHECK(threw, "getRegressor must throw std::invalid_argument on dimension mismatch");

  // batched host sibling: normal equations of 10000 samples; G pi == b because tau comes from RNEA
  const int64_t n = 10000;
  std::vector<double> q(6 * n), dq(6 * n), ddq(6 * n), G(70 * 70), b(70);
  double tt = 0;
  for (int s = 0; s < 3; s++) rdb_fill_uniform_host(s == 0 ? q.data() : (s == 1 ? dq.data() : ddq.data()), 6, n, n, 0x5EED0001, s);
  rdb_samples in{n, n, q.data(), dq.data(), ddq.data(), nullptr};
  chain.getRegressorGramHost(in, nullptr, G.data(), b.data(), &tt, false);
  double worst = 0, scale = 0;
  for (int r = 0; r < 70; r++)
  {
    double acc = 0;
    for (int c = 0; c < 70; c++) acc += G[c * 70 + r] * pi[c];
    worst = std::fmax(worst, std::fabs(acc - b[r]));
    scale = std::fmax(scale, std::fabs(b[r]));
  }
  CHECK(worst <= 1e-9 * scale, "G pi_nom == b: %g (scale %g)", worst, scale);
  // latency of the per-sample getters (N = 1 batches: one launch + one synchronisation each), next to the reference's published per-call
  // figures on a CPU core (reference README.md:29-45): the facade is a drop-in for signatures, the speed-up exists only for batches
  {
    VectorXd q1(6, 0.3), d1(6, 0.1), dd1(6, -0.2);
    struct Row { const char* name; double ref_us; int which; };
    const Row rows[] = {{"getTransformation", 0.7597, 0}, {"getJacobian", 1.0656, 1}, {"getJointTorque", 3.7673, 2}, {"getRegressor", -1.0, 3},
                        {"getJointInertia", 10.0676, 4}};
    for (const Row& r : rows)
    {
      const int reps = 2000;
      for (int pass = 0; pass < 2; pass++)
      {
        const auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < reps; k++)
        {
          q1[0] = 0.3 + 1e-6 * k;
          if (r.which == 0) chain.getTransformation(q1);
          else if (r.which == 1) chain.getJacobian(q1);
          else if (r.which == 2) chain.getJointTorque(q1, d1, dd1);
          else if (r.which == 3) chain.getRegressor(q1, d1, dd1);
          else chain.getJointInertia(q1);
        }
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
        if (pass == 1)
        {
          if (r.ref_us > 0) std::printf("latency %-18s %8.2f us per call (N = 1)   reference (CPU core, README.md:29-45): %6.2f us\n", r.name, us, r.ref_us);
          else std::printf("latency %-18s %8.2f us per call (N = 1)   reference: not published\n", r.name, us);
        }
      }
    }
  }
  std::printf("facade_check: %s (kernels launched: %llu)\n", g_fail ? "FAILED" : "ok", (unsigned long long)rdb_kernel_launch_count());
  return g_fail ? 1 : 0;
}
