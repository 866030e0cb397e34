// Compile test of the Eigen-typed overloads of include/rosdyn_b200/chain.hpp (the reference's signatures, primitives.h:452-548) against the
// Eigen subset of oracle/shim -- Eigen3 itself is not installed in the build image.  tests/test_cpp_headers.py compiles and links this file;
// it runs only where a CUDA device exists (the engine has no CPU fallback) and then checks the Eigen results against the std-container ones.
#include <cmath>
#include <cstdio>

#include "rosdyn_b200/chain.hpp"

#ifndef ROSDYN_B200_HAVE_EIGEN
#error "compile with -I oracle/shim (or a real Eigen3) so that <Eigen/Core> exists"
#endif

using rosdyn_b200::Chain;

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

int main()
{
  if (rdb_device_count() <= 0)
  {
    std::printf("no CUDA device: compiled and linked only\n");
    return 0;
  }
  std::vector<rdb_joint_desc> J = {joint(RDB_JOINT_REVOLUTE, 0, 0, 0.1, 0, 0, 0, 0, 0, 1, 0), joint(RDB_JOINT_REVOLUTE, 0, 0.2, 0, 0, 1.57, 0, 0, 1, 0, 1),
                                   joint(RDB_JOINT_PRISMATIC, 0.3, 0, 0, 0.1, 0, 0.2, 1, 0, 0, 2)};
  std::vector<rdb_link_desc> L(4);
  for (int l = 1; l < 4; l++)
  {
    L[l].mass = 1.0 + l;
    L[l].cog[2] = 0.1 * l;
    L[l].inertial_rot[0] = L[l].inertial_rot[4] = L[l].inertial_rot[8] = 1.0;
    L[l].inertia[0] = L[l].inertia[3] = 0.02 * l;
    L[l].inertia[5] = 0.01 * l;
  }
  rdb_chain_desc desc{3, 3, {0, 0, -9.806}, J.data(), L.data()};
  Chain ch(desc);
  Eigen::VectorXd q(3), dq(3), ddq(3), dddq(3);
  q << 0.3, -0.7, 0.2;
  dq << 0.5, 0.1, -0.4;
  ddq << -0.2, 0.9, 0.3;
  dddq << 0.1, 0.2, 0.3;
  int fail = 0;
  const Eigen::Affine3d T = ch.getTransformation(q);
  const Chain::EVectorOfAffine3d Ts = ch.getTransformations(q);
  const Chain::EMatrix6Xd Jac = ch.getJacobian(q);
  const Chain::EVectorOfVector6d tw = ch.getTwist(q, dq), dtw = ch.getDTwist(q, dq, ddq), ddtw = ch.getDDTwist(q, dq, ddq, dddq);
  const Eigen::VectorXd tau = ch.getJointTorque(q, dq, ddq), taun = ch.getJointTorqueNonLinearPart(q, dq);
  const Eigen::MatrixXd Phi = ch.getRegressor(q, dq, ddq), M = ch.getJointInertia(q);
  const Eigen::VectorXd pi = ch.getNominalParametersEigen();
  // invariants of SURVEY.md section 4, on the Eigen objects
  const Eigen::VectorXd t2 = Phi * pi, t3 = M * ddq + taun;
  const Chain::EVector6d vt = Jac * dq;
  for (int k = 0; k < 3; k++)
  {
    if (std::fabs(t2(k) - tau(k)) > 1e-10 * (1 + std::fabs(tau(k)))) fail++;
    if (std::fabs(t3(k) - tau(k)) > 1e-10 * (1 + std::fabs(tau(k)))) fail++;
  }
  for (int k = 0; k < 6; k++)
    if (std::fabs(vt(k) - tw.back()(k)) > 1e-12) fail++;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
      if (T.matrix()(r, c) != Ts.back().matrix()(r, c)) fail++;
  (void)dtw;
  (void)ddtw;
  // batched sibling: samples as columns, Eigen-record results
  Eigen::MatrixXd Q(3, 5), DQ(3, 5), DDQ(3, 5);
  for (int i = 0; i < 5; i++)
    for (int k = 0; k < 3; k++)
    {
      Q(k, i) = q(k) + 0.1 * i;
      DQ(k, i) = dq(k);
      DDQ(k, i) = ddq(k);
    }
  std::vector<double> rec, trec;
  ch.getRegressorBatch(Q, DQ, DDQ, rec, &trec);
  for (int c = 0; c < 30; c++)
    for (int r = 0; r < 3; r++)
      if (rec[(size_t)c * 3 + r] != Phi(r, c)) fail++;  // sample 0 == the per-sample call
  std::printf(fail ? "FAIL %d\n" : "eigen facade ok\n", fail);
  return fail ? 1 : 0;
}
