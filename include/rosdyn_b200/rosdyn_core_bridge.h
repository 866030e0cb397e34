// rosdyn_core_bridge.h -- the reference-side binding: a rosdyn::Chain of rosdyn_core (the reference, header-only C++ on Eigen) converted to
// the POD chain descriptor of the C-ABI (include/rosdyn_b200.h), so that code holding a rosdyn::ChainPtr can hand its model to the B200 engine.
//
// It is written ONLY against accessors the reference really has
//   Chain : getJoints(), getLinks(), getActiveJointsName(), getGravity()                 (primitives.h:385-409)
//   Joint : getName(), getType(), getTransformation(q), getScrew_of_child_in_parent()     (primitives.h:115-151)
//   Link  : getMass(), getCog(), getSpatialInertia()                                      (primitives.h:204-219)
// and only against element access / small fixed-size products of Eigen, so it compiles with Eigen3 proper and with the stand-in Eigen subset
// of oracle/shim (that is how this repo compile- and run-tests it: oracle/ref_driver.cpp::oracle_bridge_descriptor, tests/test_cpp_headers.py).
//
// Include AFTER <rosdyn_core/primitives.h>.  rosdyn::toB200Desc fills caller-owned arrays (no GPU needed); rosdyn::toB200 also creates the handle.
#ifndef ROSDYN_B200_ROSDYN_CORE_BRIDGE_H
#define ROSDYN_B200_ROSDYN_CORE_BRIDGE_H

#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "../rosdyn_b200.h"

namespace rosdyn
{

// One chain link as Link::fromUrdf left it (primitives_impl.h:288-319).  getSpatialInertia() is the 6x6 inertia about the link ORIGIN in link
// axes, [[m I, m c^T], [m c^, R I_cog R^T + m c^ c^T]] (spacevect_algebra.h:232-239): the descriptor wants the inertia about the cog, so the
// parallel-axis term is taken off again and the inertial-frame rotation is the identity (the rotation is already applied).
inline rdb_link_desc toB200Link(Link& l)
{
  rdb_link_desc d{};
  d.mass = l.getMass();
  const Eigen::Vector3d c = l.getCog();
  const Eigen::Matrix66d& S = l.getSpatialInertia();
  const double cs[3][3] = {{0.0, -c(2), c(1)}, {c(2), 0.0, -c(0)}, {-c(1), c(0), 0.0}};
  double Ic[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
    {
      double cc = 0.0;  // (c^ c^T)(i, j)
      for (int k = 0; k < 3; k++) cc += cs[i][k] * cs[j][k];
      Ic[i][j] = S(3 + i, 3 + j) - d.mass * cc;
    }
  for (int k = 0; k < 3; k++) d.cog[k] = c(k);
  d.inertial_rot[0] = d.inertial_rot[4] = d.inertial_rot[8] = 1.0;
  d.inertia[0] = Ic[0][0]; d.inertia[1] = Ic[0][1]; d.inertia[2] = Ic[0][2];
  d.inertia[3] = Ic[1][1]; d.inertia[4] = Ic[1][2]; d.inertia[5] = Ic[2][2];
  return d;
}

// One chain joint as Joint::fromUrdf left it (primitives_impl.h:50-83).  T_pj = getTransformation(0) (R_jc(0) = I, primitives_impl.h:40-46); the
// axis in the parent frame is the non-zero half of the screw (primitives_impl.h:25-35), the axis in the joint frame is R_pj^T of it.
inline rdb_joint_desc toB200Joint(Joint& j, const std::vector<std::string>& active_joints)
{
  rdb_joint_desc d{};
  d.type = j.getType() == Joint::REVOLUTE ? RDB_JOINT_REVOLUTE : (j.getType() == Joint::PRISMATIC ? RDB_JOINT_PRISMATIC : RDB_JOINT_FIXED);
  const Eigen::Affine3d T = j.getTransformation(0.0);
  const Eigen::Matrix3d R = T.linear();
  const Eigen::Vector3d t = T.translation();
  const Eigen::Vector6d s = j.getScrew_of_child_in_parent();
  const int o = d.type == RDB_JOINT_REVOLUTE ? 3 : 0;  // [0; axis] revolute, [axis; 0] prismatic, 0 fixed
  for (int r = 0; r < 3; r++)
  {
    d.xyz[r] = t(r);
    for (int c = 0; c < 3; c++) d.rot[3 * r + c] = R(r, c);
    d.axis[r] = 0.0;
    for (int k = 0; k < 3; k++) d.axis[r] += R(k, r) * s(o + k);  // R_pj^T axis_p
  }
  const auto it = std::find(active_joints.begin(), active_joints.end(), j.getName());
  d.input_index = it == active_joints.end() ? -1 : int(it - active_joints.begin());
  if (d.type == RDB_JOINT_FIXED) d.input_index = -1;
  return d;
}

// The chain base -> tool as a descriptor; `joints` / `links` own the storage the descriptor points into.
inline rdb_chain_desc toB200Desc(Chain& c, std::vector<rdb_joint_desc>& joints, std::vector<rdb_link_desc>& links)
{
  joints.clear();
  links.clear();
  const std::vector<std::string>& active = c.getActiveJointsName();
  for (const LinkPtr& l : c.getLinks()) links.push_back(toB200Link(*l));
  for (const JointPtr& j : c.getJoints()) joints.push_back(toB200Joint(*j, active));
  rdb_chain_desc desc{};
  desc.n_joints = int32_t(joints.size());
  desc.n_inputs = int32_t(active.size());
  const Eigen::Vector3d g = c.getGravity();
  for (int k = 0; k < 3; k++) desc.gravity[k] = g(k);
  desc.joints = joints.data();
  desc.links = links.data();
  return desc;
}

// rosdyn::Chain -> handle of the B200 engine (throws std::runtime_error like the reference's constructors, primitives_impl.h:498-501)
inline rdb_chain* toB200(Chain& c)
{
  std::vector<rdb_joint_desc> J;
  std::vector<rdb_link_desc> L;
  const rdb_chain_desc desc = toB200Desc(c, J, L);
  rdb_chain* h = nullptr;
  if (rdb_chain_create(&desc, &h) != RDB_OK) throw std::runtime_error(std::string("rosdyn_b200: ") + rdb_last_error());
  return h;
}

}  // namespace rosdyn

#endif  // ROSDYN_B200_ROSDYN_CORE_BRIDGE_H
