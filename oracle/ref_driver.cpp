// ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Batch driver around the REFERENCE's own rosdyn::Chain (the headers under /root/reference/rosdyn_core/include are compiled where
// they lie; see oracle/Makefile target _ref).  Eigen3 / urdfdom / roscpp are not installed in this image, so the third-party headers
// the reference includes are the small stand-ins of oracle/shim/ (dense arithmetic restated as plain loops, urdf structs without the
// XML parser, logging dropped).  Every kinematic / dynamic formula that runs here is the reference's source text:
//   Joint::fromUrdf, Link::fromUrdf, Chain::init, setInputJointsName, computeFrames, computeScrews, getTransformations, getJacobian,
//   getTwist, getDTwist*, getDDTwist*, getWrench, getJointTorque, getRegressor, getJointInertia, getNominalParameters
//   (internal/primitives_impl.h) and spacevect_algebra.h.
// The exported symbols mirror oracle/rosdyn_oracle.c one to one (same names, same SoA layouts), so oracle/oracle.py drives either
// library; tests/test_reference_build.py checks the restatement against this build.  Only tests/ and bench.py's CPU legs load it.
#include <rosdyn_core/primitives.h>

#ifdef _OPENMP
#include <omp.h>
#endif
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/rosdyn_b200.h" /* descriptor structs only */
#include "../include/rosdyn_b200/rosdyn_core_bridge.h" /* the reference-side binding under test: rosdyn::Chain -> rdb_chain_desc (no GPU) */

namespace
{
struct RefChain
{
  std::shared_ptr<urdf::Model> model;
  rosdyn::ChainPtr chain;
  std::vector<std::string> input_names;
  std::vector<std::shared_ptr<urdf::Joint>> ujoints;  // by chain joint (limits are re-read by createChain)
  std::vector<int> input_of_joint;
  Eigen::Vector3d gravity;
  std::string base_name, tool_name;
  std::vector<rosdyn::ChainPtr> per_thread;  // rosdyn::Chain is stateful and not reentrant: one clone() per worker (primitives.h:554)
  int nJ = 0, nL = 0, n_in = 0;
  // chains for `nthreads` workers (0 = all); worker 0 uses the original
  int workers(int nthreads)
  {
    int nt = 1;
#ifdef _OPENMP
    nt = nthreads <= 0 ? omp_get_max_threads() : nthreads;
#endif
    (void)nthreads;
    while ((int)per_thread.size() < nt)
    {
      if (per_thread.empty())
        per_thread.push_back(chain);
      else
      {
        rosdyn::ChainPtr c = chain->clone();
        c->setInputJointsName(input_names);
        per_thread.push_back(c);
      }
    }
    return nt;
  }
};
int tid()
{
#ifdef _OPENMP
  return omp_get_thread_num();
#else
  return 0;
#endif
}

// rotation matrix (row-major) -> unit quaternion, Shepperd's method; the reference turns it back into a matrix (urdf_parser.h:44-50)
urdf::Rotation rot_to_quat(const double* R)
{
  urdf::Rotation q;
  const double m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
  const double tr = m00 + m11 + m22;
  if (tr > 0)
  {
    const double s = std::sqrt(tr + 1.0) * 2;
    q.w = 0.25 * s;
    q.x = (m21 - m12) / s;
    q.y = (m02 - m20) / s;
    q.z = (m10 - m01) / s;
  }
  else if (m00 > m11 && m00 > m22)
  {
    const double s = std::sqrt(1.0 + m00 - m11 - m22) * 2;
    q.w = (m21 - m12) / s;
    q.x = 0.25 * s;
    q.y = (m01 + m10) / s;
    q.z = (m02 + m20) / s;
  }
  else if (m11 > m22)
  {
    const double s = std::sqrt(1.0 + m11 - m00 - m22) * 2;
    q.w = (m02 - m20) / s;
    q.x = (m01 + m10) / s;
    q.y = 0.25 * s;
    q.z = (m12 + m21) / s;
  }
  else
  {
    const double s = std::sqrt(1.0 + m22 - m00 - m11) * 2;
    q.w = (m10 - m01) / s;
    q.x = (m02 + m20) / s;
    q.y = (m12 + m21) / s;
    q.z = 0.25 * s;
  }
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.x /= n;
  q.y /= n;
  q.z /= n;
  q.w /= n;
  return q;
}

Eigen::VectorXd gather(const double* x, int n_in, int64_t ld, int64_t i)
{
  Eigen::VectorXd v(n_in);
  for (int j = 0; j < n_in; j++) v(j) = x ? x[(int64_t)j * ld + i] : 0.0;
  return v;
}

void put_vec6(double* dst, int64_t ld, int64_t i, const rosdyn::VectorOfVector6d& v)
{
  if (!dst) return;
  for (size_t l = 0; l < v.size(); l++)
    for (int k = 0; k < 6; k++) dst[(int64_t)(6 * l + k) * ld + i] = v[l](k);
}
}  // namespace

extern "C" {

const char* oracle_kind(void) { return "reference"; }
int oracle_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void* oracle_chain_create(const rdb_chain_desc* d)
{
  try
  {
    if (!d || d->n_joints < 0) return nullptr;
    auto rc = std::make_unique<RefChain>();
    rc->nJ = d->n_joints;
    rc->nL = d->n_joints + 1;
    rc->n_in = d->n_inputs;
    rc->model = std::make_shared<urdf::Model>();
    std::vector<std::shared_ptr<urdf::Link>> links(rc->nL);
    for (int l = 0; l < rc->nL; l++)
    {
      links[l] = std::make_shared<urdf::Link>();
      links[l]->name = "link_" + std::to_string(l);
      const rdb_link_desc& L = d->links[l];
      links[l]->inertial = std::make_shared<urdf::Inertial>();
      urdf::Inertial& in = *links[l]->inertial;
      in.mass = L.mass;
      in.origin.position.x = L.cog[0];
      in.origin.position.y = L.cog[1];
      in.origin.position.z = L.cog[2];
      in.origin.rotation = rot_to_quat(L.inertial_rot);
      in.ixx = L.inertia[0];
      in.ixy = L.inertia[1];
      in.ixz = L.inertia[2];
      in.iyy = L.inertia[3];
      in.iyz = L.inertia[4];
      in.izz = L.inertia[5];
    }
    std::vector<std::string> input_names(rc->n_in);
    for (int k = 0; k < rc->n_in; k++) input_names[k] = "__unlisted_input_" + std::to_string(k);  // no such joint: column of S stays zero
    for (int j = 0; j < rc->nJ; j++)
    {
      const rdb_joint_desc& J = d->joints[j];
      auto uj = std::make_shared<urdf::Joint>();
      uj->name = "joint_" + std::to_string(j);
      uj->type = J.type == RDB_JOINT_REVOLUTE ? urdf::Joint::REVOLUTE : (J.type == RDB_JOINT_PRISMATIC ? urdf::Joint::PRISMATIC : urdf::Joint::FIXED);
      uj->axis.x = J.axis[0];
      uj->axis.y = J.axis[1];
      uj->axis.z = J.axis[2];
      uj->parent_to_joint_origin_transform.position.x = J.xyz[0];
      uj->parent_to_joint_origin_transform.position.y = J.xyz[1];
      uj->parent_to_joint_origin_transform.position.z = J.xyz[2];
      uj->parent_to_joint_origin_transform.rotation = rot_to_quat(J.rot);
      uj->limits = std::make_shared<urdf::JointLimits>();
      uj->limits->lower = -1.0e9;
      uj->limits->upper = 1.0e9;
      uj->limits->velocity = 10.0;
      uj->limits->effort = 100.0;
      uj->parent_link_name = links[j]->name;
      uj->child_link_name = links[j + 1]->name;
      links[j]->child_joints.push_back(uj);
      links[j]->child_links.push_back(links[j + 1]);
      links[j + 1]->parent_joint = uj;
      if (J.input_index >= 0 && J.input_index < rc->n_in) input_names[J.input_index] = uj->name;
      rc->ujoints.push_back(uj);
      rc->input_of_joint.push_back(J.input_index);
    }
    rc->model->root_link_ = links[0];
    Eigen::Vector3d g;
    g << d->gravity[0], d->gravity[1], d->gravity[2];
    rc->gravity = g;
    rc->base_name = links.front()->name;
    rc->tool_name = links.back()->name;
    rc->chain = rosdyn::createChain(*rc->model, links.front()->name, links.back()->name, g);
    if (!rc->chain) return nullptr;
    rc->chain->setInputJointsName(input_names);
    rc->input_names = input_names;
    return rc.release();
  }
  catch (const std::exception&)
  {
    return nullptr;
  }
}

void oracle_chain_destroy(void* c) { delete static_cast<RefChain*>(c); }
int oracle_chain_joints_number(const void* c) { return (int)static_cast<const RefChain*>(c)->chain->getJointsNumber(); }
int oracle_chain_inputs_number(const void* c) { return (int)static_cast<const RefChain*>(c)->chain->getActiveJointsNumber(); }

void oracle_nominal_parameters(const void* c, double* out)
{
  const Eigen::VectorXd p = static_cast<const RefChain*>(c)->chain->getNominalParameters();
  for (Eigen::Index k = 0; k < p.size(); k++) out[k] = p(k);
}

void oracle_kinematics_batch(const void* cv, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq, const double* dddq,
                             int64_t ld_out, double* T_tool, double* T_links, double* jacobian, double* twist, double* dtwist,
                             double* dtwist_lin, double* dtwist_nonlin, double* ddtwist, double* ddtwist_lin, double* ddtwist_nonlin,
                             double* torque, int nthreads)
{
  RefChain* rc = static_cast<RefChain*>(const_cast<void*>(cv));
  const int nt = rc->workers(nthreads);
  const int n_in = rc->n_in, nL = rc->nL;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    rosdyn::Chain& ch = *rc->per_thread[tid()];
    const Eigen::VectorXd vq = gather(q, n_in, ld, i), vdq = gather(dq, n_in, ld, i), vddq = gather(ddq, n_in, ld, i),
                          vdddq = gather(dddq, n_in, ld, i);
    if (T_tool || T_links)
    {
      const rosdyn::VectorOfAffine3d& T = ch.getTransformations(vq);
      for (int l = 0; l < nL; l++)
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 4; c++)
          {
            const double v = T[l].matrix()(r, c);
            if (T_links) T_links[(int64_t)(12 * l + 4 * r + c) * ld_out + i] = v;
            if (T_tool && l == nL - 1) T_tool[(int64_t)(4 * r + c) * ld_out + i] = v;
          }
    }
    if (jacobian)
    {
      const Eigen::Matrix6Xd& J = ch.getJacobian(vq);
      for (int c = 0; c < n_in; c++)
        for (int r = 0; r < 6; r++) jacobian[(int64_t)(6 * c + r) * ld_out + i] = J(r, c);
    }
    // order matters: the full recursions first (their direct paths), the linear / non-linear parts afterwards, so that the
    // reference's "sum of the cached parts" shortcuts (primitives_impl.h:1108-1112, 1205-1209) are not taken
    if (twist) put_vec6(twist, ld_out, i, ch.getTwist(vq, vdq));
    if (dtwist) put_vec6(dtwist, ld_out, i, ch.getDTwist(vq, vdq, vddq));
    if (ddtwist) put_vec6(ddtwist, ld_out, i, ch.getDDTwist(vq, vdq, vddq, vdddq));
    if (torque)
    {
      const Eigen::VectorXd& t = ch.getJointTorque(vq, vdq, vddq);
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = t(k);
    }
    if (dtwist_lin) put_vec6(dtwist_lin, ld_out, i, ch.getDTwistLinearPart(vq, vddq));
    if (dtwist_nonlin) put_vec6(dtwist_nonlin, ld_out, i, ch.getDTwistNonLinearPart(vq, vdq));
    if (ddtwist_lin) put_vec6(ddtwist_lin, ld_out, i, ch.getDDTwistLinearPart(vq, vdddq));
    if (ddtwist_nonlin) put_vec6(ddtwist_nonlin, ld_out, i, ch.getDDTwistNonLinearPart(vq, vdq, vddq));
  }
}

void oracle_regressor_torque_batch(const void* cv, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq, int64_t ld_out,
                                   double* phi, double* torque, int nthreads)
{
  RefChain* rc = static_cast<RefChain*>(const_cast<void*>(cv));
  const int nt = rc->workers(nthreads);
  const int n_in = rc->n_in, P = 10 * rc->nJ;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    rosdyn::Chain& ch = *rc->per_thread[tid()];
    const Eigen::VectorXd vq = gather(q, n_in, ld, i), vdq = gather(dq, n_in, ld, i), vddq = gather(ddq, n_in, ld, i);
    const Eigen::VectorXd t = ch.getJointTorque(vq, vdq, vddq);
    const Eigen::MatrixXd R = ch.getRegressor(vq, vdq, vddq);
    if (phi)
      for (int c = 0; c < P; c++)
        for (int r = 0; r < n_in; r++) phi[(int64_t)(c * n_in + r) * ld_out + i] = R(r, c);
    if (torque)
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = t(k);
  }
}

void oracle_inertia_batch(const void* cv, int64_t n, int64_t ld, const double* q, int64_t ld_out, double* inertia, int nthreads)
{
  RefChain* rc = static_cast<RefChain*>(const_cast<void*>(cv));
  const int nt = rc->workers(nthreads);
  const int n_in = rc->n_in;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (int64_t i = 0; i < n; i++)
  {
    rosdyn::Chain& ch = *rc->per_thread[tid()];
    const Eigen::MatrixXd& M = ch.getJointInertia(gather(q, n_in, ld, i));
    for (int c = 0; c < n_in; c++)
      for (int r = 0; r < n_in; r++) inertia[(int64_t)(c * n_in + r) * ld_out + i] = M(r, c);
  }
}

// Chain::getWrench + getJointTorque(q,Dq,DDq,ext) and Chain::getJacobianLink, through the reference's own methods
void oracle_wrench_batch(const void* cv, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq, const double* ext,
                         int64_t ld_ext, int64_t ld_out, double* torque, double* wrenches)
{
  const RefChain* rc = static_cast<const RefChain*>(cv);
  rosdyn::Chain& ch = *rc->chain;
  const int n_in = rc->n_in, nL = rc->nL;
  for (int64_t i = 0; i < n; i++)
  {
    const Eigen::VectorXd vq = gather(q, n_in, ld, i), vdq = gather(dq, n_in, ld, i), vddq = gather(ddq, n_in, ld, i);
    rosdyn::VectorOfVector6d e(nL);
    for (int l = 0; l < nL; l++)
      for (int k = 0; k < 6; k++) e[l](k) = ext ? ext[(int64_t)(6 * l + k) * ld_ext + i] : 0.0;
    const Eigen::VectorXd t = ch.getJointTorque(vq, vdq, vddq, e);
    if (torque)
      for (int k = 0; k < n_in; k++) torque[(int64_t)k * ld_out + i] = t(k);
    if (wrenches) put_vec6(wrenches, ld_out, i, ch.getWrench(vq, vdq, vddq, e));
  }
}
void oracle_jacobian_link_batch(const void* cv, int64_t n, int64_t ld, const double* q, int link, int64_t ld_out, double* jac)
{
  const RefChain* rc = static_cast<const RefChain*>(cv);
  rosdyn::Chain& ch = *rc->chain;
  const int n_in = rc->n_in;
  const std::string name = ch.getLinksName().at(link);
  for (int64_t i = 0; i < n; i++)
  {
    const Eigen::VectorXd vq = gather(q, n_in, ld, i);
    ch.getTransformation(vq);  // getJacobianLink itself ignores q (maybe_unused(q), primitives_impl.h:953): bring the frames up to date first
    const Eigen::Matrix6Xd J = ch.getJacobianLink(vq, name);
    for (int c = 0; c < n_in; c++)
      for (int r = 0; r < 6; r++) jac[(int64_t)(6 * c + r) * ld_out + i] = c < J.cols() ? J(r, c) : 0.0;
  }
}

// normal equations of the reference's regressor / torque with long-double accumulation
// Chain::computeLocalIk / computeWeigthedLocalIk of the reference (primitives_impl.h:1398-1468) for n targets.  The joint limits go in
// through the urdf model (Joint::fromUrdf reads them, Chain::setInputJointsName copies them to m_q_min / m_q_max); the wall clock is the
// shim's tick clock, so that `max_iter + 1` loop iterations run (the restatement allows max_iter steps and max_iter + 1 checks; the
// reference checks at the top of each iteration only); solve_quadprog is the stand-in of shim/eigen_matrix_utils.
void oracle_local_ik_batch(void* cv, int64_t n, int64_t ld, const double* target, const double* seed, const double* q_min, const double* q_max,
                           const double* weight, double toll, int max_iter, double* sol_out, int32_t* status, int32_t* iters, double* err_norm)
{
  RefChain* rc = static_cast<RefChain*>(cv);
  try
  {
  for (size_t j = 0; j < rc->ujoints.size(); j++)
  {
    const int r = rc->input_of_joint[j];
    if (r < 0 || r >= rc->n_in) continue;
    rc->ujoints[j]->limits->lower = q_min ? q_min[r] : -1e10;
    rc->ujoints[j]->limits->upper = q_max ? q_max[r] : 1e10;
  }
  rosdyn::ChainPtr ch = rosdyn::createChain(*rc->model, rc->base_name, rc->tool_name, rc->gravity);
  ch->setInputJointsName(rc->input_names);
  const int n_in = rc->n_in;
  for (int64_t i = 0; i < n; i++)
  {
    Eigen::Affine3d T;
    for (int r = 0; r < 3; r++)
      for (int k = 0; k < 4; k++) T.matrix()(r, k) = target[(int64_t)(4 * r + k) * ld + i];
    Eigen::VectorXd sd(n_in), sol(n_in);
    for (int r = 0; r < n_in; r++) sd(r) = seed[(int64_t)r * ld + i];
    ros::Time::ticks() = 0;
    bool ok;
    if (weight)
    {
      Eigen::Vector6d w;
      for (int k = 0; k < 6; k++) w(k) = weight[k];
      ok = ch->computeWeigthedLocalIk(sol, T, w, sd, toll, ros::Duration(max_iter + 1.5));
    }
    else
      ok = ch->computeLocalIk(sol, T, sd, toll, ros::Duration(max_iter + 1.5));
    const long used = ros::Time::ticks();  // 1 (tini) + one per loop test
    ros::Time::ticks() = -1;
    for (int r = 0; r < n_in; r++) sol_out[(int64_t)r * ld + i] = sol(r);
    if (status) status[i] = ok ? 1 : 0;
    if (iters) iters[i] = (int32_t)(used - 2);  // steps taken before the successful check
    if (err_norm) err_norm[i] = 0.0;
  }
  }
  catch (const std::exception& e)
  {
    fprintf(stderr, "oracle_local_ik_batch (reference build): %s\n", e.what());
    if (status)
      for (int64_t i = 0; i < n; i++) status[i] = -1;
  }
}

// Chain::getMultiplicity of the reference (primitives_impl.h:1470-1517); limits through the urdf model as for the IK
int64_t oracle_multiplicity(void* cv, const double* q, const double* q_min, const double* q_max, double* out, int64_t capacity)
{
  RefChain* rc = static_cast<RefChain*>(cv);
  for (size_t j = 0; j < rc->ujoints.size(); j++)
  {
    const int r = rc->input_of_joint[j];
    if (r < 0 || r >= rc->n_in) continue;
    rc->ujoints[j]->limits->lower = q_min[r];
    rc->ujoints[j]->limits->upper = q_max[r];
  }
  rosdyn::ChainPtr ch = rosdyn::createChain(*rc->model, rc->base_name, rc->tool_name, rc->gravity);
  ch->setInputJointsName(rc->input_names);
  Eigen::VectorXd v(rc->n_in);
  for (int r = 0; r < rc->n_in; r++) v(r) = q[r];
  const std::vector<Eigen::VectorXd> all = ch->getMultiplicity(v);
  for (size_t k = 0; k < all.size() && (int64_t)k < capacity; k++)
    for (int r = 0; r < rc->n_in; r++) out[k * rc->n_in + r] = all[k](r);
  return (int64_t)all.size();
}

void oracle_regressor_gram(const void* cv, int64_t n, int64_t ld, const double* q, const double* dq, const double* ddq, const double* tau_meas,
                           double* gram, double* rhs, double* tau_sq)
{
  const RefChain* rc = static_cast<const RefChain*>(cv);
  rosdyn::Chain& ch = *rc->chain;
  const int n_in = rc->n_in, P = 10 * rc->nJ;
  std::vector<long double> G((size_t)P * P, 0.0L), b(P, 0.0L);
  long double tt = 0.0L;
  for (int64_t i = 0; i < n; i++)
  {
    const Eigen::VectorXd vq = gather(q, n_in, ld, i), vdq = gather(dq, n_in, ld, i), vddq = gather(ddq, n_in, ld, i);
    Eigen::VectorXd t = ch.getJointTorque(vq, vdq, vddq);
    if (tau_meas) t = gather(tau_meas, n_in, ld, i);
    const Eigen::MatrixXd R = ch.getRegressor(vq, vdq, vddq);
    for (int r = 0; r < n_in; r++)
    {
      tt += (long double)t(r) * t(r);
      for (int c = 0; c < P; c++)
      {
        const long double x = R(r, c);
        if (x == 0.0L) continue;
        b[c] += x * t(r);
        for (int c2 = 0; c2 < P; c2++) G[(size_t)c * P + c2] += x * R(r, c2);
      }
    }
  }
  for (int k = 0; k < P * P; k++) gram[k] = (double)G[k];
  for (int k = 0; k < P; k++) rhs[k] = (double)b[k];
  if (tau_sq) *tau_sq = (double)tt;
}

// The reference's component classes themselves (friction_polynomial1.h, friction_polynomial2.h, ideal_spring.h), constructed through the
// stand-in parameter store exactly as rosdyn_identification would configure them; nominal coefficients are arbitrary (the regressor
// does not depend on them).
void oracle_components_regressor_batch(int n_comp, const rdb_component_desc* comp, int n_in, int64_t n, int64_t ld, const double* q,
                                       const double* dq, int64_t ld_out, double* phi_c)
{
  ros::NodeHandle nh;
  const std::string robot = "checker_robot";
  std::vector<std::string> names(n_in);
  for (int k = 0; k < n_in; k++) names[k] = "in_" + std::to_string(k);
  nh.setParam(robot + "/joint_names", names);
  int col = 0;
  for (int k = 0; k < n_comp; k++)
  {
    const rdb_component_desc& c = comp[k];
    const std::string jn = names.at(c.input_index);
    rosdyn::ComponentPtr cp;
    if (c.type == RDB_COMPONENT_IDEAL_SPRING)
    {
      nh.setParam(robot + "/" + jn + "/spring/coefficients", std::map<std::string, double>{{"elasticity", 3.0}, {"offset_effort", 0.5}});
      nh.setParam(robot + "/" + jn + "/spring/constants", std::map<std::string, double>{});
      cp.reset(new rosdyn::IdealSpring(jn, robot, nh));
    }
    else
    {
      nh.setParam(robot + "/" + jn + "/friction/constants", std::map<std::string, double>{{"min_velocity", c.min_velocity}, {"max_velocity", c.max_velocity}});
      if (c.type == RDB_COMPONENT_FRICTION_POLY1)
      {
        nh.setParam(robot + "/" + jn + "/friction/coefficients", std::map<std::string, double>{{"coloumb", 1.0}, {"viscous", 2.0}});
        cp.reset(new rosdyn::FirstOrderPolynomialFriction(jn, robot, nh));
      }
      else
      {
        nh.setParam(robot + "/" + jn + "/friction/coefficients",
                    std::map<std::string, double>{{"coloumb", 1.0}, {"first_order_viscous", 2.0}, {"second_order_viscous", 0.3}});
        cp.reset(new rosdyn::SecondOrderPolynomialFriction(jn, robot, nh));
      }
    }
    const int nc = (int)cp->getParametersNumber();
    for (int64_t i = 0; i < n; i++)
    {
      Eigen::VectorXd vq = gather(q, n_in, ld, i), vdq = gather(dq, n_in, ld, i), vddq = Eigen::VectorXd::Zero(n_in);
      const Eigen::MatrixXd R = cp->getRegressor(vq, vdq, vddq);
      for (int p = 0; p < nc; p++)
        for (int r = 0; r < n_in; r++) phi_c[((int64_t)(col + p) * n_in + r) * ld_out + i] = R(r, p);
    }
    col += nc;
  }
}

// The reference-side binding (include/rosdyn_b200/rosdyn_core_bridge.h) run on the reference's own Chain: joints[nJ], links[nJ + 1],
// gravity[3] as rosdyn::toB200Desc reports them; returns the number of inputs.  tests/test_cpp_headers.py feeds the result back into the
// restatement and compares it with the reference's outputs on the original chain.
int oracle_bridge_descriptor(void* cv, rdb_joint_desc* joints, rdb_link_desc* links, double gravity[3])
{
  RefChain& rc = *static_cast<RefChain*>(cv);
  std::vector<rdb_joint_desc> J;
  std::vector<rdb_link_desc> L;
  const rdb_chain_desc d = rosdyn::toB200Desc(*rc.chain, J, L);
  for (size_t k = 0; k < J.size(); k++) joints[k] = J[k];
  for (size_t k = 0; k < L.size(); k++) links[k] = L[k];
  for (int k = 0; k < 3; k++) gravity[k] = d.gravity[k];
  return d.n_inputs;
}

}  // extern "C"
