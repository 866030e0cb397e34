// TEST INFRASTRUCTURE: stand-in for urdfdom's <urdf/model.h> / <urdf_model/model.h> -- just the data members that
// rosdyn_core reads in Joint::fromUrdf / Link::fromUrdf (primitives_impl.h:50-149, 276-331) and urdf_parser.h:44-57.
// Models are built programmatically by oracle/ref_driver.cpp from the flat chain descriptor; there is no XML parser here.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace urdf
{
struct Vector3
{
  double x = 0, y = 0, z = 0;
};
struct Rotation
{
  double x = 0, y = 0, z = 0, w = 1;
  // urdfdom_headers Rotation::setFromRPY (published formula)
  void setFromRPY(double roll, double pitch, double yaw)
  {
    const double phi = roll / 2.0, the = pitch / 2.0, psi = yaw / 2.0;
    x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
    y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
    z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
    w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
    const double s = std::sqrt(x * x + y * y + z * z + w * w);
    if (s == 0.0)
    {
      x = y = z = 0.0;
      w = 1.0;
    }
    else
    {
      x /= s;
      y /= s;
      z /= s;
      w /= s;
    }
  }
};
struct Pose
{
  Vector3 position;
  Rotation rotation;
};
struct JointLimits
{
  double lower = 0, upper = 0, effort = 0, velocity = 0;
};
struct Inertial
{
  Pose origin;
  double mass = 0, ixx = 0, ixy = 0, ixz = 0, iyy = 0, iyz = 0, izz = 0;
};
class Link;
class Joint
{
public:
  enum
  {
    UNKNOWN,
    REVOLUTE,
    CONTINUOUS,
    PRISMATIC,
    FLOATING,
    PLANAR,
    FIXED
  } type = UNKNOWN;
  std::string name, child_link_name, parent_link_name;
  Vector3 axis;
  Pose parent_to_joint_origin_transform;
  std::shared_ptr<JointLimits> limits;
};
class Link
{
public:
  std::string name;
  std::shared_ptr<Inertial> inertial;
  std::shared_ptr<Joint> parent_joint;
  std::vector<std::shared_ptr<Joint>> child_joints;
  std::vector<std::shared_ptr<Link>> child_links;
};
class ModelInterface
{
public:
  std::shared_ptr<Link> root_link_;
  std::string name_;
  std::map<std::string, std::shared_ptr<Link>> links_;
  std::map<std::string, std::shared_ptr<Joint>> joints_;
};
class Model : public ModelInterface
{
public:
  bool initParam(const std::string&) { return false; }   // no parameter server in the checker
  bool initString(const std::string&) { return false; }  // no XML parser in the checker
};
}  // namespace urdf
