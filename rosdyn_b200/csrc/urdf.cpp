// urdf.cpp -- URDF string -> chain descriptor, host only, no third-party parser.
//
// Replaces, for the chain between two links, the model build the reference delegates to urdfdom + Link::fromUrdf /
// Joint::fromUrdf / Chain::init (primitives_impl.h:50-149, 276-331, 580-703, 1518-1527; urdf_parser.h:44-57):
//   * a minimal XML reader for the URDF subset that matters (robot / link / inertial / origin / mass / inertia /
//     joint / parent / child / axis / limit); everything else (visual, collision, transmission, gazebo ...) is skipped;
//   * rpy -> rotation through the URDF quaternion exactly like urdfdom's setFromRPY followed by Eigen's
//     Quaterniond -> matrix (urdf_parser.h:44-50);
//   * joint types: revolute and continuous -> REVOLUTE, prismatic -> PRISMATIC, anything else -> FIXED (primitives_impl.h:74-83);
//     axis defaults to (1,0,0) (URDF), normalised later by rdb_chain_create;
//   * limits with the reference's malformed-URDF defaults (primitives_impl.h:85-143);
//   * the chain is found by climbing parent joints from the tool link to the base link (primitives_impl.h:615-626):
//     RDB_ERR_NOT_FOUND with "Base link not found" / "Tool link not found" (primitives_impl.h:601-613).
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/rosdyn_b200.h"

namespace rdb
{
rdb_status set_error(rdb_status s, const std::string& what);  // capi.cu

namespace
{
struct XmlNode
{
  std::string name;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const XmlNode* child(const char* n) const
  {
    for (const auto& k : kids)
      if (k->name == n) return k.get();
    return nullptr;
  }
  const char* get(const char* a) const
  {
    auto it = attr.find(a);
    return it == attr.end() ? nullptr : it->second.c_str();
  }
};

class XmlReader
{
public:
  explicit XmlReader(const char* s) : p_(s) {}
  std::unique_ptr<XmlNode> parse(std::string& err)
  {
    std::unique_ptr<XmlNode> root;
    while (skip_misc())
    {
      if (*p_ != '<')
      {
        err = "text outside the root element";
        return nullptr;
      }
      root = element(err);
      if (!root) return nullptr;
      break;
    }
    if (!root) err = "no root element";
    return root;
  }

private:
  const char* p_;
  void ws()
  {
    while (*p_ && std::isspace((unsigned char)*p_)) p_++;
  }
  // skips whitespace, comments, processing instructions and DOCTYPE; false at end of input
  bool skip_misc()
  {
    for (;;)
    {
      ws();
      if (!*p_) return false;
      if (!std::strncmp(p_, "<!--", 4))
      {
        const char* e = std::strstr(p_ + 4, "-->");
        p_ = e ? e + 3 : p_ + std::strlen(p_);
      }
      else if (!std::strncmp(p_, "<?", 2))
      {
        const char* e = std::strstr(p_ + 2, "?>");
        p_ = e ? e + 2 : p_ + std::strlen(p_);
      }
      else if (!std::strncmp(p_, "<!", 2))
      {
        const char* e = std::strchr(p_, '>');
        p_ = e ? e + 1 : p_ + std::strlen(p_);
      }
      else
        return true;
    }
  }
  static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
  std::string name()
  {
    const char* b = p_;
    while (name_char(*p_)) p_++;
    return std::string(b, p_);
  }
  // nesting is bounded: the reader recurses once per level and a hostile document must not overflow the stack (a URDF nests 5 deep)
  static constexpr int MAX_DEPTH = 256;
  int depth_ = 0;
  struct DepthGuard
  {
    int& d;
    explicit DepthGuard(int& x) : d(x) { d++; }
    ~DepthGuard() { d--; }
  };
  std::unique_ptr<XmlNode> element(std::string& err)
  {
    DepthGuard guard(depth_);
    if (depth_ > MAX_DEPTH)
    {
      err = "elements nested deeper than 256 levels";
      return nullptr;
    }
    p_++;  // '<'
    std::unique_ptr<XmlNode> n(new XmlNode);
    n->name = name();
    if (n->name.empty())
    {
      err = "malformed tag";
      return nullptr;
    }
    for (;;)
    {
      ws();
      if (!*p_)
      {
        err = "unterminated tag <" + n->name + ">";
        return nullptr;
      }
      if (*p_ == '/' && p_[1] == '>')
      {
        p_ += 2;
        return n;
      }
      if (*p_ == '>')
      {
        p_++;
        break;
      }
      const std::string key = name();
      ws();
      if (key.empty() || *p_ != '=')
      {
        err = "malformed attribute in <" + n->name + ">";
        return nullptr;
      }
      p_++;
      ws();
      const char q = *p_;
      if (q != '"' && q != '\'')
      {
        err = "unquoted attribute in <" + n->name + ">";
        return nullptr;
      }
      const char* e = std::strchr(p_ + 1, q);
      if (!e)
      {
        err = "unterminated attribute in <" + n->name + ">";
        return nullptr;
      }
      n->attr[key] = std::string(p_ + 1, e);
      p_ = e + 1;
    }
    // content
    for (;;)
    {
      while (*p_ && *p_ != '<') p_++;  // character data is irrelevant for URDF
      if (!*p_)
      {
        err = "missing </" + n->name + ">";
        return nullptr;
      }
      if (!std::strncmp(p_, "<![CDATA[", 9))
      {
        const char* e = std::strstr(p_, "]]>");
        p_ = e ? e + 3 : p_ + std::strlen(p_);
        continue;
      }
      if (p_[1] == '!' || p_[1] == '?')
      {
        if (!skip_misc())
        {
          err = "missing </" + n->name + ">";
          return nullptr;
        }
        continue;
      }
      if (p_[1] == '/')
      {
        p_ += 2;
        const std::string close = name();
        ws();
        if (close != n->name || *p_ != '>')
        {
          err = "mismatched </" + close + "> for <" + n->name + ">";
          return nullptr;
        }
        p_++;
        return n;
      }
      std::unique_ptr<XmlNode> k = element(err);
      if (!k) return nullptr;
      n->kids.push_back(std::move(k));
    }
  }
};

bool vec3(const char* s, double v[3])
{
  if (!s) return false;
  char* e = nullptr;
  for (int k = 0; k < 3; k++)
  {
    v[k] = std::strtod(s, &e);
    if (e == s) return false;
    s = e;
  }
  return true;
}
double num(const char* s, double dflt)
{
  if (!s) return dflt;
  char* e = nullptr;
  const double v = std::strtod(s, &e);
  return e == s ? dflt : v;
}

// urdfdom Rotation::setFromRPY + Eigen::Quaterniond::toRotationMatrix, row-major
void rpy_to_rot(const double rpy[3], double R[9])
{
  const double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
  double x = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
  double y = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
  double z = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
  double w = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
  const double s = std::sqrt(x * x + y * y + z * z + w * w);
  x /= s; y /= s; z /= s; w /= s;
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
void origin(const XmlNode* o, double xyz[3], double R[9])
{
  double rpy[3] = {0, 0, 0};
  xyz[0] = xyz[1] = xyz[2] = 0;
  if (o)
  {
    double t[3];
    if (vec3(o->get("xyz"), t)) std::memcpy(xyz, t, sizeof(t));
    if (vec3(o->get("rpy"), t)) std::memcpy(rpy, t, sizeof(t));
  }
  rpy_to_rot(rpy, R);
}

struct UJoint
{
  std::string name, parent, child;
  rdb_joint_desc d;
  double q_max, q_min, dq_max, ddq_max, tau_max;
};
}  // namespace

struct UrdfChain;
struct UrdfHolder  // standard layout: a pointer to `pub` is a pointer to the holder
{
  rdb_urdf_chain pub;
  UrdfChain* impl;
};
struct UrdfChain
{
  UrdfHolder holder;
  std::vector<rdb_joint_desc> joints;
  std::vector<rdb_link_desc> links;
  std::vector<std::string> jn, ln;
  std::vector<const char*> jnp, lnp;
  std::vector<double> q_max, q_min, dq_max, ddq_max, tau_max;
};

}  // namespace rdb

using namespace rdb;

extern "C" {

rdb_status rdb_urdf_parse(const char* urdf_xml, const char* base_link, const char* tool_link, const double gravity[3], rdb_urdf_chain** out)
{
  if (!urdf_xml || !base_link || !tool_link || !out) return set_error(RDB_ERR_INVALID_ARG, "rdb_urdf_parse: null argument");
  *out = nullptr;
  std::string err;
  XmlReader rd(urdf_xml);
  std::unique_ptr<XmlNode> root = rd.parse(err);
  if (!root) return set_error(RDB_ERR_INVALID_ARG, "URDF: " + err);
  if (root->name != "robot") return set_error(RDB_ERR_INVALID_ARG, "URDF: root element is <" + root->name + ">, expected <robot>");

  std::map<std::string, rdb_link_desc> links;
  std::map<std::string, UJoint> joint_of_child;  // child link -> its parent joint
  for (const auto& k : root->kids)
  {
    if (k->name == "link")
    {
      const char* nm = k->get("name");
      if (!nm) return set_error(RDB_ERR_INVALID_ARG, "URDF: <link> without a name");
      rdb_link_desc L{};
      L.inertial_rot[0] = L.inertial_rot[4] = L.inertial_rot[8] = 1.0;
      if (const XmlNode* in = k->child("inertial"))  // Link::fromUrdf, primitives_impl.h:291-319
      {
        origin(in->child("origin"), L.cog, L.inertial_rot);
        if (const XmlNode* m = in->child("mass")) L.mass = num(m->get("value"), 0.0);
        if (const XmlNode* I = in->child("inertia"))
        {
          static const char* key[6] = {"ixx", "ixy", "ixz", "iyy", "iyz", "izz"};
          for (int a = 0; a < 6; a++) L.inertia[a] = num(I->get(key[a]), 0.0);
        }
      }
      links[nm] = L;
    }
    else if (k->name == "joint")
    {
      const char* nm = k->get("name");
      const XmlNode *pa = k->child("parent"), *chd = k->child("child");
      if (!nm || !pa || !chd || !pa->get("link") || !chd->get("link")) return set_error(RDB_ERR_INVALID_ARG, "URDF: malformed <joint>");
      UJoint J;
      J.name = nm;
      J.parent = pa->get("link");
      J.child = chd->get("link");
      std::memset(&J.d, 0, sizeof(J.d));
      const std::string type = k->get("type") ? k->get("type") : "";
      const bool revolute = type == "revolute", continuous = type == "continuous", prismatic = type == "prismatic";
      J.d.type = (revolute || continuous) ? RDB_JOINT_REVOLUTE : (prismatic ? RDB_JOINT_PRISMATIC : RDB_JOINT_FIXED);  // primitives_impl.h:74-83
      J.d.input_index = -1;
      origin(k->child("origin"), J.d.xyz, J.d.rot);
      J.d.axis[0] = 1.0;  // URDF default axis
      if (const XmlNode* ax = k->child("axis"))
      {
        double t[3];
        if (vec3(ax->get("xyz"), t)) std::memcpy(J.d.axis, t, sizeof(t));
      }
      // limits, primitives_impl.h:85-143.  (m_Dq_max is left uninitialised by the reference when <limit> is missing; 0 here.)
      J.q_max = J.q_min = J.dq_max = J.ddq_max = J.tau_max = 0.0;
      const XmlNode* lim = k->child("limit");
      if (revolute || prismatic)
      {
        if (!lim)
        {
          J.q_max = 1e10; J.q_min = -1e10; J.ddq_max = 10.0 * J.dq_max; J.tau_max = 1e10;
        }
        else
        {
          J.q_max = num(lim->get("upper"), 0.0);
          J.q_min = num(lim->get("lower"), 0.0);
          if (J.q_max <= J.q_min)
          {
            J.q_max = 2 * M_PI;
            J.q_min = -2 * M_PI;
          }
          J.dq_max = num(lim->get("velocity"), 0.0);
          if (J.dq_max <= 0.0) J.dq_max = 2 * M_PI;
          J.ddq_max = 10.0 * J.dq_max;
          J.tau_max = num(lim->get("effort"), 0.0);
        }
      }
      else if (continuous)
      {
        J.q_max = 1e10; J.q_min = -1e10;
        if (!lim)
        {
          J.ddq_max = 10.0 * J.dq_max; J.tau_max = 1e10;
        }
        else
        {
          J.dq_max = num(lim->get("velocity"), 0.0);
          J.ddq_max = 10.0 * J.dq_max;
          J.tau_max = num(lim->get("effort"), 0.0);
        }
      }
      joint_of_child[J.child] = J;
    }
  }
  // Chain::init: base must exist, tool must be a descendant of base (findChild from base, primitives_impl.h:600-613)
  if (!links.count(base_link)) return set_error(RDB_ERR_NOT_FOUND, "Base link not found");
  if (!links.count(tool_link)) return set_error(RDB_ERR_NOT_FOUND, "Tool link not found");
  std::vector<UJoint> chain;  // tool -> base while climbing
  std::string act = tool_link;
  size_t guard = 0;
  while (act != base_link)
  {
    auto it = joint_of_child.find(act);
    if (it == joint_of_child.end() || ++guard > joint_of_child.size() + 1) return set_error(RDB_ERR_NOT_FOUND, "Tool link not found");
    chain.push_back(it->second);
    act = it->second.parent;
    if (!links.count(act)) return set_error(RDB_ERR_INVALID_ARG, "URDF: joint '" + it->second.name + "' has an unknown parent link");
  }
  if (chain.size() > RDB_MAX_JOINTS) return set_error(RDB_ERR_INVALID_ARG, "URDF: chain longer than RDB_MAX_JOINTS");

  std::unique_ptr<UrdfChain> u(new UrdfChain);
  const int nj = (int)chain.size();
  u->links.push_back(links[base_link]);
  u->ln.push_back(base_link);
  int n_in = 0;
  for (int k = nj - 1; k >= 0; k--)  // base -> tool
  {
    UJoint& J = chain[k];
    if (J.d.type != RDB_JOINT_FIXED) J.d.input_index = n_in++;  // default inputs: the moveable joints base -> tool (primitives_impl.h:631-636, 700)
    u->joints.push_back(J.d);
    u->jn.push_back(J.name);
    u->links.push_back(links[J.child]);
    u->ln.push_back(J.child);
    u->q_max.push_back(J.q_max);
    u->q_min.push_back(J.q_min);
    u->dq_max.push_back(J.dq_max);
    u->ddq_max.push_back(J.ddq_max);
    u->tau_max.push_back(J.tau_max);
  }
  for (const auto& s : u->jn) u->jnp.push_back(s.c_str());
  for (const auto& s : u->ln) u->lnp.push_back(s.c_str());
  rdb_urdf_chain& p = u->holder.pub;
  u->holder.impl = u.get();
  p.desc.n_joints = nj;
  p.desc.n_inputs = n_in;
  for (int k = 0; k < 3; k++) p.desc.gravity[k] = gravity ? gravity[k] : 0.0;  // ctor default gravity is zero (primitives.h:346)
  p.desc.joints = u->joints.data();
  p.desc.links = u->links.data();
  p.joint_names = u->jnp.data();
  p.link_names = u->lnp.data();
  p.q_max = u->q_max.data();
  p.q_min = u->q_min.data();
  p.dq_max = u->dq_max.data();
  p.ddq_max = u->ddq_max.data();
  p.tau_max = u->tau_max.data();
  *out = &u.release()->holder.pub;
  return RDB_OK;
}

void rdb_urdf_chain_free(rdb_urdf_chain* c)
{
  if (c) delete reinterpret_cast<UrdfHolder*>(c)->impl;
}

rdb_status rdb_chain_from_urdf(const char* urdf_xml, const char* base_link, const char* tool_link, const double gravity[3], rdb_chain** out)
{
  if (!out) return set_error(RDB_ERR_INVALID_ARG, "rdb_chain_from_urdf: null output");
  *out = nullptr;
  rdb_urdf_chain* u = nullptr;
  rdb_status s = rdb_urdf_parse(urdf_xml, base_link, tool_link, gravity, &u);
  if (s != RDB_OK) return s;
  s = rdb_chain_create(&u->desc, out);
  rdb_urdf_chain_free(u);
  return s;
}

}  // extern "C"
