// TEST INFRASTRUCTURE: stand-in for <ros/ros.h>: an in-process parameter store (enough for ComponentBase, base_component.h:84-119),
// a steady clock for ros::Time / ros::Duration, and the logging macros.
#pragma once
#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "console.h"

namespace ros
{
struct ParamStore
{
  std::map<std::string, double> doubles;
  std::map<std::string, bool> bools;
  std::map<std::string, std::vector<std::string>> string_lists;
  std::map<std::string, std::map<std::string, double>> double_maps;
  static ParamStore& instance()
  {
    static ParamStore s;
    return s;
  }
};
namespace param
{
inline bool has(const std::string& k)
{
  ParamStore& s = ParamStore::instance();
  if (s.doubles.count(k) || s.bools.count(k) || s.string_lists.count(k) || s.double_maps.count(k)) return true;
  const std::string prefix = k + "/";  // a namespace exists when any key lives under it
  for (auto& kv : s.doubles)
    if (kv.first.compare(0, prefix.size(), prefix) == 0) return true;
  for (auto& kv : s.bools)
    if (kv.first.compare(0, prefix.size(), prefix) == 0) return true;
  return false;
}
inline bool get(const std::string& k, double& v)
{
  auto& m = ParamStore::instance().doubles;
  auto it = m.find(k);
  if (it == m.end()) return false;
  v = it->second;
  return true;
}
inline bool get(const std::string& k, bool& v)
{
  auto& m = ParamStore::instance().bools;
  auto it = m.find(k);
  if (it == m.end()) return false;
  v = it->second;
  return true;
}
inline void set(const std::string& k, double v) { ParamStore::instance().doubles[k] = v; }
inline void set(const std::string& k, bool v) { ParamStore::instance().bools[k] = v; }
}  // namespace param

class NodeHandle
{
public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  bool getParam(const std::string& k, std::map<std::string, double>& v) const
  {
    auto& m = ParamStore::instance().double_maps;
    auto it = m.find(k);
    if (it == m.end()) return false;
    v = it->second;
    return true;
  }
  bool getParam(const std::string& k, std::vector<std::string>& v) const
  {
    auto& m = ParamStore::instance().string_lists;
    auto it = m.find(k);
    if (it == m.end()) return false;
    v = it->second;
    return true;
  }
  bool getParam(const std::string& k, double& v) const { return param::get(k, v); }
  void setParam(const std::string& k, const std::map<std::string, double>& v) const { ParamStore::instance().double_maps[k] = v; }
  void setParam(const std::string& k, const std::vector<std::string>& v) const { ParamStore::instance().string_lists[k] = v; }
  void setParam(const std::string& k, double v) const
  {
    ParamStore::instance().doubles[k] = v;
    // "<ns>/<map>/<key>" also updates the map parameter, as the parameter server would
    const size_t p = k.rfind('/');
    if (p != std::string::npos)
    {
      auto& m = ParamStore::instance().double_maps;
      auto it = m.find(k.substr(0, p));
      if (it != m.end()) it->second[k.substr(p + 1)] = v;
    }
  }
};

class Duration
{
public:
  double s = 0;
  Duration() {}
  explicit Duration(double sec) : s(sec) {}
  double toSec() const { return s; }
  bool operator<(const Duration& o) const { return s < o.s; }
  bool operator>(const Duration& o) const { return s > o.s; }
};
class Time
{
public:
  double s = 0;
  // >= 0: deterministic clock for the checker, one second per call (turns the wall-clock budget of Chain::computeLocalIk,
  // primitives_impl.h:1405, into an iteration budget); < 0: the steady clock
  static long& ticks()
  {
    static thread_local long t = -1;
    return t;
  }
  static Time now()
  {
    Time t;
    long& k = ticks();
    if (k >= 0) t.s = (double)(k++);
    else t.s = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return t;
  }
  Duration operator-(const Time& o) const { return Duration(s - o.s); }
};
}  // namespace ros
