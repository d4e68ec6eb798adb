// STAND-IN for cnr_param (CNR-STIIMA-IRAS/cnr_param, absent from this image): a tiny in-process
// key/value store for scalar parameters (everything else is "not available", so the reference's
// get_param helpers fall back to their defaults).
// Test infrastructure for oracle/_ref only.
#pragma once
// the real cnr_param / yaml-cpp / boost headers pull these in transitively; the reference relies on that
#include <algorithm>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <numeric>
#include <random>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include <yaml-cpp/yaml.h>
namespace cnr
{
namespace param
{
enum class ModulesID { ROS1, ROS2, MAPPED_FILE };
inline bool is_mapped_file_avaliable() { return true; }
inline bool is_ros1_avaliable() { return false; }
inline bool is_ros2_avaliable() { return false; }
// the stand-in parameter server: scalar parameters only, set by oracle/ref_driver.cpp
inline std::map<std::string, double>& shim_store()
{
  static std::map<std::string, double> s;
  return s;
}
inline void shim_set(const std::string& key, double v) { shim_store()[key] = v; }
inline bool has(const std::string& key, std::string& what)
{
  if (shim_store().count(key))
    return true;
  what = "cnr_param stand-in: no such parameter";
  return false;
}
template <typename T>
inline bool get(const std::string& key, T& value, std::string& what, const bool& = true,
                const std::vector<ModulesID>& = std::vector<ModulesID>())
{
  if constexpr (std::is_arithmetic<T>::value)
  {
    auto it = shim_store().find(key);
    if (it != shim_store().end())
    {
      value = (T)it->second;
      return true;
    }
  }
  what = "cnr_param stand-in: no such parameter (or not a scalar)";
  return false;
}
}  // namespace param
}  // namespace cnr
