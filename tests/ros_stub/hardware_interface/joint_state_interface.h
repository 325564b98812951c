// Minimal stand-in for ros_control's hardware_interface (handles and name-keyed interfaces), enough for
// MjHWInterface (reference src/mujoco_sim/mj_hw_interface.cpp) to compile and run without ROS.
#pragma once
#include <ros/ros.h>  // as the real ros_control headers do

#include <list>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
namespace hardware_interface {
class JointStateHandle {
 public:
  JointStateHandle() = default;
  JointStateHandle(const std::string& name, const double* pos, const double* vel, const double* eff) : name_(name), pos_(pos), vel_(vel), eff_(eff) {}
  const std::string& getName() const { return name_; }
  double getPosition() const { return *pos_; }
  double getVelocity() const { return *vel_; }
  double getEffort() const { return *eff_; }
 private:
  std::string name_;
  const double *pos_ = nullptr, *vel_ = nullptr, *eff_ = nullptr;
};
template <class Handle>
class ResourceManager {
 public:
  void registerHandle(const Handle& h) { handles_[h.getName()] = h; }
  Handle getHandle(const std::string& name) {
    auto it = handles_.find(name);
    if (it == handles_.end()) throw std::runtime_error("no handle named " + name);
    return it->second;
  }
  std::vector<std::string> getNames() const {
    std::vector<std::string> n;
    for (auto& kv : handles_) n.push_back(kv.first);
    return n;
  }
 private:
  std::map<std::string, Handle> handles_;
};
class JointStateInterface : public ResourceManager<JointStateHandle> {};
struct InterfaceResources {
  std::string hardware_interface;
  std::set<std::string> resources;
};
struct ControllerInfo {
  std::string name, type;
  std::vector<InterfaceResources> claimed_resources;
};
}  // namespace hardware_interface
