#pragma once
#include <typeinfo>
#include "joint_command_interface.h"
namespace hardware_interface {
class RobotHW {
 public:
  virtual ~RobotHW() = default;
  template <class T> void registerInterface(T* iface) { ifaces_[typeid(T).name()] = iface; }
  template <class T> T* get() { auto it = ifaces_.find(typeid(T).name()); return it == ifaces_.end() ? nullptr : static_cast<T*>(it->second); }
  virtual void doSwitch(const std::list<ControllerInfo>&, const std::list<ControllerInfo>&) {}
 private:
  std::map<std::string, void*> ifaces_;
};
}  // namespace hardware_interface
