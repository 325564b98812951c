#pragma once
#include "joint_state_interface.h"
namespace hardware_interface {
class JointHandle : public JointStateHandle {
 public:
  JointHandle() = default;
  JointHandle(const JointStateHandle& js, double* cmd) : JointStateHandle(js), cmd_(cmd) {}
  void setCommand(double c) { *cmd_ = c; }
  double getCommand() const { return *cmd_; }
 private:
  double* cmd_ = nullptr;
};
class JointCommandInterface : public ResourceManager<JointHandle> {};
class EffortJointInterface : public JointCommandInterface {};
class VelocityJointInterface : public JointCommandInterface {};
class PositionJointInterface : public JointCommandInterface {};
}  // namespace hardware_interface
