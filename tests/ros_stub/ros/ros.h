// Minimal stand-in for <ros/ros.h>: just enough for the reference's hot-path sources to compile without ROS
// (tests/test_source_compat.py).  Not a ROS implementation.
#pragma once
#include <cstdio>
#include <string>
namespace ros {
class NodeHandle {
 public:
  NodeHandle() = default;
  explicit NodeHandle(const std::string&) {}
};
}  // namespace ros
#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_ERROR(...) do { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
