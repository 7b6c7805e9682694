// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for the ROS symbols that the
// reference hot-path headers mention (state_6dof.h ROS_ERROR; parameters.h NodeHandle/Duration).
#ifndef ORACLE_SHIM_ROS_H
#define ORACLE_SHIM_ROS_H
#include <array>
#include <cstdio>
#include <map>
#include <memory>
#include <vector>
#include <string>
#define ROS_ERROR(...) (std::fprintf(stderr, __VA_ARGS__), std::fputc('\n', stderr))
#define ROS_WARN(...) (std::fprintf(stderr, __VA_ARGS__), std::fputc('\n', stderr))
#define ROS_INFO(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
namespace ros
{
class NodeHandle
{
};
class Duration
{
public:
  Duration() : sec_(0) {}
  explicit Duration(double s) : sec_(s) {}
  double toSec() const { return sec_; }
private:
  double sec_;
};
}  // namespace ros
#endif
