// MOCK of the small part of roscpp that integration/ros/vlo_loam_node.cpp touches (compile check only).
#pragma once
#include <string>
#include <vector>
#include <iostream>
#include <sstream>
#include <memory>
#include <boost_shim.h>
namespace ros {
struct Time { double t = 0; double toSec() const { return t; } };
struct Publisher { template <class M> void publish(const M &) const {} };
struct Subscriber {};
class NodeHandle {
 public:
  explicit NodeHandle(const std::string & = "") {}
  template <class T> bool getParam(const std::string &, T &) const { return false; }
  template <class M> Publisher advertise(const std::string &, int) { return Publisher(); }
  template <class M> Subscriber subscribe(const std::string &, int, void (*)(const boost::shared_ptr<M const> &)) { return Subscriber(); }
};
inline void init(int &, char **, const std::string &) {}
inline void spin() {}
}  // namespace ros
#define ROS_WARN_STREAM(x) do { std::ostringstream ss_; ss_ << x; std::cerr << ss_.str() << std::endl; } while (0)
#define ROS_FATAL(...) do { std::cerr << "fatal" << std::endl; } while (0)
