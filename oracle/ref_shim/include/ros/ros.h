// Stub of <ros/ros.h> -- just enough surface for the UNMODIFIED reference node sources to compile without ROS.
// TEST INFRASTRUCTURE ONLY (oracle/ref_shim): lets tests pin the oracle restatement against the reference's own
// translation units.  Nothing here is derived from ROS sources; every call is a no-op except NodeHandle::param,
// which reads a process-global table that the shim driver fills before constructing the node.
#pragma once
#include <math.h>
#include <stdio.h>

#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace ref_shim {
inline std::map<std::string, double> &param_table() {
  static std::map<std::string, double> t;
  return t;
}
}  // namespace ref_shim

namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
template <class... A>
int bind(A &&...) { return 0; }
}  // namespace boost
struct ref_shim_placeholder {};
static ref_shim_placeholder _1;

namespace ros {
struct Time {
  double sec_ = 0.0;
  Time() {}
  Time(double s) : sec_(s) {}
  static Time now() { return Time(0.0); }
  double toSec() const { return sec_; }
};
struct Duration {
  Duration() {}
  Duration(double) {}
};
struct Rate {
  Rate(double) {}
  void sleep() {}
};
inline bool ok() { return false; }
inline void spinOnce() {}
inline void init(int &, char **, const std::string &) {}

struct Publisher {
  template <class M>
  void publish(const M &) const {}
};
struct Subscriber {};

struct NodeHandle {
  NodeHandle() {}
  NodeHandle(const std::string &) {}
  template <class M>
  Publisher advertise(const std::string &, int) { return Publisher(); }
  template <class M = void, class... A>
  Subscriber subscribe(const std::string &, int, A &&...) { return Subscriber(); }
  // nh_.param(name, var, default): launch-file overrides come from ref_shim::param_table()
  template <class T>
  void param(const std::string &name, T &var, const T &def) const {
    auto it = ref_shim::param_table().find(name);
    var = it == ref_shim::param_table().end() ? def : (T)it->second;
  }
  void param(const std::string &, std::string &var, const std::string &def) const { var = def; }
};
}  // namespace ros

#define ROS_ERROR(...) do { fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_INFO(...) do { } while (0)

namespace std_msgs {
struct Header {
  unsigned seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs
