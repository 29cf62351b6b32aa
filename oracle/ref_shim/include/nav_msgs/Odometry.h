#pragma once
#include <geometry_msgs/msgs.h>
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header;
  struct { geometry_msgs::Pose pose; } pose;
  struct { geometry_msgs::Twist twist; } twist;
  typedef std::shared_ptr<const Odometry> ConstPtr;
};
}  // namespace nav_msgs
