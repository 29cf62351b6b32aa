#pragma once
#include <geometry_msgs/msgs.h>
namespace nav_msgs {
struct Path {
  std_msgs::Header header;
  std::vector<geometry_msgs::PoseStamped> poses;
  typedef std::shared_ptr<const Path> ConstPtr;
};
}  // namespace nav_msgs
