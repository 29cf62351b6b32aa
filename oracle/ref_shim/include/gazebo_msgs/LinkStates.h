#pragma once
#include <geometry_msgs/msgs.h>
namespace gazebo_msgs {
struct LinkStates {
  std::vector<std::string> name; std::vector<geometry_msgs::Pose> pose; std::vector<geometry_msgs::Twist> twist;
  typedef std::shared_ptr<const LinkStates> ConstPtr;
};
}  // namespace gazebo_msgs
