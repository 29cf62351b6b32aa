#pragma once
#include <ros/ros.h>
namespace sensor_msgs {
struct JointState {
  std_msgs::Header header;
  std::vector<std::string> name;
  std::vector<double> position, velocity, effort;
  typedef std::shared_ptr<const JointState> ConstPtr;
};
}  // namespace sensor_msgs
