#pragma once
#include <geometry_msgs/msgs.h>
namespace sensor_msgs {
struct Imu {
  std_msgs::Header header;
  geometry_msgs::Quaternion orientation;
  geometry_msgs::Vector3 angular_velocity, linear_acceleration;
  typedef std::shared_ptr<const Imu> ConstPtr;
};
}  // namespace sensor_msgs
