#pragma once
#include <geometry_msgs/msgs.h>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, ADD = 0 };
  std_msgs::Header header;
  std::string ns;
  int id = 0, type = 0, action = 0;
  geometry_msgs::Pose pose;
  geometry_msgs::Vector3 scale;
  struct { float r = 0, g = 0, b = 0, a = 0; } color;
  ros::Duration lifetime;
  std::vector<geometry_msgs::Point> points;
};
}  // namespace visualization_msgs
