// Stub geometry_msgs (plain structs with the field names the reference sources touch).
#pragma once
#include <ros/ros.h>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseArray { std_msgs::Header header; std::vector<Pose> poses; };
struct Twist { Vector3 linear, angular; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::Header header; std::string child_frame_id; Transform transform; };
struct Wrench { Vector3 force, torque; };
struct WrenchStamped {
  std_msgs::Header header; Wrench wrench;
  typedef std::shared_ptr<const WrenchStamped> ConstPtr;
};
}  // namespace geometry_msgs
