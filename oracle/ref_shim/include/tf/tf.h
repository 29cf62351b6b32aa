// Stub of <tf/tf.h>: the few tf types the reference nodes name.  Yaw <-> quaternion helpers are ordinary
// textbook formulas (input preparation only; the solve itself never touches tf).
#pragma once
#include <geometry_msgs/msgs.h>

#include <stdexcept>
namespace tf {
struct TransformException : public std::runtime_error {
  TransformException(const std::string &m) : std::runtime_error(m) {}
};
struct Vector3 {
  double v[3];
  Vector3() : v{0, 0, 0} {}
  Vector3(double x, double y, double z) : v{x, y, z} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
struct Quaternion {
  double x_, y_, z_, w_;
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternion(double x, double y, double z, double w) : x_(x), y_(y), z_(z), w_(w) {}
};
struct Matrix3x3 {
  double m[3][3];
  Matrix3x3() : m{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}} {}
  explicit Matrix3x3(const Quaternion &q) {
    double x = q.x_, y = q.y_, z = q.z_, w = q.w_;
    m[0][0] = 1 - 2 * (y * y + z * z); m[0][1] = 2 * (x * y - z * w); m[0][2] = 2 * (x * z + y * w);
    m[1][0] = 2 * (x * y + z * w); m[1][1] = 1 - 2 * (x * x + z * z); m[1][2] = 2 * (y * z - x * w);
    m[2][0] = 2 * (x * z - y * w); m[2][1] = 2 * (y * z + x * w); m[2][2] = 1 - 2 * (x * x + y * y);
  }
  void getRPY(double &roll, double &pitch, double &yaw) const {
    pitch = asin(-m[2][0]);
    roll = atan2(m[2][1], m[2][2]);
    yaw = atan2(m[1][0], m[0][0]);
  }
  Vector3 operator*(const Vector3 &a) const {
    return Vector3(m[0][0] * a.x() + m[0][1] * a.y() + m[0][2] * a.z(), m[1][0] * a.x() + m[1][1] * a.y() + m[1][2] * a.z(),
                   m[2][0] * a.x() + m[2][1] * a.y() + m[2][2] * a.z());
  }
};
struct StampedTransform {
  Matrix3x3 basis;
  geometry_msgs::TransformStamped msg;
  const Matrix3x3 &getBasis() const { return basis; }
};
inline void transformStampedTFToMsg(const StampedTransform &t, geometry_msgs::TransformStamped &m) { m = t.msg; }
inline double getYaw(const geometry_msgs::Quaternion &q) {
  return atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
inline geometry_msgs::Quaternion createQuaternionMsgFromYaw(double yaw) {
  geometry_msgs::Quaternion q;
  q.x = 0; q.y = 0; q.z = sin(yaw / 2); q.w = cos(yaw / 2);
  return q;
}
}  // namespace tf
