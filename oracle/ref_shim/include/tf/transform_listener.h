#pragma once
#include <tf/tf.h>
namespace tf {
struct TransformListener {
  StampedTransform current;  // what lookupTransform hands out; the shim driver sets it
  void lookupTransform(const std::string &, const std::string &, const ros::Time &, StampedTransform &out) const { out = current; }
};
}  // namespace tf
