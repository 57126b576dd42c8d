// MOCK of tf::createQuaternionMsgFromRollPitchYaw
#pragma once
#include <cmath>
#include <geometry_msgs/Quaternion.h>
namespace tf {
inline geometry_msgs::Quaternion createQuaternionMsgFromRollPitchYaw(double r, double p, double y) {
  double cr = std::cos(r / 2), sr = std::sin(r / 2), cp = std::cos(p / 2), sp = std::sin(p / 2), cy = std::cos(y / 2), sy = std::sin(y / 2);
  geometry_msgs::Quaternion q;
  q.w = cr * cp * cy + sr * sp * sy; q.x = sr * cp * cy - cr * sp * sy; q.y = cr * sp * cy + sr * cp * sy; q.z = cr * cp * sy - sr * sp * cy;
  return q;
}
}
