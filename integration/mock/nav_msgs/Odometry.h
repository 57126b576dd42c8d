// MOCK of nav_msgs/Odometry
#pragma once
#include <string>
#include <std_msgs/Header.h>
#include <geometry_msgs/Quaternion.h>
namespace nav_msgs {
struct Odometry {
  std_msgs::Header header; std::string child_frame_id;
  struct { struct { geometry_msgs::Point position; geometry_msgs::Quaternion orientation; } pose; double covariance[36]; } pose;
  struct { double covariance[36]; } twist;
};
}
