// MOCK of loam/OptStatus (Header header; float32[36] hessian -- degerate_odometry_filter.cpp:30-31)
#pragma once
#include <std_msgs/Header.h>
namespace loam { struct OptStatus { std_msgs::Header header; float hessian[36]; }; }
