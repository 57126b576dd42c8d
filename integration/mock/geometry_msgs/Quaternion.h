#pragma once
namespace geometry_msgs { struct Quaternion { double x = 0, y = 0, z = 0, w = 1; }; struct Point { double x = 0, y = 0, z = 0; }; }
