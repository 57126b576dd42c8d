// MOCK: boost::shared_ptr as ROS message pointers use it
#pragma once
#include <memory>
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }
