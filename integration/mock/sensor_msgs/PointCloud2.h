// MOCK of sensor_msgs/PointCloud2 (fields as in SURVEY.md Appendix C)
#pragma once
#include <cstdint>
#include <vector>
#include <string>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 7; uint32_t count = 1; };
struct PointCloud2 {
  typedef boost::shared_ptr<PointCloud2 const> ConstPtr;
  std_msgs::Header header; uint32_t height = 1, width = 0; std::vector<PointField> fields; bool is_bigendian = false;
  uint32_t point_step = 16, row_step = 0; std::vector<uint8_t> data; bool is_dense = true;
};
}
