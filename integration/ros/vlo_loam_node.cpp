// Drop-in replacement for the four `loam` nodelets in front of gtsam_fusion.
// Needs ROS (roscpp, sensor_msgs, nav_msgs, tf) and the `loam` message package to build for real; in this repository
// it is compiled against the minimal mock headers under integration/mock/ (tests/test_abi_and_host.py) so that the
// calls into include/vlo.h stay in step with the C-ABI.
// vlo_loam_node.cpp  -- drop-in for `loam` (publishes what gtsam_fusion_filter subscribes to)
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <nav_msgs/Odometry.h>
#include <loam/OptStatus.h>          // Header header; float32[36] hessian
#include <tf/transform_datatypes.h>
#include <cmath>
#include "vlo.h"

static vlo_handle *H;
static ros::Publisher pubOdom, pubStatus, pubMapped, pubMapStatus;

// LOAM pose (rx ry rz tx ty tz, R = Ry Rx Rz, axes x left / y up / z forward) -> the ROS-convention frame the *_CORRECTED
// topics carry (loam.launch:18-23: /camera_init -> /camera_init_CORRECTED is the axis change ROS (x, y, z) = LOAM (z, x, y),
// the same permutation as vil_fusion/python/loam_frame_transform.py:52-90)
static void loamPoseToRos(const float p[6], double q[4], double t[3]) {
  const double sx = sin(p[0]), cx = cos(p[0]), sy = sin(p[1]), cy = cos(p[1]), sz = sin(p[2]), cz = cos(p[2]);
  const double L[3][3] = {{cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx}, {cx * sz, cx * cz, -sx},
                          {-sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx}};
  const int m[3] = {2, 0, 1};                       // ROS axis i is LOAM axis m[i]
  double R[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = L[m[i]][m[j]];
  for (int i = 0; i < 3; i++) t[i] = p[3 + m[i]];
  const double tr = R[0][0] + R[1][1] + R[2][2];    // rotation matrix -> quaternion (w x y z)
  if (tr > 0) { double s = 2 * sqrt(tr + 1); q[0] = s / 4; q[1] = (R[2][1] - R[1][2]) / s; q[2] = (R[0][2] - R[2][0]) / s; q[3] = (R[1][0] - R[0][1]) / s; }
  else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) { double s = 2 * sqrt(1 + R[0][0] - R[1][1] - R[2][2]); q[0] = (R[2][1] - R[1][2]) / s; q[1] = s / 4; q[2] = (R[0][1] + R[1][0]) / s; q[3] = (R[0][2] + R[2][0]) / s; }
  else if (R[1][1] > R[2][2]) { double s = 2 * sqrt(1 + R[1][1] - R[0][0] - R[2][2]); q[0] = (R[0][2] - R[2][0]) / s; q[1] = (R[0][1] + R[1][0]) / s; q[2] = s / 4; q[3] = (R[1][2] + R[2][1]) / s; }
  else { double s = 2 * sqrt(1 + R[2][2] - R[0][0] - R[1][1]); q[0] = (R[1][0] - R[0][1]) / s; q[1] = (R[0][2] + R[2][0]) / s; q[2] = (R[1][2] + R[2][1]) / s; q[3] = s / 4; }
}

static void publish(const std_msgs::Header &hdr, const float pose[6], const vlo_result &r,
                    ros::Publisher &po, ros::Publisher &ps) {
  nav_msgs::Odometry o; o.header = hdr; o.header.frame_id = "/camera_init_CORRECTED"; o.child_frame_id = "/laser_odom";
  double q[4], t[3]; loamPoseToRos(pose, q, t);
  o.pose.pose.orientation.w = q[0]; o.pose.pose.orientation.x = q[1]; o.pose.pose.orientation.y = q[2]; o.pose.pose.orientation.z = q[3];
  o.pose.pose.position.x = t[0]; o.pose.pose.position.y = t[1]; o.pose.pose.position.z = t[2];
  for (int i = 0; i < 36; i++) o.twist.covariance[i] = r.cov[i];   // consumed at SensorManagerRos.cpp:154-156 (LOAM order rx ry rz tx ty tz)
  loam::OptStatus s; s.header = o.header;                           // exact same stamp
  for (int i = 0; i < 36; i++) s.hessian[i] = r.hessian[i];          // read at degerate_odometry_filter.cpp:30-31
  po.publish(o); ps.publish(s);
}

static void cloudCb(const sensor_msgs::PointCloud2::ConstPtr &msg) {
  // the x / y / z byte offsets come from the message's own fields[] (velodyne_pointcloud PointXYZIR: point_step 22 or 32;
  // CARLA: 3 or 4 floats, carla_to_ros_transforms.py:69-70; the tooling's float32 reshape, downsample_pointcloud.py:45-46)
  int off[3] = {-1, -1, -1};
  for (const auto &f : msg->fields) {
    const int a = f.name == "x" ? 0 : f.name == "y" ? 1 : f.name == "z" ? 2 : -1;
    if (a >= 0 && f.datatype == 7 /* FLOAT32 */) off[a] = (int)f.offset;
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0 || msg->is_bigendian || (msg->point_step & 3)) {
    // a point_step that is not a multiple of 4 (the packed 22-byte PointXYZIR) would need a repack: refused, not garbled
    ROS_WARN_STREAM("vlo: unsupported PointCloud2 layout (need little-endian FLOAT32 x/y/z, point_step % 4 == 0), sweep dropped");
    return;
  }
  const int n = msg->width * msg->height;
  vlo_result odom, mapped;
  // one tick = multiScanRegistration + laserOdometry + (every ioRatio-th sweep) laserMapping with its map maintenance;
  // the map lives on the device (vlo_map_reset at start-up; vlo_map_insert preloads a prior map)
  int rc = vlo_process_scan_pc2(H, msg->data.data(), n, (int)msg->point_step, off[0], off[1], off[2],
                                msg->header.stamp.toSec(), &odom, &mapped);
  if (rc < 0) { ROS_WARN_STREAM("vlo: " << vlo_last_error(H)); }           // a capacity report concerns this sweep only: results below are valid
  if (rc < 0 && rc != VLO_ERR_CAPACITY) return;                             // soft drop, like the reference
  float sum[6], aft[6]; vlo_online_pose(H, sum, aft);
  if (odom.status == VLO_OK)   publish(msg->header, sum, odom, pubOdom, pubStatus);
  if (mapped.status == VLO_OK) publish(msg->header, aft, mapped, pubMapped, pubMapStatus);
}

int main(int argc, char **argv) {
  ros::init(argc, argv, "vlo_loam"); ros::NodeHandle nh("~");
  vlo_config c; vlo_default_config(&c);
  std::string lidar = "VLP-16"; nh.getParam("lidar", lidar); vlo_set_lidar(&c, lidar.c_str());   // loam_params.yaml:22
  nh.getParam("scanPeriod", c.scan_period); nh.getParam("featureRegions", c.feature_regions);    // :3,25 ... same names
  nh.getParam("curvatureRegion", c.curvature_region); nh.getParam("maxCornerSharp", c.max_corner_sharp);             // :26-27
  nh.getParam("maxCornerLessSharp", c.max_corner_less_sharp); nh.getParam("maxSurfaceFlat", c.max_surface_flat);    // :28-29
  nh.getParam("surfaceCurvatureThreshold", c.surface_curvature_threshold);                                         // :30
  nh.getParam("lessFlatFilterSize", c.less_flat_filter_size);                                                      // :31
  bool undistort = false, rotate = false, ring_fields = false;
  // :34 undistortInputCloud = the fork's ego-motion compensation of the input cloud (external prior / motion model / IMU topic,
  // loam.launch:38): not implemented by libvlo -- vlo_create refuses it (VLO_ERR_UNSUPPORTED) rather than do something else.
  // LaserOdometry's own per-point interpolation (upstream's transformToStart) is c.deskew, on by default as upstream.
  nh.getParam("undistortInputCloud", undistort); c.undistort_input_cloud = undistort ? 1 : 0;
  nh.getParam("odomMaxIterations", c.odom_max_iterations); nh.getParam("odomDeltaTAbort", c.odom_delta_t_abort);    // :36-37
  nh.getParam("odomDeltaRAbort", c.odom_delta_r_abort);                                                            // :38
  nh.getParam("mapMaxIterations", c.map_max_iterations); nh.getParam("mapDeltaTAbort", c.map_delta_t_abort);        // :44-45
  nh.getParam("mapDeltaRAbort", c.map_delta_r_abort);                                                              // :46
  std::vector<double> ypr;                                                                                         // :4-5
  nh.getParam("rotateInputCloud", rotate); c.rotate_input = rotate ? 1 : 0;
  if (nh.getParam("inputCloudRotation", ypr) && ypr.size() == 3) for (int i = 0; i < 3; i++) c.input_rotation[i] = (float)ypr[i];
  // :23 useCloudIntensityandRingFields: ring ids from the cloud's `ring` field.  velodyne_pointcloud's PointXYZIR: x y z
  // intensity (float32) + ring (uint16) at byte 16 -> ring_field 16, type 1 (UINT16); an all-float32 cloud with a fifth
  // `ring` column would be ring_field 4, type 0
  nh.getParam("useCloudIntensityandRingFields", ring_fields); c.ring_field = ring_fields ? 16 : -1; c.ring_field_type = 1;
  int ring_off = -1, ring_type = -1;               // a deployment with another layout states it: ~ringFieldOffset (bytes), ~ringFieldType (0 f32 index, 1 u16, 2 u8)
  if (ring_fields && nh.getParam("ringFieldOffset", ring_off) && nh.getParam("ringFieldType", ring_type)) { c.ring_field = ring_off; c.ring_field_type = ring_type; }
  nh.getParam("odomDegenEigVal", c.odom_degen_eig); nh.getParam("mapDegenEigVal", c.map_degen_eig); // :39,53
  nh.getParam("cornerFilterSize", c.corner_filter_size); nh.getParam("surfaceFilterSize", c.surface_filter_size);   // :47-48
  nh.getParam("mapCubeSize", c.map_cube_size); nh.getParam("numNeighborSubmapCubes", c.n_neighbor_cubes);          // :49,52
  nh.getParam("ioRatio", c.io_ratio);                                                                              // :35
  std::vector<int> dims, start;                                                                                    // :50-51
  if (nh.getParam("mapDimensionsInCubes", dims) && dims.size() == 3) std::copy(dims.begin(), dims.end(), c.map_dims);
  if (nh.getParam("mapStartLocationInCubes", start) && start.size() == 3) std::copy(start.begin(), start.end(), c.map_start_cubes);
  c.hessian_order = 0;   // OptStatus.hessian in LOAM's order, what gtsam_fusion_filter was tuned on (fusion_params.yaml:35-36)
  c.max_scans = 2; c.max_points = 1 << 17; c.max_map_points = 1 << 21;
  const int crc = vlo_create(&c, &H);
  if (crc == VLO_ERR_UNSUPPORTED) { ROS_FATAL("undistortInputCloud:=true is not implemented by libvlo (set it to false)"); return 1; }
  if (crc != VLO_OK) { ROS_FATAL("vlo_create failed (no GPU?)"); return 1; }
  vlo_map_reset(H);      // empty maintained map, cube window centred on mapStartLocationInCubes
  pubOdom = nh.advertise<nav_msgs::Odometry>("/laser_odom_to_init_CORRECTED", 5);
  pubStatus = nh.advertise<loam::OptStatus>("/laser_odom_optimization_status", 5);
  pubMapped = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init_CORRECTED", 5);
  pubMapStatus = nh.advertise<loam::OptStatus>("/laser_mapping_optimization_status", 5);
  ros::Subscriber sub = nh.subscribe("/multi_scan_points", 2, cloudCb);   // loam.launch:37 remap
  ros::spin(); vlo_destroy(H); return 0;
}
