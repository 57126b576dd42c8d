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
#include "vlo.h"

static vlo_handle *H;
static ros::Publisher pubOdom, pubStatus, pubMapped, pubMapStatus;

static void publish(const std_msgs::Header &hdr, const float pose[6], const vlo_result &r,
                    ros::Publisher &po, ros::Publisher &ps) {
  nav_msgs::Odometry o; o.header = hdr; o.header.frame_id = "/loam_init"; o.child_frame_id = "/laser_odom";
  // LOAM axes (x left, y up, z fwd) -> ROS (x,y,z) = LOAM (z,x,y): vil_fusion/python/loam_frame_transform.py:52-90
  geometry_msgs::Quaternion q = tf::createQuaternionMsgFromRollPitchYaw(pose[2], -pose[0], -pose[1]);
  o.pose.pose.orientation.x = -q.y; o.pose.pose.orientation.y = -q.z; o.pose.pose.orientation.z = q.x; o.pose.pose.orientation.w = q.w;
  o.pose.pose.position.x = pose[3]; o.pose.pose.position.y = pose[4]; o.pose.pose.position.z = pose[5];
  for (int i = 0; i < 36; i++) o.twist.covariance[i] = r.cov[i];   // consumed at SensorManagerRos.cpp:154-156
  loam::OptStatus s; s.header = o.header;                           // exact same stamp
  for (int i = 0; i < 36; i++) s.hessian[i] = r.hessian[i];          // read at degerate_odometry_filter.cpp:30-31
  po.publish(o); ps.publish(s);
}

static void cloudCb(const sensor_msgs::PointCloud2::ConstPtr &msg) {
  // all-float32 little-endian fields, xyz first (downsample_pointcloud.py:45-46; CARLA: exactly 3 floats,
  // carla_to_ros_transforms.py:69-70).  A driver with another field order goes through
  // vlo_scans_upload_pc2(H, data, offsets, 1, point_step, fields["x"].offset, fields["y"].offset, fields["z"].offset, 0)
  // + vlo_scans_organise / _extract / vlo_register_pairs / vlo_map_process instead of the fused tick.
  const int stride = msg->point_step / 4, n = msg->width * msg->height;
  vlo_result odom, mapped;
  // one tick = multiScanRegistration + laserOdometry + (every ioRatio-th sweep) laserMapping with its map maintenance;
  // the map lives on the device (vlo_map_reset at start-up; vlo_map_insert preloads a prior map)
  int rc = vlo_process_scan(H, reinterpret_cast<const float *>(msg->data.data()), n, stride,
                            msg->header.stamp.toSec(), &odom, &mapped);
  if (rc < 0) { ROS_WARN_STREAM("vlo: " << vlo_last_error(H)); return; }   // soft drop, like the reference
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
  bool undistort = true, rotate = false, ring_fields = false;
  nh.getParam("undistortInputCloud", undistort); c.deskew = undistort ? 1 : 0;                                     // :34
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
  nh.getParam("odomDegenEigVal", c.odom_degen_eig); nh.getParam("mapDegenEigVal", c.map_degen_eig); // :39,53
  nh.getParam("cornerFilterSize", c.corner_filter_size); nh.getParam("surfaceFilterSize", c.surface_filter_size);   // :47-48
  nh.getParam("mapCubeSize", c.map_cube_size); nh.getParam("numNeighborSubmapCubes", c.n_neighbor_cubes);          // :49,52
  nh.getParam("ioRatio", c.io_ratio);                                                                              // :35
  std::vector<int> dims, start;                                                                                    // :50-51
  if (nh.getParam("mapDimensionsInCubes", dims) && dims.size() == 3) std::copy(dims.begin(), dims.end(), c.map_dims);
  if (nh.getParam("mapStartLocationInCubes", start) && start.size() == 3) std::copy(start.begin(), start.end(), c.map_start_cubes);
  c.hessian_order = 0;   // OptStatus.hessian in LOAM's order, what gtsam_fusion_filter was tuned on (fusion_params.yaml:35-36)
  c.max_scans = 2; c.max_points = 1 << 17; c.max_map_points = 1 << 21;
  if (vlo_create(&c, &H) != VLO_OK) { ROS_FATAL("vlo_create failed (no GPU?)"); return 1; }
  vlo_map_reset(H);      // empty maintained map, cube window centred on mapStartLocationInCubes
  pubOdom = nh.advertise<nav_msgs::Odometry>("/laser_odom_to_init_CORRECTED", 5);
  pubStatus = nh.advertise<loam::OptStatus>("/laser_odom_optimization_status", 5);
  pubMapped = nh.advertise<nav_msgs::Odometry>("/aft_mapped_to_init_CORRECTED", 5);
  pubMapStatus = nh.advertise<loam::OptStatus>("/laser_mapping_optimization_status", 5);
  ros::Subscriber sub = nh.subscribe("/multi_scan_points", 2, cloudCb);   // loam.launch:37 remap
  ros::spin(); vlo_destroy(H); return 0;
}
