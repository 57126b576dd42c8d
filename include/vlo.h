/* vlo.h -- C-ABI of libvlo.so: the B200 (sm_100a) LiDAR-odometry hot path behind gtsam_fusion.
 *
 * Plain C, caller-owned buffers, int status codes, no exceptions, only C scalar and pointer types.  One handle =
 * one CUDA stream + all device memory; re-entrant per handle, not thread-safe within a handle
 * (the reference's callers are single-threaded spinners: gtsam_fusion/src/gtsam_fusion_node.cpp:101,
 * gtsam_fusion/src/degerate_odometry_filter.cpp:50).
 *
 * What each entry point replaces in the reference (file:line under /root/reference):
 *   vlo_config                 gtsam_fusion/config/carla/loam_params.yaml:1-56 (+ fusion_params.yaml:31-36)
 *   vlo_scans_upload / _organise   the `loam` nodelet multiScanRegistration consuming /lidar PointCloud2
 *                              (gtsam_fusion/launch/loam.launch:33-38; raw layout per
 *                              vil_fusion/python/downsample_pointcloud.py:45-46)
 *   vlo_scans_extract          multiScanRegistration feature clouds (/laser_cloud_sharp ... ; loam.launch:33-38)
 *   vlo_register_pairs         the `loam` nodelet laserOdometry (loam.launch:40-45) -> /laser_odom_to_init +
 *                              /laser_odom_optimization_status {hessian[36]} read by
 *                              gtsam_fusion/src/degerate_odometry_filter.cpp:25-31
 *   vlo_map_build / vlo_register_map   the `loam` nodelet laserMapping (loam.launch:47-52)
 *   vlo_result.logdet_* / pass_dopt    degerate_odometry_filter.cpp:30-46 (the gate itself)
 *   vlo_imu_preintegrate_batch VILFusion::IMUManager::getFactor, gtsam_fusion/src/gtsam_fusion/IMUManager.cpp:27-74
 *                              (+ gtsam PreintegratedCombinedMeasurements::integrateMeasurement it calls, :50-52,64)
 *   vlo_pose_diff              SensorManagerRos::poseDiff, gtsam_fusion/src/gtsam_fusion/SensorManagerRos.cpp:122-158
 *   vlo_process_scan           one LOAM tick of the online path
 *   vlo_bag_register_*         offline replay of a whole bag through the same nodelets (loam.launch:54-57)
 *
 * Status: 0 ok; <0 error (vlo_last_error has text); >0 soft conditions mirroring the reference's
 * silent drops (e.g. VLO_SOFT_TOO_FEW_CORR).
 */
#ifndef VLO_H
#define VLO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLO_OK                     0
#define VLO_ERR_INVALID_ARG       -1
#define VLO_ERR_CUDA              -2
#define VLO_ERR_CAPACITY          -3   /* more scans / points / ring points than the handle was created for */
#define VLO_ERR_NO_DEVICE         -4
#define VLO_ERR_STATE             -5   /* call order violated (e.g. register before extract) */
#define VLO_ERR_UNSUPPORTED       -6   /* the configuration asks for something this library does not implement (see vlo_config.undistort_input_cloud) */
#define VLO_SOFT_TOO_FEW_CORR      1   /* < 10 (odometry) / < 50 (mapping) correspondences: no update */
#define VLO_SOFT_DEGENERATE_DROP   2   /* D-optimality gate would drop this odometry message */

#define VLO_MAX_RINGS   128
#define VLO_MAX_REGIONS 8

typedef struct vlo_handle vlo_handle;

/* Field names follow loam_params.yaml (line numbers in comments). */
typedef struct vlo_config {
    /* ---- capacities (not in the reference; they size device memory) ---- */
    int   max_scans;                   /* scans resident at once (batch size); >= 2 */
    int   max_points;                  /* raw points per scan */
    int   max_ring_points;             /* points per ring after organising (shared-memory staging), <= 4096 */
    int   max_map_points;              /* per map cloud (corner / surf) */
    int   max_imu_factors, max_imu_samples;
    int   device;                      /* CUDA device ordinal */
    /* ---- MultiScanRegistration ---- */
    float scan_period;                 /* scanPeriod 0.1 (3) */
    int   n_rings;                     /* lidar (22): VLP-16 16, HDL-32 32, HDL-64E 64 */
    float lower_deg, upper_deg;        /* vertical FoV of the preset */
    int   feature_regions;             /* featureRegions 6 (25) */
    int   curvature_region;            /* curvatureRegion 5 (26) */
    int   max_corner_sharp;            /* maxCornerSharp 2 (27) */
    int   max_corner_less_sharp;       /* maxCornerLessSharp 20 (28) */
    int   max_surface_flat;            /* maxSurfaceFlat 4 (29) */
    float surface_curvature_threshold; /* surfaceCurvatureThreshold 0.1 (30) */
    float less_flat_filter_size;       /* lessFlatFilterSize 0.2 (31) */
    /* ---- LaserOdometry ---- */
    int   odom_max_iterations;         /* odomMaxIterations 25 (36) */
    float odom_delta_t_abort;          /* odomDeltaTAbort 0.05 (37) */
    float odom_delta_r_abort;          /* odomDeltaRAbort 0.05 (38) */
    float odom_degen_eig;              /* odomDegenEigVal 30 (39) */
    int   deskew;                      /* 1: per-point s = relTime/scanPeriod (upstream transformToStart); 0: rigid */
    int   odom_forward_bound_quirk;    /* 1: reproduce upstream's forward partner-loop bound */
    /* ---- LaserMapping ---- */
    int   map_max_iterations;          /* mapMaxIterations 10 (44) */
    float map_delta_t_abort;           /* mapDeltaTAbort 0.05 (45) */
    float map_delta_r_abort;           /* mapDeltaRAbort 0.05 (46) */
    float map_degen_eig;               /* mapDegenEigVal 40 (53) */
    float map_cell_size;               /* voxel-hash cell edge for the map grids (m); not in the reference */
    float odom_cell_size;              /* voxel-hash cell edge of the scan-to-scan surface (less-flat) target grids (m) */
    float odom_corner_cell_size;       /* same for the sparse corner (less-sharp) target grids; 5 m = the 25 m^2 search radius */
    /* ---- gtsam_fusion_filter (fusion_params.yaml:35-36) ---- */
    float dopt_rot_threshold;          /* filter/rot_degen_threshold 11.5 */
    float dopt_trans_threshold;        /* filter/trans_degen_threshold 28.9 */
    /* ---- IMU noise (fusion_params.yaml:22-27, ImuManagerRos.cpp:20-33) ---- */
    double cov_accel, cov_gyro, cov_integration, cov_bias_acc, cov_bias_omega, cov_bias_acc_omega_int;
    /* ---- LaserMapping, map side (loam_params.yaml:35,47-52) ---- */
    float corner_filter_size;          /* cornerFilterSize 0.2 (47): VoxelGrid leaf of the corner stack and the corner map; 0 = off */
    float surface_filter_size;         /* surfaceFilterSize 0.4 (48) */
    float map_cube_size;               /* mapCubeSize 10.0 (49) */
    int   map_dims[3];                 /* mapDimensionsInCubes [101,51,101] (50) */
    int   map_start_cubes[3];          /* mapStartLocationInCubes [50,25,50] (51) */
    int   n_neighbor_cubes;            /* numNeighborSubmapCubes 5 (52); <= 5 */
    int   io_ratio;                    /* ioRatio 2 (35): the online tick runs LaserMapping on every io_ratio-th sweep */
    int   hessian_order;               /* 0: OptStatus.hessian in LOAM's order (rx ry rz tx ty tz); 1: (tx ty tz rx ry rz), i.e.
                                          block(0,0) really is translation as degerate_odometry_filter.cpp:32-33 labels it */
    /* ---- MultiScanRegistration input options (loam_params.yaml:4-5,23) ---- */
    int   rotate_input;                /* rotateInputCloud false (4) */
    float input_rotation[3];           /* inputCloudRotation [0,0,0] (5): yaw pitch roll (rad); p' = Rz(yaw) Ry(pitch) Rx(roll) p, ROS frame */
    int   ring_field;                  /* useCloudIntensityandRingFields (23): where the point's `ring` field is (ring ids as delivered by
                                          the driver, e.g. Velodyne / Ouster / Bpearl clouds); -1 = ring from the vertical angle.
                                          ring_field_type 0: index in float32 units of a FLOAT32 field; 1 / 2: BYTE offset of a
                                          UINT16 / UINT8 field (PointCloud2 fields[].offset of the usual uint16 `ring`) */
    int   ring_field_type;
    /* undistortInputCloud (34): the fork's ego-motion compensation of the INPUT cloud inside MultiScanRegistration from an
     * external prior / motion model / the IMU topic the nodelet subscribes to (loam.launch:38, imuHistorySize :24).  Its
     * code is not in the reference and it is NOT implemented here: vlo_create refuses a non-zero value with
     * VLO_ERR_UNSUPPORTED instead of silently doing something else.  (`deskew` above is a different thing: LaserOdometry's own
     * per-point interpolation of the sweep motion, upstream's transformToStart, which is always on upstream.) */
    int   undistort_input_cloud;
} vlo_config;

/* One registration result = one nav_msgs/Odometry + loam/OptStatus pair of the reference. */
typedef struct vlo_result {
    float  transform[6];       /* rx ry rz tx ty tz, LOAM order and axes (x left, y up, z forward) */
    float  hessian[36];        /* OptStatus.hessian: AtA of the last linearisation, row-major, LOAM order */
    float  eig[6];             /* eigenvalues of AtA at iteration 0, ascending */
    float  P[36];              /* solution-remapping projection (identity when not degenerate) */
    int    is_degenerate;      /* some eigenvalue < odomDegenEigVal / mapDegenEigVal */
    int    iterations;
    int    n_corr_edge, n_corr_plane;
    float  logdet_rot, logdet_trans;   /* degerate_odometry_filter.cpp:35-36 (block(3,3), block(0,0)) */
    int    pass_dopt;          /* 1 = the filter would publish this odometry (:39-46) */
    int    status;             /* VLO_OK or VLO_SOFT_* */
    double cov[36];            /* twist.covariance: sigma^2 (AtA)^-1 */
} vlo_result;

typedef struct vlo_feature_counts {
    int n_valid;               /* organised points */
    int n_sharp, n_less_sharp, n_flat, n_less_flat;
} vlo_feature_counts;

typedef struct vlo_preint {    /* gtsam::PreintegratedCombinedMeasurements as POD */
    double dR[9];              /* deltaRij, row-major */
    double dP[3], dV[3];
    double dR_dbg[9], dP_dba[9], dP_dbg[9], dV_dba[9], dV_dbg[9];
    double cov[225];           /* preintMeasCov, order (theta, p, v, bias_acc, bias_gyro) */
    double dt;                 /* deltaTij */
    int    n_integrated;
    int    _pad;
} vlo_preint;

/* ---------------------------------------------------------------- lifecycle */
void        vlo_default_config(vlo_config *cfg);            /* loam_params.yaml / fusion_params.yaml defaults, VLP-16 */
int         vlo_set_lidar(vlo_config *cfg, const char *name); /* "VLP-16" | "HDL-32" | "HDL-64E" | "O1-16" | "O1-64" | "Bperl-32" (loam_params.yaml:22) */
int         vlo_create(const vlo_config *cfg, vlo_handle **out);
void        vlo_destroy(vlo_handle *h);
const char *vlo_last_error(const vlo_handle *h);
const char *vlo_version(void);
int         vlo_synchronize(vlo_handle *h);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
long long   vlo_launch_count(const vlo_handle *h);

/* per-stage device timing (CUDA events on the handle's stream around every launch group); stage
 * names via vlo_stage_name(0 .. vlo_stage_count()-1); vlo_get_stage_times drains the accumulators */
int         vlo_set_profiling(vlo_handle *h, int enable);
int         vlo_stage_count(void);
const char *vlo_stage_name(int stage);
int         vlo_get_stage_times(vlo_handle *h, float *ms, int *launches);
/* the handle's cudaStream_t (so a caller can record its own events on it) */
void       *vlo_stream(vlo_handle *h);

/* ---------------------------------------------------------------- scans (batch of B <= max_scans) */
/* raw: concatenated clouds, `stride_floats` float32 per point, x y z first (ROS axes), scan s owns
 * points [offsets[s], offsets[s+1]).  `on_device` != 0: raw is a device pointer (offsets stay host). */
int vlo_scans_upload(vlo_handle *h, const float *raw, const int *offsets, int n_scans, int stride_floats, int on_device);
/* the same for a sensor_msgs/PointCloud2 payload as it arrives (Appendix C of SURVEY.md; the reference tooling reshapes it as
 * float32[-1, point_step/4], vil_fusion/python/downsample_pointcloud.py:45-46): `data` = concatenated `data` blobs,
 * point_step bytes per point, FLOAT32 x / y / z at the byte offsets of fields[] (multiples of 4) */
int vlo_scans_upload_pc2(vlo_handle *h, const void *data, const int *offsets, int n_scans, int point_step,
                         int x_offset, int y_offset, int z_offset, int on_device);
int vlo_scans_organise(vlo_handle *h);                      /* K0: axis swap, ring id, rel-time, ring-major float4 */
int vlo_scans_extract(vlo_handle *h);                       /* K1: curvature, masks, sector selection, less-flat voxel filter */
int vlo_scans_counts(vlo_handle *h, vlo_feature_counts *counts /* n_scans */);
/* copy-backs for parity tests and tooling (any pointer may be NULL) */
int vlo_scan_get_cloud(vlo_handle *h, int scan, float *cloud_xyzi /* n_valid*4 */, int *ring_start /* R+1 */, int *src_index);
int vlo_scan_get_features(vlo_handle *h, int scan, int8_t *label, float *curvature, uint8_t *picked,
                          int *sharp_idx, int *less_sharp_idx, int *flat_idx, float *less_flat_xyzi,
                          int *less_sharp_ring_start, int *less_flat_ring_start);

/* ---------------------------------------------------------------- scan-to-scan (LaserOdometry) */
/* For each pair p = 0..n_pairs-1 registers scan cur[p] against scan last[p] (indices into the
 * resident batch).  seeds: n_pairs*6 or NULL (zeros).  last_transforms: n_pairs*6 or NULL -- when
 * given (online mode) the `last` feature clouds are first moved to their sweep end with it
 * (transformToEnd).  Results are written to `out` (host). */
int vlo_register_pairs(vlo_handle *h, const int *last, const int *cur, int n_pairs,
                       const float *seeds, const float *last_transforms, vlo_result *out);
/* the same without the host synchronisation (rigid batches: no last_transforms): see vlo_register_map_enqueue below -- the records
 * land in `out_pinned` when the stream gets there; vlo_synchronize, then vlo_results_finish(h, out, n_pairs) */
int vlo_register_pairs_enqueue(vlo_handle *h, const int *last, const int *cur, int n_pairs, const float *seeds, vlo_result *out_pinned);
/* parity hooks: with tracing enabled, the correspondence indices of association round `round`
 * (iterations 0,5,10,.. -> round 0,1,2,..) of pair `pair` of the last vlo_register_pairs call */
int vlo_set_trace(vlo_handle *h, int enable);
int vlo_pair_get_correspondences(vlo_handle *h, int pair, int round, int *corner_idx /* n_sharp*2 */, int *surf_idx /* n_flat*3 */);

/* ---------------------------------------------------------------- scan-to-map (LaserMapping) */
int vlo_map_build(vlo_handle *h, const float *corner_xyzi, int n_corner, const float *surf_xyzi, int n_surf, int on_device);
/* registers the down-sampled corner (less sharp) / surface (less flat) stacks of resident scans against the map */
int vlo_register_map(vlo_handle *h, const int *scans, int n, const float *seeds /* n*6 transformTobeMapped */, vlo_result *out);
/* the same without the host synchronisation: everything, including the copy of the n result records into `out`, is only
 * ENQUEUED on the handle's stream (`out`: pinned host memory that stays valid until vlo_synchronize; the next batch can be
 * uploaded and registered meanwhile).  After vlo_synchronize: vlo_results_finish(h, out, n) completes the records on the host
 * (the float64 covariance) and returns the soft status vlo_register_map would have returned. */
int vlo_register_map_enqueue(vlo_handle *h, const int *scans, int n, const float *seeds, vlo_result *out_pinned);
int vlo_results_finish(vlo_handle *h, vlo_result *out, int n);
int vlo_map_get_correspondences(vlo_handle *h, int slot, int *corner_idx5, int *surf_idx5);
/* exact k-NN service on the map grids (k = 1 or 5, neighbours with d2 < max_d2): which = 0 corner, 1 surf;
 * queries host xyzi; missing neighbours are idx -1 / d2 +inf */
int vlo_map_knn(vlo_handle *h, int which, const float *queries_xyzi, int nq, int k, float max_d2, int *idx, float *d2);

/* ---- maintained map: BasicLaserMapping::process of the `loam` nodelet laserMapping (loam.launch:47-52), map side.
 * The map lives on the device as voxel centroids (corner leaf cornerFilterSize, surface leaf surfaceFilterSize) tagged
 * with their cube (mapCubeSize); a point's index = order of voxel creation.
 *   vlo_map_reset    empties the map and re-centres the cube window (mapStartLocationInCubes)
 *   vlo_map_insert   upstream's insertion step alone: points (host, sensor frame) moved with pose6, binned into cubes,
 *                    touched voxels re-filtered (e.g. to preload a prior map)
 *   vlo_map_process  one LaserMapping tick for resident scan `scan`: window shift + FOV-valid sub-map around `seed6`,
 *                    stack down-sampling, optimisation against the sub-map (skipped when it holds <= 10 corner or <= 100
 *                    surface points), insertion with the optimised pose.  info[6] (optional) = down-sampled stack sizes,
 *                    sub-map sizes, map sizes after insertion (corner, surface each)
 *   vlo_map_get_points  copy-back for parity tests / tooling: xyz0 per point + packed cube (bit 30 = evicted) */
int vlo_map_reset(vlo_handle *h);
int vlo_map_insert(vlo_handle *h, const float *corner_xyzi, int n_corner, const float *surf_xyzi, int n_surf, const float *pose6);
int vlo_map_process(vlo_handle *h, int scan, const float *seed6, vlo_result *out, int *info);
int vlo_map_size(vlo_handle *h, int *n_corner, int *n_surf);
int vlo_map_get_points(vlo_handle *h, int which, float *xyzi, int *cube);
/* down-sampled stacks (cornerFilterSize / surfaceFilterSize VoxelGrid of the less-sharp / less-flat clouds) of a resident
 * scan: the query clouds of vlo_register_map / vlo_map_process.  Any pointer may be NULL. */
int vlo_scan_get_stack(vlo_handle *h, int scan, float *corner_xyzi, int *n_corner, float *surf_xyzi, int *n_surf);
/* sizes of the down-sampled stacks of every resident scan (n_scans ints each) */
int vlo_scans_stack_counts(vlo_handle *h, int *n_corner, int *n_surf);

/* ---------------------------------------------------------------- online tick */
/* One LOAM tick: organise + extract + scan-to-scan against the previous tick (seeded with the
 * previous transform) [+ scan-to-map when a map is resident]; `odom` / `mapped` may be NULL. */
int vlo_process_scan(vlo_handle *h, const float *raw, int n_points, int stride_floats, double stamp,
                     vlo_result *odom, vlo_result *mapped);
/* the same tick for a sensor_msgs/PointCloud2 payload as it arrives: point_step bytes per point, FLOAT32 x / y / z at the byte
 * offsets of fields[] (multiples of 4); ring ids per vlo_config.ring_field */
int vlo_process_scan_pc2(vlo_handle *h, const void *data, int n_points, int point_step, int x_offset, int y_offset, int z_offset,
                         double stamp, vlo_result *odom, vlo_result *mapped);
int vlo_online_reset(vlo_handle *h);
/* accumulated odometry pose (transformSum) and last mapped pose (transformAftMapped), LOAM order/axes */
int vlo_online_pose(vlo_handle *h, float *sum6, float *mapped6);
int vlo_online_set_map_pose(vlo_handle *h, const float *pose6);

/* ---------------------------------------------------------------- whole-bag streaming */
/* Offline reprocessing of a bag handed over as host batches (each <= max_scans/2 scans; `raw`/`offsets` as in
 * vlo_scans_upload, host memory, pinned for full PCIe rate).  Batch k+1 is copied to the device on a second
 * stream while batch k's kernels run; one synchronisation at the end.  Replaces replaying the bag through the
 * `loam` nodelets (gtsam_fusion/launch/loam.launch:54-57 `rosbag play`; vil_fusion/python/quick_autoexperiments.py:49-50).
 *   vlo_bag_register_map  : scan-to-map of every scan from seeds[n_scans*6]; out = sum(n_scans) records in order
 *   vlo_bag_register_pairs: scan-to-scan of the consecutive pairs (i, i+1) inside each batch from seeds[(n_scans-1)*6]
 *                           (NULL = zero seed); out = sum(n_scans - 1) records; overlap batches by one frame to chain */
typedef struct vlo_bag_batch {
    const float *raw;          /* concatenated clouds of this batch: host memory (pinned, for the copy to overlap) or device memory */
    const int   *offsets;      /* n_scans + 1 point offsets into raw */
    int          n_scans;
    const float *seeds;
} vlo_bag_batch;
int vlo_bag_register_map(vlo_handle *h, const vlo_bag_batch *batches, int n_batches, int stride_floats, vlo_result *out);
int vlo_bag_register_pairs(vlo_handle *h, const vlo_bag_batch *batches, int n_batches, int stride_floats, vlo_result *out);

/* ---------------------------------------------------------------- IMU */
/* Batched IMUManager::getFactor over one time-sorted sample stream (host pointers):
 * factor f integrates the window (t0[f], t1[f]] exactly as IMUManager.cpp:33-66 does on a buffer that
 * holds every sample; bias6 = acc(3), gyro(3) (one bias for the batch). */
int vlo_imu_preintegrate_batch(vlo_handle *h, const double *t, const double *acc, const double *gyro, int n_samples,
                               const double *t0, const double *t1, const double *bias6, int n_factors, vlo_preint *out);

/* ---------------------------------------------------------------- helpers */
/* pose = x y z qw qx qy qz (host, float64); SensorManagerRos.cpp:122-158 */
void vlo_pose_diff(const double *before7, const double *after7, double *out7);
/* degerate_odometry_filter.cpp:30-46 on a host hessian (float32[36]); returns 1 if published */
int  vlo_dopt_gate(const float *hessian36, double rot_thr, double trans_thr, float *logdet_rot, float *logdet_trans);
/* transformSum accumulation (LOAM axes); fudge = upstream's 1.05 factor on ry / tz (1.0 = exact) */
void vlo_accumulate_pose(const float *sum_in6, const float *transform6, float fudge, float *sum_out6);

#ifdef __cplusplus
}
#endif
#endif
