/* TEST INFRASTRUCTURE -- CPU oracle for the LiDAR-odometry hot path of gtsam_fusion.
 *
 * PARITY UNPINNED for the LOAM stages: the arithmetic of feature extraction, correspondence
 * search, linearisation and the eigen-degeneracy remap is NOT in /root/reference; it lives in
 * the un-vendored, un-pinned catkin dependency `loam`
 * (gtsam_fusion/package.xml:26, gtsam_fusion/README.md:21-25 -> ItsTimmy/cerberus_loam_dev
 * @feature/publish_covariance, a fork of laboshinl/loam_velodyne).  This oracle restates that
 * published algorithm and is anchored on the reference's own call
 * sites / parameters: gtsam_fusion/config/carla/loam_params.yaml:1-56,
 * gtsam_fusion/launch/loam.launch:31-58.
 * PINNED parts: the D-optimality gate (gtsam_fusion/src/degerate_odometry_filter.cpp:29-48),
 * the IMU window/interpolation rule (gtsam_fusion/src/gtsam_fusion/IMUManager.cpp:27-74) with
 * the known-answer test gtsam_fusion/test/UnitTests.cpp:30-66, and poseDiff
 * (gtsam_fusion/src/gtsam_fusion/SensorManagerRos.cpp:122-158, KAT UnitTests.cpp:228-233).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Build: see oracle/Makefile (gcc -O3 -ffp-contract=off, no fast-math).
 */
#ifndef VLO_ORACLE_H
#define VLO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z, w; } orc_pt;   /* w = intensity = ring + relTime */

/* Names follow loam_params.yaml (line numbers in comments). */
typedef struct {
    float scan_period;                 /* scanPeriod 0.1 (3) */
    int   n_rings;                     /* lidar preset (22): VLP-16 16, HDL-32 32, HDL-64E 64 */
    float lower_deg, upper_deg;        /* preset vertical FoV */
    int   feature_regions;             /* featureRegions 6 (25) */
    int   curvature_region;            /* curvatureRegion 5 (26) */
    int   max_corner_sharp;            /* maxCornerSharp 2 (27) */
    int   max_corner_less_sharp;       /* maxCornerLessSharp 20 (28) */
    int   max_surface_flat;            /* maxSurfaceFlat 4 (29) */
    float surface_curvature_threshold; /* 0.1 (30) */
    float less_flat_filter_size;       /* 0.2 (31) */
    int   odom_max_iterations;         /* 25 (36) */
    float odom_delta_t_abort;          /* 0.05 (37) */
    float odom_delta_r_abort;          /* 0.05 (38) */
    float odom_degen_eig;              /* 30 (39) */
    int   map_max_iterations;          /* 10 (44) */
    float map_delta_t_abort;           /* 0.05 (45) */
    float map_delta_r_abort;           /* 0.05 (46) */
    float map_degen_eig;               /* 40 (53) */
    int   deskew;                      /* 1: s = relTime/scanPeriod per point (upstream); 0: rigid, s = 1 */
    int   odom_forward_bound_quirk;    /* 1: forward partner loop bounded by #query features (upstream quirk) */
    float dopt_rot_threshold;          /* fusion_params.yaml:35  11.5 */
    float dopt_trans_threshold;        /* fusion_params.yaml:36  28.9 */
    /* LaserMapping map side (loam_params.yaml:35,47-52) */
    float corner_filter_size;          /* cornerFilterSize 0.2 (47); 0 = stack not down-sampled */
    float surface_filter_size;         /* surfaceFilterSize 0.4 (48) */
    float map_cube_size;               /* mapCubeSize 10.0 (49) */
    int   map_dims[3];                 /* mapDimensionsInCubes [101,51,101] (50) */
    int   map_start_cubes[3];          /* mapStartLocationInCubes [50,25,50] (51) */
    int   n_neighbor_cubes;            /* numNeighborSubmapCubes 5 (52) */
    int   io_ratio;                    /* ioRatio 2 (35): mapping runs on every io_ratio-th sweep */
    int   rotate_input;                /* rotateInputCloud false (4) */
    float input_rotation[3];           /* inputCloudRotation [0,0,0] (5): yaw pitch roll (rad), p' = Rz(yaw) Ry(pitch) Rx(roll) p in the ROS frame */
    int   ring_field;                  /* useCloudIntensityandRingFields (23): float index of a FLOAT32 ring field in the point, -1 = ring from the vertical angle */
    int   ring_field_type;             /* 0: ring_field counts float32 units (FLOAT32 field); 1 / 2: ring_field is the BYTE offset of a UINT16 / UINT8 field */
} orc_config;

void orc_default_config(orc_config *c);

/* ---- A.1 MultiScanRegistration::process: axis swap, ring id, rel-time, ring-major layout ---- */
/* raw: n points, `stride` floats each (x,y,z first, ROS frame).  out: ring-major cloud in the
 * LOAM frame; ring_start[R+1] exclusive offsets; src_index (optional) = raw index of each out point.
 * returns number of valid points. */
int orc_organise(const orc_config *c, const float *raw, int n, int stride,
                 orc_pt *out, int *ring_start, int *src_index);

/* ---- A.2/A.3 BasicScanRegistration::extractFeatures ---- */
/* label int8: 2 sharp, 1 less sharp, 0 default, -1 flat.  curvature is written only inside
 * sector ranges (else 0).  picked = final neighbour-picked mask.  Lists hold cloud indices in
 * upstream push order (ring, sector, pick order).  less_flat holds the per-ring voxel-grid
 * centroids (leaf less_flat_filter_size) with lflat_ring_start[R+1]; lsharp_ring_start[R+1] likewise. */
typedef struct {
    int n_sharp, n_less_sharp, n_flat, n_less_flat;
} orc_feature_counts;

void orc_extract(const orc_config *c, const orc_pt *cloud, const int *ring_start,
                 int8_t *label, float *curvature, uint8_t *picked,
                 int *sharp_idx, int *less_sharp_idx, int *flat_idx,
                 orc_pt *less_flat, int *lsharp_ring_start, int *lflat_ring_start,
                 orc_feature_counts *counts);

/* ---- nearest neighbours: exact, (d2, index) lexicographic order ---- */
void orc_knn_brute(const orc_pt *cloud, int n, const orc_pt *q, int nq, int k, int *idx, float *d2);
typedef struct orc_kdtree orc_kdtree;
orc_kdtree *orc_kdtree_build(const orc_pt *cloud, int n);
void orc_kdtree_free(orc_kdtree *t);
void orc_kdtree_knn(const orc_kdtree *t, const orc_pt *q, int nq, int k, int *idx, float *d2);

/* ---- dense 6x6 (A.7) ---- */
void orc_solve6_colpiv_qr(const float *A /*36 row-major*/, const float *b, float *x);
void orc_eig6_jacobi(const float *A, float *eval /*ascending*/, float *evec /*row i = eigenvector i*/);
int  orc_degeneracy(const float *A, float thr, float *eval, float *P /*36*/);
/* degerate_odometry_filter.cpp:30-46: returns 1 if the odometry message is published (passes) */
int  orc_dopt_gate(const float *hessian36, double rot_thr, double trans_thr, float *logdet_rot, float *logdet_trans);

/* ---- scan-to-scan (A.4-A.7) ---- */
typedef struct {
    float transform[6];        /* rx ry rz tx ty tz (LOAM order) */
    float hessian[36];         /* AtA of the last linearisation */
    float eig[6];              /* eigenvalues at iteration 0 */
    float P[36];               /* projection used for remapping */
    int   is_degenerate;
    int   iterations;          /* number of GN iterations executed */
    int   n_corr_edge, n_corr_plane; /* of the last linearisation */
    float logdet_rot, logdet_trans;
    int   pass_dopt;
    double cov[36];
    int   status;              /* 0 ok, 1 too few correspondences at every iteration */
} orc_reg_result;

/* cur_sharp/cur_flat: feature points of the current sweep (LOAM frame, intensity ring+relTime).
 * last_corner/last_surf: previous sweep's less-sharp / less-flat clouds (already at sweep end,
 * intensity = ring), ring-major with ring_start[R+1].  seed: initial transform.
 * trace (optional): per association round r (iter 0,5,10,..): corner idx (2 per query) then surf idx
 * (3 per query) appended; trace_T: transform after every iteration (6 floats each). */
void orc_odometry_register(const orc_config *c,
                           const orc_pt *cur_sharp, int n_sharp, const orc_pt *cur_flat, int n_flat,
                           const orc_pt *last_corner, int n_last_corner, const int *last_corner_ring_start,
                           const orc_pt *last_surf, int n_last_surf, const int *last_surf_ring_start,
                           const float *seed, int use_kdtree,
                           orc_reg_result *res, int *trace_idx, float *trace_T);

void orc_odometry_associate(const orc_config *c, const float *T,
                            const orc_pt *cur_sharp, int n_sharp, const orc_pt *cur_flat, int n_flat,
                            const orc_pt *last_corner, int n_last_corner,
                            const orc_pt *last_surf, int n_last_surf, int use_kdtree,
                            int *corner_idx /*2 per*/, int *surf_idx /*3 per*/);

void orc_transform_to_start(const orc_config *c, const float *T, const orc_pt *in, int n, orc_pt *out);
void orc_transform_to_end(const orc_config *c, const float *T, orc_pt *cloud, int n);
/* transformSum accumulation (BasicLaserOdometry tail, no IMU, fudge factor explicit) */
void orc_accumulate_pose(const float *sum_in, const float *T, float fudge, float *sum_out);

/* ---- scan-to-map (A.8) ---- */
void orc_mapping_register(const orc_config *c,
                          const orc_pt *corner_q, int n_corner_q, const orc_pt *surf_q, int n_surf_q,
                          const orc_pt *corner_map, int n_corner_map, const orc_pt *surf_map, int n_surf_map,
                          const float *seed /*transformTobeMapped*/, int use_kdtree,
                          orc_reg_result *res, int *trace_idx /*5 per query, first association*/, float *trace_T);

/* pcl::VoxelGrid restated (V1/V2 of scan_registration.c): centroids in order of first appearance; returns count */
int orc_voxel_downsample(const orc_pt *in, int n, float leaf, orc_pt *out);

/* ---- LaserMapping map maintenance (A.8 map side; laser_map.c) ---- */
typedef struct orc_lmap orc_lmap;
orc_lmap *orc_lmap_create(const orc_config *c, int cap);
void orc_lmap_free(orc_lmap *m);
int  orc_lmap_size(const orc_lmap *m, int which);
void orc_lmap_get(const orc_lmap *m, int which, orc_pt *pts, int *cube);
void orc_lmap_window(const orc_lmap *m, int *cen3);
int  orc_lmap_submap(const orc_lmap *m, int which, int *ids);
void orc_lmap_insert(orc_lmap *m, const orc_pt *corner, int nc, const orc_pt *surf, int ns, const float *T);
void orc_lmap_select(orc_lmap *m, const float *T, int *centre_abs, uint8_t *mask);
void orc_lmap_process(orc_lmap *m, const orc_pt *corner_stack, int nc, const orc_pt *surf_stack, int ns,
                      const float *seed, orc_reg_result *res, int *info);

/* ---- IMU (Appendix B + IMUManager.cpp:27-74) ---- */
typedef struct {
    double cov_accel, cov_gyro, cov_integration, cov_bias_acc, cov_bias_omega, cov_bias_acc_omega_int;
} orc_imu_params;   /* isotropic, ImuManagerRos.cpp:20-33 */

typedef struct {
    double dR[9];       /* deltaRij row-major */
    double dP[3], dV[3];
    double dR_dbg[9], dP_dba[9], dP_dbg[9], dV_dba[9], dV_dbg[9];
    double cov[225];    /* preintMeasCov, order (theta, p, v, ba, bg) */
    double dt;          /* deltaTij */
    int    n_integrated;
} orc_preint;

void orc_imu_get_factor(const orc_imu_params *p, const double *t, const double *acc, const double *gyro, int n,
                        double t0, double t1, const double *bias6 /*acc(3) gyro(3)*/, orc_preint *out);
void orc_imu_batch(const orc_imu_params *p, const double *t, const double *acc, const double *gyro, int n,
                   const double *t0, const double *t1, const double *bias6, int n_factors, orc_preint *out,
                   int n_threads);

/* ---- SensorManagerRos::poseDiff (SensorManagerRos.cpp:122-158) ---- */
/* pose = x y z qw qx qy qz */
void orc_pose_diff(const double *before7, const double *after7, double *out7);

#ifdef __cplusplus
}
#endif
#endif
