// temporary: entry points not implemented yet
#include "vlo_internal.cuh"
#define NOTIMPL(h) do { if (h) (h)->err = "not implemented"; return VLO_ERR_STATE; } while (0)
extern "C" int vlo_map_build(vlo_handle *h, const float *, int, const float *, int, int) { NOTIMPL(h); }
extern "C" int vlo_register_map(vlo_handle *h, const int *, int, const float *, vlo_result *) { NOTIMPL(h); }
extern "C" int vlo_map_get_correspondences(vlo_handle *h, int, int *, int *) { NOTIMPL(h); }
extern "C" int vlo_map_knn(vlo_handle *h, int, const float *, int, int, int *, float *) { NOTIMPL(h); }
extern "C" int vlo_process_scan(vlo_handle *h, const float *, int, int, double, vlo_result *, vlo_result *) { NOTIMPL(h); }
extern "C" int vlo_imu_preintegrate_batch(vlo_handle *h, const double *, const double *, const double *, int, const double *, const double *, const double *, int, vlo_preint *) { NOTIMPL(h); }
extern "C" void vlo_pose_diff(const double *, const double *, double *) {}
extern "C" int vlo_dopt_gate(const float *, double, double, float *, float *) { return 0; }
extern "C" void vlo_accumulate_pose(const float *, const float *, float, float *) {}
