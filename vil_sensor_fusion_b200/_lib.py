"""ctypes view of include/vlo.h.  Loading fails loudly when libvlo.so is missing: there is no
CPU fallback in the product."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VLO_LIB_PATH: load another build of the same library (tuning experiments: tools/build_variant.py)
LIB_PATH = os.environ.get("VLO_LIB_PATH") or os.path.join(HERE, "lib", "libvlo.so")


class Config(C.Structure):
    _fields_ = [
        ("max_scans", C.c_int), ("max_points", C.c_int), ("max_ring_points", C.c_int), ("max_map_points", C.c_int),
        ("max_imu_factors", C.c_int), ("max_imu_samples", C.c_int), ("device", C.c_int),
        ("scan_period", C.c_float), ("n_rings", C.c_int), ("lower_deg", C.c_float), ("upper_deg", C.c_float),
        ("feature_regions", C.c_int), ("curvature_region", C.c_int), ("max_corner_sharp", C.c_int),
        ("max_corner_less_sharp", C.c_int), ("max_surface_flat", C.c_int),
        ("surface_curvature_threshold", C.c_float), ("less_flat_filter_size", C.c_float),
        ("odom_max_iterations", C.c_int), ("odom_delta_t_abort", C.c_float), ("odom_delta_r_abort", C.c_float),
        ("odom_degen_eig", C.c_float), ("deskew", C.c_int), ("odom_forward_bound_quirk", C.c_int),
        ("map_max_iterations", C.c_int), ("map_delta_t_abort", C.c_float), ("map_delta_r_abort", C.c_float),
        ("map_degen_eig", C.c_float), ("map_cell_size", C.c_float), ("odom_cell_size", C.c_float), ("odom_corner_cell_size", C.c_float),
        ("dopt_rot_threshold", C.c_float), ("dopt_trans_threshold", C.c_float),
        ("cov_accel", C.c_double), ("cov_gyro", C.c_double), ("cov_integration", C.c_double),
        ("cov_bias_acc", C.c_double), ("cov_bias_omega", C.c_double), ("cov_bias_acc_omega_int", C.c_double),
        ("corner_filter_size", C.c_float), ("surface_filter_size", C.c_float), ("map_cube_size", C.c_float),
        ("map_dims", C.c_int * 3), ("map_start_cubes", C.c_int * 3), ("n_neighbor_cubes", C.c_int), ("io_ratio", C.c_int),
        ("hessian_order", C.c_int),
        ("rotate_input", C.c_int), ("input_rotation", C.c_float * 3), ("ring_field", C.c_int), ("ring_field_type", C.c_int),
        ("undistort_input_cloud", C.c_int),
    ]


class Result(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 6), ("hessian", C.c_float * 36), ("eig", C.c_float * 6), ("P", C.c_float * 36),
        ("is_degenerate", C.c_int), ("iterations", C.c_int), ("n_corr_edge", C.c_int), ("n_corr_plane", C.c_int),
        ("logdet_rot", C.c_float), ("logdet_trans", C.c_float), ("pass_dopt", C.c_int), ("status", C.c_int),
        ("cov", C.c_double * 36),
    ]


class BagBatch(C.Structure):
    _fields_ = [("raw", C.c_void_p), ("offsets", C.c_void_p), ("n_scans", C.c_int), ("seeds", C.c_void_p)]


class FeatureCounts(C.Structure):
    _fields_ = [("n_valid", C.c_int), ("n_sharp", C.c_int), ("n_less_sharp", C.c_int), ("n_flat", C.c_int),
                ("n_less_flat", C.c_int)]


class Preint(C.Structure):
    _fields_ = [
        ("dR", C.c_double * 9), ("dP", C.c_double * 3), ("dV", C.c_double * 3),
        ("dR_dbg", C.c_double * 9), ("dP_dba", C.c_double * 9), ("dP_dbg", C.c_double * 9),
        ("dV_dba", C.c_double * 9), ("dV_dbg", C.c_double * 9), ("cov", C.c_double * 225),
        ("dt", C.c_double), ("n_integrated", C.c_int), ("_pad", C.c_int),
    ]


# every symbol include/vlo.h declares: (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "vlo_default_config": (None, [C.POINTER(Config)]),
    "vlo_set_lidar": (C.c_int, [C.POINTER(Config), C.c_char_p]),
    "vlo_create": (C.c_int, [C.POINTER(Config), C.POINTER(_VP)]),
    "vlo_destroy": (None, [_VP]),
    "vlo_last_error": (C.c_char_p, [_VP]),
    "vlo_version": (C.c_char_p, []),
    "vlo_synchronize": (C.c_int, [_VP]),
    "vlo_launch_count": (C.c_longlong, [_VP]),
    "vlo_set_profiling": (C.c_int, [_VP, C.c_int]),
    "vlo_stage_count": (C.c_int, []),
    "vlo_stage_name": (C.c_char_p, [C.c_int]),
    "vlo_get_stage_times": (C.c_int, [_VP, _VP, _VP]),
    "vlo_stream": (C.c_void_p, [_VP]),
    "vlo_scans_upload": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int]),
    "vlo_scans_upload_pc2": (C.c_int, [_VP, _VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vlo_scans_organise": (C.c_int, [_VP]),
    "vlo_scans_extract": (C.c_int, [_VP]),
    "vlo_scans_counts": (C.c_int, [_VP, C.POINTER(FeatureCounts)]),
    "vlo_scan_get_cloud": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP]),
    "vlo_scan_get_features": (C.c_int, [_VP, C.c_int] + [_VP] * 9),
    "vlo_register_pairs": (C.c_int, [_VP, _VP, _VP, C.c_int, _VP, _VP, _VP]),
    "vlo_set_trace": (C.c_int, [_VP, C.c_int]),
    "vlo_pair_get_correspondences": (C.c_int, [_VP, C.c_int, C.c_int, _VP, _VP]),
    "vlo_map_build": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_int, C.c_int]),
    "vlo_register_map": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    "vlo_register_map_enqueue": (C.c_int, [_VP, _VP, C.c_int, _VP, _VP]),
    "vlo_register_pairs_enqueue": (C.c_int, [_VP, _VP, _VP, C.c_int, _VP, _VP]),
    "vlo_results_finish": (C.c_int, [_VP, _VP, C.c_int]),
    "vlo_map_get_correspondences": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "vlo_map_knn": (C.c_int, [_VP, C.c_int, _VP, C.c_int, C.c_int, C.c_float, _VP, _VP]),
    "vlo_map_reset": (C.c_int, [_VP]),
    "vlo_map_insert": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_int, _VP]),
    "vlo_map_process": (C.c_int, [_VP, C.c_int, _VP, C.POINTER(Result), _VP]),
    "vlo_map_size": (C.c_int, [_VP, _VP, _VP]),
    "vlo_map_get_points": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "vlo_scan_get_stack": (C.c_int, [_VP, C.c_int, _VP, _VP, _VP, _VP]),
    "vlo_scans_stack_counts": (C.c_int, [_VP, _VP, _VP]),
    "vlo_online_reset": (C.c_int, [_VP]),
    "vlo_online_pose": (C.c_int, [_VP, _VP, _VP]),
    "vlo_online_set_map_pose": (C.c_int, [_VP, _VP]),
    "vlo_process_scan": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_double, C.POINTER(Result), C.POINTER(Result)]),
    "vlo_process_scan_pc2": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(Result), C.POINTER(Result)]),
    "vlo_bag_register_map": (C.c_int, [_VP, C.POINTER(BagBatch), C.c_int, C.c_int, _VP]),
    "vlo_bag_register_pairs": (C.c_int, [_VP, C.POINTER(BagBatch), C.c_int, C.c_int, _VP]),
    "vlo_imu_preintegrate_batch": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, _VP, _VP, _VP, C.c_int, _VP]),
    "vlo_pose_diff": (None, [_VP, _VP, _VP]),
    "vlo_dopt_gate": (C.c_int, [_VP, C.c_double, C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vlo_accumulate_pose": (None, [_VP, _VP, C.c_float, _VP]),
}

_lib = None


def load(path: str | None = None):
    """Load libvlo.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "libvlo.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU fallback" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        if os.environ.get("VLO_LIB_PATH") and not hasattr(lib, name):
            continue                     # an older build loaded for a tuning comparison: it may predate a symbol
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
