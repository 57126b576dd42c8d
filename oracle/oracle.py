"""TEST INFRASTRUCTURE: ctypes loader for the CPU oracle (oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  See oracle/vlo_oracle.h for what is pinned by the reference and what is not.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class Config(C.Structure):
    _fields_ = [
        ("scan_period", C.c_float), ("n_rings", C.c_int), ("lower_deg", C.c_float), ("upper_deg", C.c_float),
        ("feature_regions", C.c_int), ("curvature_region", C.c_int), ("max_corner_sharp", C.c_int),
        ("max_corner_less_sharp", C.c_int), ("max_surface_flat", C.c_int),
        ("surface_curvature_threshold", C.c_float), ("less_flat_filter_size", C.c_float),
        ("odom_max_iterations", C.c_int), ("odom_delta_t_abort", C.c_float), ("odom_delta_r_abort", C.c_float),
        ("odom_degen_eig", C.c_float), ("map_max_iterations", C.c_int), ("map_delta_t_abort", C.c_float),
        ("map_delta_r_abort", C.c_float), ("map_degen_eig", C.c_float), ("deskew", C.c_int),
        ("odom_forward_bound_quirk", C.c_int), ("dopt_rot_threshold", C.c_float), ("dopt_trans_threshold", C.c_float),
        ("corner_filter_size", C.c_float), ("surface_filter_size", C.c_float), ("map_cube_size", C.c_float),
        ("map_dims", C.c_int * 3), ("map_start_cubes", C.c_int * 3), ("n_neighbor_cubes", C.c_int), ("io_ratio", C.c_int),
        ("rotate_input", C.c_int), ("input_rotation", C.c_float * 3), ("ring_field", C.c_int), ("ring_field_type", C.c_int),
    ]


class FeatureCounts(C.Structure):
    _fields_ = [("n_sharp", C.c_int), ("n_less_sharp", C.c_int), ("n_flat", C.c_int), ("n_less_flat", C.c_int)]


class RegResult(C.Structure):
    _fields_ = [
        ("transform", C.c_float * 6), ("hessian", C.c_float * 36), ("eig", C.c_float * 6), ("P", C.c_float * 36),
        ("is_degenerate", C.c_int), ("iterations", C.c_int), ("n_corr_edge", C.c_int), ("n_corr_plane", C.c_int),
        ("logdet_rot", C.c_float), ("logdet_trans", C.c_float), ("pass_dopt", C.c_int),
        ("cov", C.c_double * 36), ("status", C.c_int),
    ]


class ImuParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("cov_accel", "cov_gyro", "cov_integration", "cov_bias_acc",
                                          "cov_bias_omega", "cov_bias_acc_omega_int")]


class Preint(C.Structure):
    _fields_ = [
        ("dR", C.c_double * 9), ("dP", C.c_double * 3), ("dV", C.c_double * 3),
        ("dR_dbg", C.c_double * 9), ("dP_dba", C.c_double * 9), ("dP_dbg", C.c_double * 9),
        ("dV_dba", C.c_double * 9), ("dV_dbg", C.c_double * 9), ("cov", C.c_double * 225),
        ("dt", C.c_double), ("n_integrated", C.c_int),
    ]


PREINT_DTYPE = np.dtype([
    ("dR", "f8", (3, 3)), ("dP", "f8", 3), ("dV", "f8", 3), ("dR_dbg", "f8", (3, 3)), ("dP_dba", "f8", (3, 3)),
    ("dP_dbg", "f8", (3, 3)), ("dV_dba", "f8", (3, 3)), ("dV_dbg", "f8", (3, 3)), ("cov", "f8", (15, 15)),
    ("dt", "f8"), ("n_integrated", "i4"), ("_pad", "i4"),
])
assert PREINT_DTYPE.itemsize == C.sizeof(Preint)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_organise.restype = C.c_int
        _lib.orc_kdtree_build.restype = C.c_void_p
        _lib.orc_degeneracy.restype = C.c_int
        _lib.orc_dopt_gate.restype = C.c_int
        _lib.orc_edge_coeff.restype = C.c_int
        _lib.orc_plane_coeff.restype = C.c_int
        _lib.orc_voxel_downsample.restype = C.c_int
        _lib.orc_lmap_create.restype = C.c_void_p
        _lib.orc_lmap_size.restype = C.c_int
        _lib.orc_lmap_submap.restype = C.c_int
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


LIDAR = {"VLP-16": (-15.0, 15.0, 16), "HDL-32": (-30.67, 10.67, 32), "HDL-64E": (-24.9, 2.0, 64),
         "O1-16": (-16.611, 16.611, 16), "O1-64": (-16.611, 16.611, 64), "Bperl-32": (2.3125, 89.5, 32)}


def default_config(lidar: str = "VLP-16", **kw) -> Config:
    c = Config()
    lib().orc_default_config(C.byref(c))
    lo, hi, r = LIDAR[lidar]
    c.lower_deg, c.upper_deg, c.n_rings = lo, hi, r
    for k, v in kw.items():
        if k == "input_rotation":
            v = (C.c_float * 3)(*v)
        setattr(c, k, v)
    return c


def organise(cfg: Config, raw: np.ndarray):
    raw = np.ascontiguousarray(raw, dtype=np.float32)
    n, stride = raw.shape
    out = np.zeros((max(n, 1), 4), np.float32)
    ring_start = np.zeros(cfg.n_rings + 1, np.int32)
    src = np.zeros(max(n, 1), np.int32)
    m = lib().orc_organise(C.byref(cfg), _p(raw), n, stride, _p(out), _p(ring_start), _p(src))
    return out[:m].copy(), ring_start, src[:m].copy()


def extract(cfg: Config, cloud: np.ndarray, ring_start: np.ndarray):
    cloud = np.ascontiguousarray(cloud, dtype=np.float32)
    ring_start = np.ascontiguousarray(ring_start, dtype=np.int32)
    n = cloud.shape[0]
    label = np.zeros(max(n, 1), np.int8)
    curv = np.zeros(max(n, 1), np.float32)
    picked = np.zeros(max(n, 1), np.uint8)
    sharp = np.zeros(max(n, 1), np.int32)
    lsharp = np.zeros(max(n, 1), np.int32)
    flat = np.zeros(max(n, 1), np.int32)
    lflat = np.zeros((max(n, 1), 4), np.float32)
    lsr = np.zeros(cfg.n_rings + 1, np.int32)
    lfr = np.zeros(cfg.n_rings + 1, np.int32)
    cnt = FeatureCounts()
    lib().orc_extract(C.byref(cfg), _p(cloud), _p(ring_start), _p(label), _p(curv), _p(picked), _p(sharp), _p(lsharp),
                      _p(flat), _p(lflat), _p(lsr), _p(lfr), C.byref(cnt))
    return dict(label=label[:n], curvature=curv[:n], picked=picked[:n], sharp_idx=sharp[:cnt.n_sharp].copy(),
                less_sharp_idx=lsharp[:cnt.n_less_sharp].copy(), flat_idx=flat[:cnt.n_flat].copy(),
                less_flat=lflat[:cnt.n_less_flat].copy(), less_sharp_ring_start=lsr, less_flat_ring_start=lfr)


def knn_brute(cloud, q, k):
    cloud = np.ascontiguousarray(cloud, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    idx = np.zeros((q.shape[0], k), np.int32)
    d2 = np.zeros((q.shape[0], k), np.float32)
    lib().orc_knn_brute(_p(cloud), cloud.shape[0], _p(q), q.shape[0], k, _p(idx), _p(d2))
    return idx, d2


def knn_kdtree(cloud, q, k):
    cloud = np.ascontiguousarray(cloud, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    idx = np.zeros((q.shape[0], k), np.int32)
    d2 = np.zeros((q.shape[0], k), np.float32)
    t = C.c_void_p(lib().orc_kdtree_build(_p(cloud), cloud.shape[0]))
    lib().orc_kdtree_knn(t, _p(q), q.shape[0], k, _p(idx), _p(d2))
    lib().orc_kdtree_free(t)
    return idx, d2


def solve6(A, b):
    A = np.ascontiguousarray(A, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    x = np.zeros(6, np.float32)
    lib().orc_solve6_colpiv_qr(_p(A), _p(b), _p(x))
    return x


def eig6(A):
    A = np.ascontiguousarray(A, np.float32)
    ev = np.zeros(6, np.float32)
    vec = np.zeros((6, 6), np.float32)
    lib().orc_eig6_jacobi(_p(A), _p(ev), _p(vec))
    return ev, vec


def degeneracy(A, thr):
    A = np.ascontiguousarray(A, np.float32)
    ev = np.zeros(6, np.float32)
    P = np.zeros((6, 6), np.float32)
    flag = lib().orc_degeneracy(_p(A), C.c_float(thr), _p(ev), _p(P))
    return bool(flag), ev, P


def dopt_gate(H, rot_thr=11.5, trans_thr=28.9):
    H = np.ascontiguousarray(H, np.float32)
    lr, lt = C.c_float(), C.c_float()
    ok = lib().orc_dopt_gate(_p(H), C.c_double(rot_thr), C.c_double(trans_thr), C.byref(lr), C.byref(lt))
    return bool(ok), lr.value, lt.value


def _result_dict(r: RegResult):
    return dict(transform=np.array(r.transform, np.float32), hessian=np.array(r.hessian, np.float32).reshape(6, 6),
                eig=np.array(r.eig, np.float32), P=np.array(r.P, np.float32).reshape(6, 6),
                is_degenerate=bool(r.is_degenerate), iterations=r.iterations, n_corr_edge=r.n_corr_edge,
                n_corr_plane=r.n_corr_plane, logdet_rot=r.logdet_rot, logdet_trans=r.logdet_trans,
                pass_dopt=bool(r.pass_dopt), cov=np.array(r.cov).reshape(6, 6), status=r.status)


def odometry_register(cfg, cur_sharp, cur_flat, last_corner, lc_ring_start, last_surf, ls_ring_start, seed=None,
                      use_kdtree=True, trace=False):
    cs = np.ascontiguousarray(cur_sharp, np.float32)
    cf = np.ascontiguousarray(cur_flat, np.float32)
    lc = np.ascontiguousarray(last_corner, np.float32)
    ls = np.ascontiguousarray(last_surf, np.float32)
    seed = np.zeros(6, np.float32) if seed is None else np.ascontiguousarray(seed, np.float32)
    res = RegResult()
    n_assoc = (cfg.odom_max_iterations + 4) // 5
    tr_idx = np.full(n_assoc * (2 * cs.shape[0] + 3 * cf.shape[0]) + 1, -2, np.int32) if trace else None
    tr_T = np.zeros((cfg.odom_max_iterations, 6), np.float32) if trace else None
    lib().orc_odometry_register(C.byref(cfg), _p(cs), cs.shape[0], _p(cf), cf.shape[0], _p(lc), lc.shape[0],
                                _p(np.ascontiguousarray(lc_ring_start, np.int32)), _p(ls), ls.shape[0],
                                _p(np.ascontiguousarray(ls_ring_start, np.int32)), _p(seed), int(use_kdtree),
                                C.byref(res), _p(tr_idx), _p(tr_T))
    out = _result_dict(res)
    if trace:
        out["trace_idx"] = tr_idx
        out["trace_T"] = tr_T
    return out


def odometry_associate(cfg, T, cur_sharp, cur_flat, last_corner, last_surf, use_kdtree=False):
    cs = np.ascontiguousarray(cur_sharp, np.float32)
    cf = np.ascontiguousarray(cur_flat, np.float32)
    lc = np.ascontiguousarray(last_corner, np.float32)
    ls = np.ascontiguousarray(last_surf, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    ci = np.zeros((cs.shape[0], 2), np.int32)
    si = np.zeros((cf.shape[0], 3), np.int32)
    lib().orc_odometry_associate(C.byref(cfg), _p(T), _p(cs), cs.shape[0], _p(cf), cf.shape[0], _p(lc), lc.shape[0],
                                 _p(ls), ls.shape[0], int(use_kdtree), _p(ci), _p(si))
    return ci, si


def transform_to_start(cfg, T, pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros_like(pts)
    lib().orc_transform_to_start(C.byref(cfg), _p(np.ascontiguousarray(T, np.float32)), _p(pts), pts.shape[0], _p(out))
    return out


def transform_to_end(cfg, T, pts):
    pts = np.array(pts, np.float32, order="C", copy=True)
    lib().orc_transform_to_end(C.byref(cfg), _p(np.ascontiguousarray(T, np.float32)), _p(pts), pts.shape[0])
    return pts


def accumulate_pose(sum_in, T, fudge=1.0):
    out = np.zeros(6, np.float32)
    lib().orc_accumulate_pose(_p(np.ascontiguousarray(sum_in, np.float32)), _p(np.ascontiguousarray(T, np.float32)),
                              C.c_float(fudge), _p(out))
    return out


def mapping_register(cfg, corner_q, surf_q, corner_map, surf_map, seed, use_kdtree=True, trace=False):
    cq = np.ascontiguousarray(corner_q, np.float32)
    sq = np.ascontiguousarray(surf_q, np.float32)
    cm = np.ascontiguousarray(corner_map, np.float32)
    sm = np.ascontiguousarray(surf_map, np.float32)
    seed = np.ascontiguousarray(seed, np.float32)
    res = RegResult()
    tr_idx = np.full((cq.shape[0] + sq.shape[0]) * 5 + 1, -2, np.int32) if trace else None
    tr_T = np.zeros((cfg.map_max_iterations, 6), np.float32) if trace else None
    lib().orc_mapping_register(C.byref(cfg), _p(cq), cq.shape[0], _p(sq), sq.shape[0], _p(cm), cm.shape[0], _p(sm),
                               sm.shape[0], _p(seed), int(use_kdtree), C.byref(res), _p(tr_idx), _p(tr_T))
    out = _result_dict(res)
    if trace:
        out["trace_idx"] = tr_idx
        out["trace_T"] = tr_T
    return out


def voxel_downsample(pts, leaf):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 4)
    out = np.zeros((max(len(pts), 1), 4), np.float32)
    m = lib().orc_voxel_downsample(_p(pts), len(pts), C.c_float(leaf), _p(out))
    return out[:m].copy()


class LaserMap:
    """BasicLaserMapping's map side (oracle/laser_map.c): cube window, sub-map selection, stack down-sampling,
    registration against the sub-map, insertion + voxel re-filtering."""

    def __init__(self, cfg, cap=1 << 20):
        self.cfg, self.cap = cfg, cap
        self._m = C.c_void_p(lib().orc_lmap_create(C.byref(cfg), cap))

    def close(self):
        if self._m:
            lib().orc_lmap_free(self._m)
            self._m = None

    def __del__(self):
        self.close()

    def size(self, which):
        return lib().orc_lmap_size(self._m, which)

    def points(self, which):
        n = self.size(which)
        pts = np.zeros((max(n, 1), 4), np.float32)
        cube = np.zeros(max(n, 1), np.int32)
        lib().orc_lmap_get(self._m, which, _p(pts), _p(cube))
        return pts[:n].copy(), cube[:n].copy()

    def window(self):
        cen = np.zeros(3, np.int32)
        lib().orc_lmap_window(self._m, _p(cen))
        return cen

    def submap_ids(self, which):
        ids = np.zeros(max(self.size(which), 1), np.int32)
        n = lib().orc_lmap_submap(self._m, which, _p(ids))
        return ids[:n].copy()

    def insert(self, corner, surf, T):
        c = np.ascontiguousarray(corner, np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(surf, np.float32).reshape(-1, 4)
        lib().orc_lmap_insert(self._m, _p(c), len(c), _p(s), len(s), _p(np.ascontiguousarray(T, np.float32)))

    def select(self, T):
        side = 2 * self.cfg.n_neighbor_cubes + 1
        centre = np.zeros(3, np.int32)
        mask = np.zeros(side ** 3, np.uint8)
        lib().orc_lmap_select(self._m, _p(np.ascontiguousarray(T, np.float32)), _p(centre), _p(mask))
        return centre, mask.reshape(side, side, side)

    def process(self, corner_stack, surf_stack, seed):
        c = np.ascontiguousarray(corner_stack, np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(surf_stack, np.float32).reshape(-1, 4)
        res = RegResult()
        info = np.zeros(6, np.int32)
        lib().orc_lmap_process(self._m, _p(c), len(c), _p(s), len(s), _p(np.ascontiguousarray(seed, np.float32)),
                               C.byref(res), _p(info))
        out = _result_dict(res)
        out["info"] = dict(n_ds=(int(info[0]), int(info[1])), n_sub=(int(info[2]), int(info[3])), n_map=(int(info[4]), int(info[5])))
        return out


def imu_params(cov_accel=1e-6, cov_gyro=1e-6, cov_integration=1e-8, cov_bias_acc=1e-4, cov_bias_omega=1e-6,
               cov_bias_acc_omega_int=1e-4) -> ImuParams:
    """Defaults = gtsam_fusion/config/carla/fusion_params.yaml:22-27."""
    return ImuParams(cov_accel, cov_gyro, cov_integration, cov_bias_acc, cov_bias_omega, cov_bias_acc_omega_int)


def imu_get_factor(prm, t, acc, gyro, t0, t1, bias=None):
    t = np.ascontiguousarray(t, np.float64)
    acc = np.ascontiguousarray(acc, np.float64)
    gyro = np.ascontiguousarray(gyro, np.float64)
    bias = np.zeros(6) if bias is None else np.ascontiguousarray(bias, np.float64)
    out = np.zeros(1, PREINT_DTYPE)
    lib().orc_imu_get_factor(C.byref(prm), _p(t), _p(acc), _p(gyro), t.shape[0], C.c_double(t0), C.c_double(t1),
                             _p(bias), _p(out))
    return out[0]


def imu_batch(prm, t, acc, gyro, t0, t1, bias=None, n_threads=1):
    t = np.ascontiguousarray(t, np.float64)
    acc = np.ascontiguousarray(acc, np.float64)
    gyro = np.ascontiguousarray(gyro, np.float64)
    t0 = np.ascontiguousarray(t0, np.float64)
    t1 = np.ascontiguousarray(t1, np.float64)
    bias = np.zeros(6) if bias is None else np.ascontiguousarray(bias, np.float64)
    out = np.zeros(t0.shape[0], PREINT_DTYPE)
    lib().orc_imu_batch(C.byref(prm), _p(t), _p(acc), _p(gyro), t.shape[0], _p(t0), _p(t1), _p(bias), t0.shape[0],
                        _p(out), n_threads)
    return out


def pose_diff(before7, after7):
    out = np.zeros(7)
    lib().orc_pose_diff(_p(np.ascontiguousarray(before7, np.float64)), _p(np.ascontiguousarray(after7, np.float64)),
                        _p(out))
    return out


class CpuMap:
    """Map clouds + kd-trees built once (bench.py's CPU baseline and --impl reference legs)."""

    def __init__(self, cfg, corner_map, surf_map):
        self.cfg = cfg
        self.cm = np.ascontiguousarray(corner_map, np.float32)
        self.sm = np.ascontiguousarray(surf_map, np.float32)
        L = lib()
        L.orc_map_create.restype = C.c_void_p
        self._m = C.c_void_p(L.orc_map_create(C.byref(cfg), _p(self.cm), self.cm.shape[0], _p(self.sm), self.sm.shape[0]))

    def close(self):
        if self._m:
            lib().orc_map_free(self._m)
            self._m = None

    def batch_scan_to_map(self, raws, seeds, n_threads=1):
        """organise + extract + scan-to-map registration (+ D-opt gate) for every frame; frames are
        spread over n_threads threads."""
        raws = [np.ascontiguousarray(r, np.float32) for r in raws]
        stride = raws[0].shape[1]
        offs = np.zeros(len(raws) + 1, np.int32)
        offs[1:] = np.cumsum([r.shape[0] for r in raws])
        raw = np.concatenate(raws, axis=0)
        seeds = np.ascontiguousarray(seeds, np.float32).reshape(len(raws), 6)
        res = (RegResult * len(raws))()
        lib().orc_batch_scan_to_map(self._m, _p(raw), _p(offs), len(raws), stride, _p(seeds), res, int(n_threads))
        return [_result_dict(r) for r in res]
