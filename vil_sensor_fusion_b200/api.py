"""Python host side of the C-ABI (ctypes): the handle the package's tooling uses.

Mirrors the reference-facing surface: scans in (PointCloud2 payloads as float32 arrays), per-scan
odometry + OptStatus-like results out, `(6,6,T)` Hessian stacks for the analysis tooling
(vil_fusion/python/make_prettier_graphs.py:411-474,547-576 in the reference).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import BagBatch, Config, FeatureCounts, Preint, Result

LIDAR = {"VLP-16": (-15.0, 15.0, 16), "HDL-32": (-30.67, 10.67, 32), "HDL-64E": (-24.9, 2.0, 64)}

RESULT_DTYPE = np.dtype([
    ("transform", "f4", 6), ("hessian", "f4", (6, 6)), ("eig", "f4", 6), ("P", "f4", (6, 6)),
    ("is_degenerate", "i4"), ("iterations", "i4"), ("n_corr_edge", "i4"), ("n_corr_plane", "i4"),
    ("logdet_rot", "f4"), ("logdet_trans", "f4"), ("pass_dopt", "i4"), ("status", "i4"), ("cov", "f8", (6, 6)),
])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)

PREINT_DTYPE = np.dtype([
    ("dR", "f8", (3, 3)), ("dP", "f8", 3), ("dV", "f8", 3), ("dR_dbg", "f8", (3, 3)), ("dP_dba", "f8", (3, 3)),
    ("dP_dbg", "f8", (3, 3)), ("dV_dba", "f8", (3, 3)), ("dV_dbg", "f8", (3, 3)), ("cov", "f8", (15, 15)),
    ("dt", "f8"), ("n_integrated", "i4"), ("_pad", "i4"),
])
assert PREINT_DTYPE.itemsize == C.sizeof(Preint)


class VloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("vlo error %d: %s" % (code, msg))
        self.code = code


def default_config(lidar: str = "VLP-16", **kw) -> Config:
    lib = _lib.load()
    c = Config()
    lib.vlo_default_config(C.byref(c))
    if lib.vlo_set_lidar(C.byref(c), lidar.encode()) != 0:
        raise ValueError("unknown lidar preset %r" % lidar)
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        if k == "input_rotation":
            v = (C.c_float * 3)(*v)
        setattr(c, k, v)
    return c


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


class Handle:
    """One vlo_handle: one CUDA stream + all device memory for a batch of scans."""

    def __init__(self, cfg: Config):
        self.lib = _lib.load()
        self.cfg = cfg
        self._h = C.c_void_p()
        rc = self.lib.vlo_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise VloError(rc, {-4: "no usable CUDA device (there is no CPU fallback)", -1: "invalid config"}.get(rc, "vlo_create failed"))
        self.n_scans = 0

    def close(self):
        if self._h:
            self.lib.vlo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc < 0:
            raise VloError(rc, self.lib.vlo_last_error(self._h).decode())
        return rc

    # ---- scans
    def upload(self, scans, stride: int | None = None):
        """scans: list of float32 (n_i, stride) arrays (host) -> resident batch."""
        if isinstance(scans, np.ndarray) and scans.ndim == 2:
            scans = [scans]
        scans = [np.ascontiguousarray(s, np.float32) for s in scans]
        stride = stride or scans[0].shape[1]
        offs = np.zeros(len(scans) + 1, np.int32)
        offs[1:] = np.cumsum([s.shape[0] for s in scans])
        raw = np.concatenate(scans, axis=0) if len(scans) > 1 else scans[0]
        if raw.shape[0] == 0:
            raw = np.zeros((1, stride), np.float32)
        self._check(self.lib.vlo_scans_upload(self._h, _ptr(raw), _ptr(offs), len(scans), stride, 0))
        self.n_scans = len(scans)

    def upload_raw(self, raw_ptr, offsets: np.ndarray, stride: int, on_device: bool):
        offsets = np.ascontiguousarray(offsets, np.int32)
        self._check(self.lib.vlo_scans_upload(self._h, _ptr(raw_ptr), _ptr(offsets), len(offsets) - 1, stride, int(on_device)))
        self.n_scans = len(offsets) - 1

    def upload_pointcloud2(self, msgs):
        """msgs: list of dicts shaped like sensor_msgs/PointCloud2 -- {data: bytes / uint8 array, point_step: int,
        fields: {name: byte offset}} (FLOAT32 x, y, z).  All messages of one call share a layout."""
        step = int(msgs[0]["point_step"])
        f = msgs[0]["fields"]
        blobs = [np.frombuffer(m["data"], np.uint8) if not isinstance(m["data"], np.ndarray) else m["data"].view(np.uint8).ravel() for m in msgs]
        offs = np.zeros(len(msgs) + 1, np.int32)
        offs[1:] = np.cumsum([b.size // step for b in blobs])
        data = np.ascontiguousarray(np.concatenate(blobs))
        self._check(self.lib.vlo_scans_upload_pc2(self._h, _ptr(data), _ptr(offs), len(msgs), step, int(f["x"]), int(f["y"]), int(f["z"]), 0))
        self.n_scans = len(msgs)

    def organise(self):
        self._check(self.lib.vlo_scans_organise(self._h))

    def extract(self):
        self._check(self.lib.vlo_scans_extract(self._h))

    def synchronize(self):
        self._check(self.lib.vlo_synchronize(self._h))

    def set_profiling(self, enable: bool = True):
        self._check(self.lib.vlo_set_profiling(self._h, int(enable)))

    def stage_times(self):
        """{stage name: (accumulated ms, launch groups)} since the last call."""
        n = self.lib.vlo_stage_count()
        ms = np.zeros(n, np.float32)
        cnt = np.zeros(n, np.int32)
        self._check(self.lib.vlo_get_stage_times(self._h, _ptr(ms), _ptr(cnt)))
        return {self.lib.vlo_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}

    def stream_ptr(self) -> int:
        return int(self.lib.vlo_stream(self._h) or 0)

    def launch_count(self) -> int:
        return int(self.lib.vlo_launch_count(self._h))

    def counts(self):
        arr = (FeatureCounts * max(self.n_scans, self.cfg.max_scans))()     # the library writes one record per resident scan
        self._check(self.lib.vlo_scans_counts(self._h, arr))
        return [dict(n_valid=a.n_valid, n_sharp=a.n_sharp, n_less_sharp=a.n_less_sharp, n_flat=a.n_flat,
                     n_less_flat=a.n_less_flat) for a in arr[:self.n_scans]]

    def get_cloud(self, scan: int):
        c = self.counts()[scan]
        n = c["n_valid"]
        cloud = np.zeros((max(n, 1), 4), np.float32)
        rs = np.zeros(self.cfg.n_rings + 1, np.int32)
        src = np.zeros(max(n, 1), np.int32)
        self._check(self.lib.vlo_scan_get_cloud(self._h, scan, _ptr(cloud), _ptr(rs), _ptr(src)))
        return cloud[:n], rs, src[:n]

    def get_features(self, scan: int):
        c = self.counts()[scan]
        n = c["n_valid"]
        label = np.zeros(max(n, 1), np.int8)
        curv = np.zeros(max(n, 1), np.float32)
        picked = np.zeros(max(n, 1), np.uint8)
        sharp = np.zeros(max(c["n_sharp"], 1), np.int32)
        lsharp = np.zeros(max(c["n_less_sharp"], 1), np.int32)
        flat = np.zeros(max(c["n_flat"], 1), np.int32)
        lflat = np.zeros((max(c["n_less_flat"], 1), 4), np.float32)
        lsr = np.zeros(self.cfg.n_rings + 1, np.int32)
        lfr = np.zeros(self.cfg.n_rings + 1, np.int32)
        self._check(self.lib.vlo_scan_get_features(self._h, scan, _ptr(label), _ptr(curv), _ptr(picked), _ptr(sharp),
                                                   _ptr(lsharp), _ptr(flat), _ptr(lflat), _ptr(lsr), _ptr(lfr)))
        return dict(label=label[:n], curvature=curv[:n], picked=picked[:n], sharp_idx=sharp[:c["n_sharp"]],
                    less_sharp_idx=lsharp[:c["n_less_sharp"]], flat_idx=flat[:c["n_flat"]],
                    less_flat=lflat[:c["n_less_flat"]], less_sharp_ring_start=lsr, less_flat_ring_start=lfr)

    # ---- scan-to-scan
    def register_pairs(self, last, cur, seeds=None, last_transforms=None) -> np.ndarray:
        last = np.ascontiguousarray(last, np.int32)
        cur = np.ascontiguousarray(cur, np.int32)
        n = len(last)
        seeds = None if seeds is None else np.ascontiguousarray(seeds, np.float32).reshape(n, 6)
        lt = None if last_transforms is None else np.ascontiguousarray(last_transforms, np.float32).reshape(n, 6)
        out = np.zeros(n, RESULT_DTYPE)
        self._check(self.lib.vlo_register_pairs(self._h, _ptr(last), _ptr(cur), n, _ptr(seeds), _ptr(lt), _ptr(out)))
        return out

    def pair_correspondences(self, pair: int, rnd: int, n_sharp: int, n_flat: int):
        ci = np.zeros((max(n_sharp, 1), 2), np.int32)
        si = np.zeros((max(n_flat, 1), 3), np.int32)
        self._check(self.lib.vlo_pair_get_correspondences(self._h, pair, rnd, _ptr(ci), _ptr(si)))
        return ci[:n_sharp], si[:n_flat]

    # ---- scan-to-map
    def map_build(self, corner_xyzi, surf_xyzi):
        c = np.ascontiguousarray(corner_xyzi, np.float32)
        s = np.ascontiguousarray(surf_xyzi, np.float32)
        self._check(self.lib.vlo_map_build(self._h, _ptr(c), c.shape[0], _ptr(s), s.shape[0], 0))

    def register_map(self, scans, seeds) -> np.ndarray:
        scans = np.ascontiguousarray(scans, np.int32)
        seeds = np.ascontiguousarray(seeds, np.float32).reshape(len(scans), 6)
        out = np.zeros(len(scans), RESULT_DTYPE)
        self._check(self.lib.vlo_register_map(self._h, _ptr(scans), len(scans), _ptr(seeds), _ptr(out)))
        return out

    def register_map_enqueue(self, scans, seeds, out_address: int) -> None:
        """vlo_register_map without the host synchronisation: the result records land at `out_address` (pinned host memory,
        len(scans) * RESULT_DTYPE.itemsize bytes) once the handle's stream reaches them; see results_finish."""
        scans = np.ascontiguousarray(scans, np.int32)
        seeds = np.ascontiguousarray(seeds, np.float32).reshape(len(scans), 6)
        self._check(self.lib.vlo_register_map_enqueue(self._h, _ptr(scans), len(scans), _ptr(seeds), C.c_void_p(out_address)))

    def register_pairs_enqueue(self, last, cur, out_address: int, seeds=None) -> None:
        """vlo_register_pairs without the host synchronisation (rigid batches); see register_map_enqueue / results_finish"""
        last = np.ascontiguousarray(last, np.int32)
        cur = np.ascontiguousarray(cur, np.int32)
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, np.float32).reshape(len(last), 6)
        self._check(self.lib.vlo_register_pairs_enqueue(self._h, _ptr(last), _ptr(cur), len(last), _ptr(seeds), C.c_void_p(out_address)))

    def results_finish(self, records: np.ndarray) -> np.ndarray:
        """after synchronize(): completes records written by register_map_enqueue (in place) and returns them"""
        assert records.dtype == RESULT_DTYPE and records.flags["C_CONTIGUOUS"]
        rc = self.lib.vlo_results_finish(self._h, _ptr(records), len(records))
        if rc < 0:
            self._check(rc)
        return records

    def map_correspondences(self, slot: int, n_corner: int, n_surf: int):
        ci = np.zeros((max(n_corner, 1), 5), np.int32)
        si = np.zeros((max(n_surf, 1), 5), np.int32)
        self._check(self.lib.vlo_map_get_correspondences(self._h, slot, _ptr(ci), _ptr(si)))
        return ci[:n_corner], si[:n_surf]

    def map_knn(self, which: int, queries, k: int, max_d2: float = 25.0):
        q = np.ascontiguousarray(queries, np.float32)
        idx = np.zeros((q.shape[0], k), np.int32)
        d2 = np.zeros((q.shape[0], k), np.float32)
        self._check(self.lib.vlo_map_knn(self._h, which, _ptr(q), q.shape[0], k, max_d2, _ptr(idx), _ptr(d2)))
        return idx, d2

    # ---- maintained map (LaserMapping's map side)
    def map_reset(self):
        self._check(self.lib.vlo_map_reset(self._h))

    def map_insert(self, corner_xyzi, surf_xyzi, pose6):
        c = np.ascontiguousarray(corner_xyzi, np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(surf_xyzi, np.float32).reshape(-1, 4)
        self._check(self.lib.vlo_map_insert(self._h, _ptr(c), c.shape[0], _ptr(s), s.shape[0],
                                            _ptr(np.ascontiguousarray(pose6, np.float32))))

    def map_process(self, scan: int, seed6):
        """One LaserMapping tick for a resident scan; returns (result record, info dict)."""
        out = Result()
        info = np.zeros(6, np.int32)
        self._check(self.lib.vlo_map_process(self._h, scan, _ptr(np.ascontiguousarray(seed6, np.float32)), C.byref(out), _ptr(info)))
        r = np.frombuffer(bytes(out), RESULT_DTYPE)[0]
        return r, dict(n_ds=(int(info[0]), int(info[1])), n_sub=(int(info[2]), int(info[3])), n_map=(int(info[4]), int(info[5])))

    def map_size(self):
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.vlo_map_size(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def map_points(self, which: int):
        n = self.map_size()[which]
        pts = np.zeros((max(n, 1), 4), np.float32)
        cube = np.zeros(max(n, 1), np.int32)
        self._check(self.lib.vlo_map_get_points(self._h, which, _ptr(pts), _ptr(cube)))
        return pts[:n], cube[:n]

    def stack_counts(self, n_scans: int):
        nc = np.zeros(n_scans, np.int32)
        ns = np.zeros(n_scans, np.int32)
        self._check(self.lib.vlo_scans_stack_counts(self._h, _ptr(nc), _ptr(ns)))
        return nc, ns

    def get_stack(self, scan: int):
        """Down-sampled corner / surface stacks of a resident scan (the scan-to-map query clouds)."""
        nc, ns = C.c_int(0), C.c_int(0)
        self._check(self.lib.vlo_scan_get_stack(self._h, scan, None, C.byref(nc), None, C.byref(ns)))
        c = np.zeros((max(nc.value, 1), 4), np.float32)
        s = np.zeros((max(ns.value, 1), 4), np.float32)
        self._check(self.lib.vlo_scan_get_stack(self._h, scan, _ptr(c), None, _ptr(s), None))
        return c[:nc.value], s[:ns.value]

    def online_pose(self):
        s = np.zeros(6, np.float32)
        m = np.zeros(6, np.float32)
        self._check(self.lib.vlo_online_pose(self._h, _ptr(s), _ptr(m)))
        return s, m

    def online_set_map_pose(self, pose6):
        self._check(self.lib.vlo_online_set_map_pose(self._h, _ptr(np.ascontiguousarray(pose6, np.float32))))

    # ---- online
    def process_scan(self, raw: np.ndarray, stamp: float = 0.0, want_map: bool = False):
        raw = np.ascontiguousarray(raw, np.float32)
        odom = Result()
        mapped = Result() if want_map else None
        rc = self._check(self.lib.vlo_process_scan(self._h, _ptr(raw), raw.shape[0], raw.shape[1], stamp,
                                                   C.byref(odom), C.byref(mapped) if want_map else None))
        o = np.frombuffer(bytes(odom), RESULT_DTYPE)[0]
        m = np.frombuffer(bytes(mapped), RESULT_DTYPE)[0] if want_map else None
        return rc, o, m

    def process_pointcloud2(self, msg, stamp: float = 0.0, want_map: bool = False):
        """One online tick for a PointCloud2-shaped dict {data, point_step, fields: {name: byte offset}} (see upload_pointcloud2)."""
        data = np.frombuffer(msg["data"], np.uint8) if not isinstance(msg["data"], np.ndarray) else msg["data"].view(np.uint8).ravel()
        data = np.ascontiguousarray(data)
        step, f = int(msg["point_step"]), msg["fields"]
        odom = Result()
        mapped = Result() if want_map else None
        rc = self._check(self.lib.vlo_process_scan_pc2(self._h, _ptr(data), data.size // step, step, int(f["x"]), int(f["y"]), int(f["z"]), stamp,
                                                       C.byref(odom), C.byref(mapped) if want_map else None))
        o = np.frombuffer(bytes(odom), RESULT_DTYPE)[0]
        m = np.frombuffer(bytes(mapped), RESULT_DTYPE)[0] if want_map else None
        return rc, o, m

    # ---- whole-bag streaming (copy of batch k+1 overlaps the kernels of batch k)
    def _bag(self, fn, batches, stride, pairs: bool) -> np.ndarray:
        """batches: list of (raw, offsets, seeds); raw = host float32 array or an int address of pinned memory."""
        arr = (BagBatch * len(batches))()
        keep = []
        n_out = 0
        for i, (raw, offsets, seeds) in enumerate(batches):
            offsets = np.ascontiguousarray(offsets, np.int32)
            n = len(offsets) - 1
            if not isinstance(raw, int):
                raw = np.ascontiguousarray(raw, np.float32)
            if seeds is not None:
                seeds = np.ascontiguousarray(seeds, np.float32).reshape(n - 1 if pairs else n, 6)
            keep.append((raw, offsets, seeds))
            arr[i].raw = raw if isinstance(raw, int) else raw.ctypes.data
            arr[i].offsets = offsets.ctypes.data
            arr[i].n_scans = n
            arr[i].seeds = None if seeds is None else seeds.ctypes.data
            n_out += (n - 1) if pairs else n
        out = np.zeros(max(n_out, 1), RESULT_DTYPE)
        self._check(fn(self._h, arr, len(batches), stride, _ptr(out)))
        self.n_scans = 2 * (self.cfg.max_scans // 2)
        return out[:n_out]

    def bag_register_map(self, batches, stride: int = 4) -> np.ndarray:
        return self._bag(self.lib.vlo_bag_register_map, batches, stride, False)

    def bag_register_pairs(self, batches, stride: int = 4) -> np.ndarray:
        return self._bag(self.lib.vlo_bag_register_pairs, batches, stride, True)

    # ---- IMU
    def imu_preintegrate_batch(self, t, acc, gyro, t0, t1, bias=None) -> np.ndarray:
        t = np.ascontiguousarray(t, np.float64)
        acc = np.ascontiguousarray(acc, np.float64)
        gyro = np.ascontiguousarray(gyro, np.float64)
        t0 = np.ascontiguousarray(t0, np.float64)
        t1 = np.ascontiguousarray(t1, np.float64)
        bias = np.zeros(6) if bias is None else np.ascontiguousarray(bias, np.float64)
        out = np.zeros(len(t0), PREINT_DTYPE)
        self._check(self.lib.vlo_imu_preintegrate_batch(self._h, _ptr(t), _ptr(acc), _ptr(gyro), len(t), _ptr(t0),
                                                        _ptr(t1), _ptr(bias), len(t0), _ptr(out)))
        return out


def pose_diff(before7, after7):
    out = np.zeros(7)
    _lib.load().vlo_pose_diff(_ptr(np.ascontiguousarray(before7, np.float64)),
                              _ptr(np.ascontiguousarray(after7, np.float64)), _ptr(out))
    return out


def dopt_gate(hessian, rot_thr=11.5, trans_thr=28.9):
    Hm = np.ascontiguousarray(hessian, np.float32)
    lr, lt = C.c_float(), C.c_float()
    ok = _lib.load().vlo_dopt_gate(_ptr(Hm), rot_thr, trans_thr, C.byref(lr), C.byref(lt))
    return bool(ok), lr.value, lt.value


def accumulate_pose(sum_in, transform, fudge=1.0):
    out = np.zeros(6, np.float32)
    _lib.load().vlo_accumulate_pose(_ptr(np.ascontiguousarray(sum_in, np.float32)),
                                    _ptr(np.ascontiguousarray(transform, np.float32)), fudge, _ptr(out))
    return out


def hessian_stack(results: np.ndarray) -> np.ndarray:
    """(6,6,T) float64 array in the layout `apply_degen_function` consumes
    (vil_fusion/python/make_prettier_graphs.py:547-576)."""
    return np.ascontiguousarray(np.transpose(results["hessian"].astype(np.float64), (1, 2, 0)))
