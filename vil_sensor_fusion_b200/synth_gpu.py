"""Device-side scan synthesis (libvlo_synth.so, csrc/synth_scan.cu) -- a BENCH / TEST TOOL, not part of the hot path.

Frame k of a sweep sequence is generated in device memory from (seed, k): whole-bag runs then have no
host->device transfer in their timed region.  Same sensor model and scene description as synth.py (numpy), float32.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import synth

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libvlo_synth.so")


class Scene(C.Structure):
    _fields_ = [("n_planes", C.c_int), ("n_cyl", C.c_int), ("n_boxes", C.c_int),
                ("planes", (C.c_float * 4) * 8), ("cyl", (C.c_float * 3) * 16), ("boxes", (C.c_float * 6) * 8)]


class Sensor(C.Structure):
    _fields_ = [("rings", C.c_int), ("n_az", C.c_int), ("lower_deg", C.c_float), ("upper_deg", C.c_float),
                ("max_range", C.c_float), ("noise_sigma", C.c_float), ("scan_period", C.c_float), ("rolling", C.c_int),
                ("loop_a", C.c_float), ("loop_b", C.c_float), ("loop_period", C.c_float), ("yaw_wobble", C.c_float), ("bob_amp", C.c_float)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libvlo_synth.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        lib.vlo_synth_scans.restype = C.c_int
        lib.vlo_synth_scans.argtypes = [C.POINTER(Scene), C.POINTER(Sensor), C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_void_p]
        lib.vlo_synth_poses.restype = C.c_int
        lib.vlo_synth_poses.argtypes = [C.POINTER(Sensor), C.c_int, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def make_scene(scene: synth.Scene) -> Scene:
    s = Scene()
    s.n_planes, s.n_cyl, s.n_boxes = len(scene.planes), len(scene.cylinders), len(scene.boxes)
    assert s.n_planes <= 8 and s.n_cyl <= 16 and s.n_boxes <= 8
    for i, pl in enumerate(scene.planes):
        for j in range(4):
            s.planes[i][j] = float(pl[j])
    for i, cy in enumerate(scene.cylinders):
        for j in range(3):
            s.cyl[i][j] = float(cy[j])
    for i, bx in enumerate(scene.boxes):
        for j in range(6):
            s.boxes[i][j] = float(bx[j])
    return s


def make_sensor(model: str = "HDL-64E", n_az: int = 1800, noise_sigma: float = 0.01, rolling: bool = False, scan_period: float = 0.1,
                loop_a: float = 10.0, loop_b: float = 6.0, speed: float = 1.5) -> Sensor:
    lo, hi, rings = synth.LIDAR_MODELS[model]
    # ellipse perimeter (Ramanujan) / speed = seconds per lap
    h = ((loop_a - loop_b) / (loop_a + loop_b)) ** 2
    per = np.pi * (loop_a + loop_b) * (1 + 3 * h / (10 + np.sqrt(4 - 3 * h)))
    return Sensor(rings, n_az, lo, hi, 120.0, noise_sigma, scan_period, int(rolling), loop_a, loop_b, float(per / speed), 0.05, 0.05)


def synth_scans(scene: Scene, sensor: Sensor, frame_first: int, n_frames: int, seed: int, d_out_ptr: int, stream_ptr: int = 0) -> None:
    """Sweeps [frame_first, frame_first + n_frames) into device memory at d_out_ptr (n_frames x rings*n_az x 4 float32)."""
    rc = load().vlo_synth_scans(C.byref(scene), C.byref(sensor), frame_first, n_frames, seed, C.c_void_p(d_out_ptr), C.c_void_p(stream_ptr))
    if rc != 0:
        raise RuntimeError("vlo_synth_scans failed: %d" % rc)


def poses(sensor: Sensor, frame_first: int, n_frames: int):
    """Ground-truth sensor poses of the frames (ROS frame): R (n, 3, 3), p (n, 3)."""
    out = np.zeros((n_frames, 12), np.float32)
    rc = load().vlo_synth_poses(C.byref(sensor), frame_first, n_frames, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError("vlo_synth_poses failed: %d" % rc)
    return out[:, :9].reshape(-1, 3, 3).astype(np.float64), out[:, 9:].astype(np.float64)
