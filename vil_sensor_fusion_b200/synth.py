"""Seeded synthetic world + sensor model (numpy) used by tests, golden fixtures and bench.

Not part of the hot path: it only manufactures inputs shaped like the reference's
(`sensor_msgs/PointCloud2` `/lidar` at 10 Hz, `sensor_msgs/Imu` at 200 Hz; see
carla_tools/config/sensors.json:91-103 and carla_ros_bridge_settings.yaml:12 in the
reference) because the reference ships no bags (sample_bags/.gitignore:1-3).

Frames: everything here is in the ROS sensor frame (x forward, y left, z up); the LOAM
axis permutation (loam_frame_transform.py:52-90) happens inside the hot path.

Scenes:
  S1 "room"     : 40 x 30 x 6 m box room + 12 pillars (r = 0.3 m) + 6 box obstacles
  S2 "corridor" : infinite corridor along x (walls y=+-1.5, floor z=-1, ceiling z=+1.5)
  S3 "plane"    : single ground plane z = -1.5
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

LIDAR_MODELS = {
    # name: (lower_deg, upper_deg, rings) -- MultiScanMapper presets of the LOAM fork
    "VLP-16": (-15.0, 15.0, 16),
    "HDL-32": (-30.67, 10.67, 32),
    "HDL-64E": (-24.9, 2.0, 64),
    # additional presets named by loam_params.yaml:22 (data-sheet fields of view)
    "O1-16": (-16.611, 16.611, 16),
    "O1-64": (-16.611, 16.611, 64),
    "Bperl-32": (2.3125, 89.5, 32),
}


@dataclasses.dataclass
class Scene:
    """Analytic scene: axis-aligned room planes, vertical cylinders, axis-aligned boxes."""
    planes: np.ndarray      # (P,4) n.x n.y n.z d  with n.p + d = 0, unit n
    cylinders: np.ndarray   # (C,3) cx cy r   (infinite along z)
    boxes: np.ndarray       # (B,6) xmin ymin zmin xmax ymax zmax
    name: str = "scene"


def scene_room(seed: int = 0) -> Scene:
    rng = np.random.default_rng(seed)
    hx, hy, z0, z1 = 20.0, 15.0, -1.5, 4.5
    planes = np.array([
        [1, 0, 0, hx], [-1, 0, 0, hx],
        [0, 1, 0, hy], [0, -1, 0, hy],
        [0, 0, 1, -z0], [0, 0, -1, z1],
    ], dtype=np.float64)
    cyl = []
    for i in range(12):
        ang = 2 * math.pi * i / 12 + rng.uniform(-0.1, 0.1)
        rad = rng.uniform(6.0, 13.0)
        cyl.append([rad * math.cos(ang) * 1.3, rad * math.sin(ang), 0.3])
    boxes = []
    for i in range(6):
        cx, cy = rng.uniform(-16, 16), rng.uniform(-12, 12)
        if abs(cx) < 4 and abs(cy) < 4:
            cx += 8.0
        sx, sy, sz = rng.uniform(0.8, 2.5), rng.uniform(0.8, 2.5), rng.uniform(0.8, 2.5)
        boxes.append([cx - sx / 2, cy - sy / 2, z0, cx + sx / 2, cy + sy / 2, z0 + sz])
    return Scene(planes, np.array(cyl), np.array(boxes), "room")


def scene_corridor() -> Scene:
    planes = np.array([
        [0, 1, 0, 1.5], [0, -1, 0, 1.5],
        [0, 0, 1, 1.0], [0, 0, -1, 1.5],
    ], dtype=np.float64)
    return Scene(planes, np.zeros((0, 3)), np.zeros((0, 6)), "corridor")


def scene_plane() -> Scene:
    planes = np.array([[0, 0, 1, 1.5]], dtype=np.float64)
    return Scene(planes, np.zeros((0, 3)), np.zeros((0, 6)), "plane")


def raycast(scene: Scene, origin: np.ndarray, dirs: np.ndarray, max_range: float = 120.0) -> np.ndarray:
    """Range along each ray (inf where nothing is hit within max_range).

    origin: (N,3) or (3,), dirs: (N,3) unit vectors, world frame.
    """
    dirs = np.asarray(dirs, dtype=np.float64)
    origin = np.broadcast_to(np.asarray(origin, dtype=np.float64), dirs.shape)
    n = dirs.shape[0]
    best = np.full(n, np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for pl in scene.planes:
            nrm, d = pl[:3], pl[3]
            denom = dirs @ nrm
            num = -(origin @ nrm + d)
            t = num / denom
            ok = (denom < 0) & (t > 1e-6)          # only hit the inside face
            best = np.where(ok & (t < best), t, best)
        for cx, cy, r in scene.cylinders:
            ox, oy = origin[:, 0] - cx, origin[:, 1] - cy
            a = dirs[:, 0] ** 2 + dirs[:, 1] ** 2
            b = 2 * (ox * dirs[:, 0] + oy * dirs[:, 1])
            c = ox * ox + oy * oy - r * r
            disc = b * b - 4 * a * c
            sq = np.sqrt(np.maximum(disc, 0))
            t = (-b - sq) / (2 * a)
            ok = (disc > 0) & (t > 1e-6) & (a > 1e-12)
            best = np.where(ok & (t < best), t, best)
        for bx in scene.boxes:
            lo, hi = bx[:3], bx[3:]
            inv = 1.0 / dirs
            t0 = (lo - origin) * inv
            t1 = (hi - origin) * inv
            tmin = np.minimum(t0, t1).max(axis=1)
            tmax = np.maximum(t0, t1).min(axis=1)
            ok = (tmax >= tmin) & (tmin > 1e-6)
            best = np.where(ok & (tmin < best), tmin, best)
    best = np.where(best <= max_range, best, np.inf)
    return best


def rot_zyx(yaw, pitch, roll) -> np.ndarray:
    """R = Rz(yaw) Ry(pitch) Rx(roll); supports array inputs -> (...,3,3)."""
    yaw, pitch, roll = np.broadcast_arrays(np.asarray(yaw, float), np.asarray(pitch, float), np.asarray(roll, float))
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    R = np.empty(yaw.shape + (3, 3))
    R[..., 0, 0] = cy * cp
    R[..., 0, 1] = cy * sp * sr - sy * cr
    R[..., 0, 2] = cy * sp * cr + sy * sr
    R[..., 1, 0] = sy * cp
    R[..., 1, 1] = sy * sp * sr + cy * cr
    R[..., 1, 2] = sy * sp * cr - cy * sr
    R[..., 2, 0] = -sp
    R[..., 2, 1] = cp * sr
    R[..., 2, 2] = cp * cr
    return R


@dataclasses.dataclass
class Trajectory:
    """Smooth analytic trajectory: forward motion, yaw wobble, z bob."""
    speed: float = 1.5
    yaw_amp: float = 0.2
    yaw_freq: float = 0.3
    bob_amp: float = 0.05
    bob_freq: float = 1.1
    x0: float = -12.0
    pitch_amp: float = 0.0
    roll_amp: float = 0.0

    def position(self, t):
        t = np.asarray(t, float)
        # integrate heading-following motion approximately: x along heading cos, y along sin
        yaw = self.yaw(t)
        # closed form is not needed: use straight-line x with lateral sway consistent with yaw
        x = self.x0 + self.speed * t
        y = (self.speed * self.yaw_amp / max(self.yaw_freq, 1e-9)) * (1 - np.cos(self.yaw_freq * t)) * 0.5
        z = self.bob_amp * np.sin(self.bob_freq * t)
        return np.stack([x, y, z], axis=-1)

    def yaw(self, t):
        return self.yaw_amp * np.sin(self.yaw_freq * np.asarray(t, float))

    def pitch(self, t):
        return self.pitch_amp * np.sin(0.7 * np.asarray(t, float))

    def roll(self, t):
        return self.roll_amp * np.sin(0.9 * np.asarray(t, float))

    def rotation(self, t):
        return rot_zyx(self.yaw(t), self.pitch(t), self.roll(t))


def lidar_dirs(model: str, n_az: int = 1800, start_az: float = 0.0):
    """Sensor-frame unit directions, azimuth-major firing order (column k, ring r).

    The head spins clockwise seen from above (azimuth atan2(y,x) decreasing), which
    is what makes LOAM's `ori = -atan2(y,x)` increase with time.
    Returns dirs (n_az, R, 3) and per-column time fraction (n_az,) in [0,1).
    """
    lo, hi, rings = LIDAR_MODELS[model]
    elev = np.deg2rad(np.linspace(lo, hi, rings))
    frac = np.arange(n_az) / n_az
    az = start_az - 2 * math.pi * frac
    ce, se = np.cos(elev), np.sin(elev)
    d = np.empty((n_az, rings, 3))
    d[..., 0] = np.cos(az)[:, None] * ce[None, :]
    d[..., 1] = np.sin(az)[:, None] * ce[None, :]
    d[..., 2] = se[None, :]
    return d, frac


def make_scan(scene: Scene, model: str = "VLP-16", *, t0: float = 0.0, traj: Trajectory | None = None,
              pose: tuple[np.ndarray, np.ndarray] | None = None, scan_period: float = 0.1,
              n_az: int = 1800, rolling: bool = True, noise_sigma: float = 0.0,
              seed: int = 0, max_range: float = 120.0, start_az: float = 0.0,
              stride_floats: int = 4) -> np.ndarray:
    """One PointCloud2-like scan: float32 (N, stride_floats) = x y z intensity[...] in ROS sensor frame.

    rolling=True  : each azimuth column is ray-cast from the pose at its own firing time and
                    expressed in the sensor frame at that instant (what a spinning head reports).
    rolling=False : snapshot from the pose at t0 (or the explicit `pose=(R, p)`).
    Rays without a return inside max_range are omitted (unorganised cloud, as real drivers publish).
    """
    dirs, frac = lidar_dirs(model, n_az, start_az)
    n_col, rings, _ = dirs.shape
    if pose is not None:
        R = np.broadcast_to(np.asarray(pose[0], float), (n_col, 3, 3))
        p = np.broadcast_to(np.asarray(pose[1], float), (n_col, 3))
    else:
        traj = traj or Trajectory()
        tt = t0 + (frac * scan_period if rolling else np.zeros_like(frac))
        R = traj.rotation(tt)
        p = traj.position(tt)
    wdirs = np.einsum("kij,krj->kri", R, dirs)
    org = np.broadcast_to(p[:, None, :], wdirs.shape)
    rng_ = raycast(scene, org.reshape(-1, 3), wdirs.reshape(-1, 3), max_range).reshape(n_col, rings)
    if noise_sigma > 0:
        g = np.random.default_rng(seed)
        rng_ = rng_ + g.normal(0.0, noise_sigma, rng_.shape)
    ok = np.isfinite(rng_)
    pts = dirs * np.where(ok, rng_, 0.0)[..., None]
    out = np.zeros((n_col, rings, stride_floats), dtype=np.float32)
    out[..., :3] = pts.astype(np.float32)
    if stride_floats > 3:
        out[..., 3] = 1.0
    return np.ascontiguousarray(out[ok])


def make_imu(traj: Trajectory, t_begin: float, t_end: float, rate: float = 200.0, *,
             noise_sigma: float = 0.0, seed: int = 0, jitter: float = 0.0):
    """200 Hz IMU (specific force incl. gravity reaction, body rates) from the analytic trajectory.

    Returns t (M,), acc (M,3), gyro (M,3) float64. Gravity n_g = (0,0,-9.81) (MakeSharedU,
    ImuManagerRos.cpp:16) so a static level IMU reads acc = (0,0,+9.81).
    """
    n = int(round((t_end - t_begin) * rate)) + 1
    t = t_begin + np.arange(n) / rate
    g = np.random.default_rng(seed)
    if jitter > 0:
        t = t + g.uniform(-jitter, jitter, n)
    h = 1e-4
    pos = lambda x: traj.position(x)
    acc_w = (pos(t + h) - 2 * pos(t) + pos(t - h)) / (h * h)
    R = traj.rotation(t)
    Rp = traj.rotation(t + h)
    Rm = traj.rotation(t - h)
    dR = (Rp - Rm) / (2 * h)
    W = np.einsum("kji,kjl->kil", R, dR)           # R^T dR = [w]x
    gyro = np.stack([W[:, 2, 1], W[:, 0, 2], W[:, 1, 0]], axis=-1)
    f_w = acc_w - np.array([0.0, 0.0, -9.81])
    acc = np.einsum("kji,kj->ki", R, f_w)
    if noise_sigma > 0:
        acc = acc + g.normal(0, noise_sigma, acc.shape)
        gyro = gyro + g.normal(0, noise_sigma, gyro.shape)
    return t, acc, gyro


def sample_map_points(scene: Scene, n_points: int, seed: int = 1, corner_frac: float = 0.2,
                      extent: float = 60.0):
    """Map clouds for the scan-to-map config (C2): points on scene surfaces ('surf' map) and
    on scene edges ('corner' map), float32 (N,4) in the LOAM map frame (x=left,y=up,z=fwd) with
    intensity 0. Returned as (corner_pts, surf_pts).
    """
    g = np.random.default_rng(seed)
    n_corner = int(n_points * corner_frac)
    n_surf = n_points - n_corner
    surf = []
    # room planes: sample uniformly inside the bounded faces
    bounds = _scene_bounds(scene, extent)
    areas = []
    faces = []
    for pl in scene.planes:
        ax = int(np.argmax(np.abs(pl[:3])))
        others = [a for a in range(3) if a != ax]
        area = (bounds[1][others[0]] - bounds[0][others[0]]) * (bounds[1][others[1]] - bounds[0][others[1]])
        faces.append(("plane", pl, ax, others))
        areas.append(area)
    for c in scene.cylinders:
        faces.append(("cyl", c, None, None))
        areas.append(2 * math.pi * c[2] * (bounds[1][2] - bounds[0][2]))
    for b in scene.boxes:
        sx, sy, sz = b[3] - b[0], b[4] - b[1], b[5] - b[2]
        faces.append(("box", b, None, None))
        areas.append(2 * (sx * sz + sy * sz) + sx * sy)
    areas = np.array(areas)
    counts = g.multinomial(n_surf, areas / areas.sum())
    for (kind, prm, ax, others), cnt in zip(faces, counts):
        if cnt == 0:
            continue
        if kind == "plane":
            p = np.empty((cnt, 3))
            p[:, ax] = -prm[3] * prm[ax]
            for o in others:
                p[:, o] = g.uniform(bounds[0][o], bounds[1][o], cnt)
        elif kind == "cyl":
            th = g.uniform(0, 2 * math.pi, cnt)
            p = np.stack([prm[0] + prm[2] * np.cos(th), prm[1] + prm[2] * np.sin(th),
                          g.uniform(bounds[0][2], bounds[1][2], cnt)], axis=-1)
        else:
            p = _sample_box_surface(g, prm, cnt)
        surf.append(p)
    surf = np.concatenate(surf) if surf else np.zeros((0, 3))
    # edges: room wall/wall + wall/floor intersections, box vertical + top edges
    segs = _scene_edges(scene, bounds)
    lens = np.array([np.linalg.norm(b - a) for a, b in segs])
    counts = g.multinomial(n_corner, lens / lens.sum()) if len(segs) else []
    corner = []
    for (a, b), cnt in zip(segs, counts):
        u = g.uniform(0, 1, cnt)[:, None]
        corner.append(a[None] * (1 - u) + b[None] * u)
    corner = np.concatenate(corner) if corner else np.zeros((0, 3))
    return _ros_to_loam4(corner), _ros_to_loam4(surf)


def _ros_to_loam4(p_ros: np.ndarray) -> np.ndarray:
    out = np.zeros((p_ros.shape[0], 4), dtype=np.float32)
    out[:, 0] = p_ros[:, 1]
    out[:, 1] = p_ros[:, 2]
    out[:, 2] = p_ros[:, 0]
    return out


def _scene_bounds(scene: Scene, extent: float):
    lo = np.array([-extent, -extent, -extent])
    hi = np.array([extent, extent, extent])
    for pl in scene.planes:
        ax = int(np.argmax(np.abs(pl[:3])))
        v = -pl[3] * pl[ax]
        if pl[ax] > 0:
            lo[ax] = max(lo[ax], v)
        else:
            hi[ax] = min(hi[ax], v)
    return lo, hi


def _sample_box_surface(g, b, cnt):
    sx, sy, sz = b[3] - b[0], b[4] - b[1], b[5] - b[2]
    a = np.array([sx * sz, sx * sz, sy * sz, sy * sz, sx * sy])
    which = g.choice(5, size=cnt, p=a / a.sum())
    u, v = g.uniform(0, 1, cnt), g.uniform(0, 1, cnt)
    p = np.empty((cnt, 3))
    for k in range(5):
        m = which == k
        if k == 0:
            p[m] = np.stack([b[0] + u[m] * sx, np.full(m.sum(), b[1]), b[2] + v[m] * sz], -1)
        elif k == 1:
            p[m] = np.stack([b[0] + u[m] * sx, np.full(m.sum(), b[4]), b[2] + v[m] * sz], -1)
        elif k == 2:
            p[m] = np.stack([np.full(m.sum(), b[0]), b[1] + u[m] * sy, b[2] + v[m] * sz], -1)
        elif k == 3:
            p[m] = np.stack([np.full(m.sum(), b[3]), b[1] + u[m] * sy, b[2] + v[m] * sz], -1)
        else:
            p[m] = np.stack([b[0] + u[m] * sx, b[1] + v[m] * sy, np.full(m.sum(), b[5])], -1)
    return p


def _scene_edges(scene: Scene, bounds):
    lo, hi = bounds
    segs = []
    fin = np.isfinite(lo) & np.isfinite(hi)
    # room box edges (all 12 when bounded)
    c = [lo, hi]
    for ax in range(3):
        o = [a for a in range(3) if a != ax]
        for i in range(2):
            for j in range(2):
                a = np.empty(3)
                b = np.empty(3)
                a[ax], b[ax] = lo[ax], hi[ax]
                a[o[0]] = b[o[0]] = c[i][o[0]]
                a[o[1]] = b[o[1]] = c[j][o[1]]
                segs.append((a, b))
    for bx in scene.boxes:
        l, h = bx[:3], bx[3:]
        cc = [l, h]
        for ax in range(3):
            o = [a for a in range(3) if a != ax]
            for i in range(2):
                for j in range(2):
                    if ax != 2 and j == 0 and o[1] == 2:
                        continue  # skip bottom edges lying on the floor
                    a = np.empty(3)
                    b = np.empty(3)
                    a[ax], b[ax] = l[ax], h[ax]
                    a[o[0]] = b[o[0]] = cc[i][o[0]]
                    a[o[1]] = b[o[1]] = cc[j][o[1]]
                    segs.append((a, b))
    return segs


# ---------------------------------------------------------------------------------------------
# Ground-truth helpers in LOAM conventions (axes x=left, y=up, z=forward; R = Ry(ry) Rx(rx) Rz(rz))
_M_ROS2LOAM = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])


def loam_euler_from_R(R: np.ndarray) -> np.ndarray:
    rx = -math.asin(max(-1.0, min(1.0, R[1, 2])))
    ry = math.atan2(R[0, 2], R[2, 2])
    rz = math.atan2(R[1, 0], R[1, 1])
    return np.array([rx, ry, rz])


def loam_R_from_euler(rx, ry, rz) -> np.ndarray:
    sx, cx, sy, cy, sz, cz = math.sin(rx), math.cos(rx), math.sin(ry), math.cos(ry), math.sin(rz), math.cos(rz)
    return np.array([
        [cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx],
        [cx * sz, cx * cz, -sx],
        [-sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx],
    ])


def loam_sweep_transform(Ra, pa, Rb, pb) -> np.ndarray:
    """LOAM's `_transform` (rx,ry,rz,tx,ty,tz) for a sweep that starts at world pose (Ra,pa) and
    ends at (Rb,pb) (ROS axes): p_end = R_T p_start + t."""
    Rt_ros = Rb.T @ Ra
    t_ros = Rb.T @ (pa - pb)
    R = _M_ROS2LOAM @ Rt_ros @ _M_ROS2LOAM.T
    t = _M_ROS2LOAM @ t_ros
    return np.concatenate([loam_euler_from_R(R), t])


def loam_map_pose(R_ros, p_ros) -> np.ndarray:
    """LOAM's `transformTobeMapped` (rx,ry,rz,tx,ty,tz): p_map = R p_sensor + t, LOAM axes."""
    R = _M_ROS2LOAM @ R_ros @ _M_ROS2LOAM.T
    t = _M_ROS2LOAM @ p_ros
    return np.concatenate([loam_euler_from_R(R), t])


def make_voxel_map(scene: Scene, n_points: int, seed: int = 1, surf_leaf: float = 0.4, corner_leaf: float = 0.2,
                   tile_pitch=(45.0, 35.0)):
    """Map clouds at LOAM's own map density (loam_params.yaml:47-48: cornerFilterSize 0.2, surfaceFilterSize
    0.4): every scene surface is covered by a jittered lattice with one point per `surf_leaf` cell and every
    edge by one point per `corner_leaf`; the scene is then replicated on a grid of tiles (a building of
    identical rooms) until the map holds ~n_points, so a 1M-point map stays at realistic density while only
    the tile around the origin is ever observed.  Returns (corner, surf) float32 (N,4), LOAM axes.
    """
    g = np.random.default_rng(seed)
    lo, hi = _scene_bounds(scene, 60.0)
    surf = []

    def lattice(u0, u1, v0, v1, leaf):
        nu, nv = max(int((u1 - u0) / leaf), 1), max(int((v1 - v0) / leaf), 1)
        uu, vv = np.meshgrid(u0 + (np.arange(nu) + 0.5) * leaf, v0 + (np.arange(nv) + 0.5) * leaf, indexing="ij")
        uu = uu + g.uniform(-0.3, 0.3, uu.shape) * leaf
        vv = vv + g.uniform(-0.3, 0.3, vv.shape) * leaf
        return uu.ravel(), vv.ravel()

    for pl in scene.planes:
        ax = int(np.argmax(np.abs(pl[:3])))
        o = [a for a in range(3) if a != ax]
        u, v = lattice(lo[o[0]], hi[o[0]], lo[o[1]], hi[o[1]], surf_leaf)
        p = np.empty((u.size, 3))
        p[:, ax] = -pl[3] * pl[ax]
        p[:, o[0]] = u
        p[:, o[1]] = v
        surf.append(p)
    for cx, cy, r in scene.cylinders:
        th, z = lattice(0.0, 2 * math.pi * r, lo[2], hi[2], surf_leaf)
        surf.append(np.stack([cx + r * np.cos(th / r), cy + r * np.sin(th / r), z], -1))
    for b in scene.boxes:
        for ax in range(3):
            o = [a for a in range(3) if a != ax]
            for side in (0, 1):
                if ax == 2 and side == 0:
                    continue
                u, v = lattice(b[o[0]], b[3 + o[0]], b[o[1]], b[3 + o[1]], surf_leaf)
                p = np.empty((u.size, 3))
                p[:, ax] = b[ax + 3 * side]
                p[:, o[0]] = u
                p[:, o[1]] = v
                surf.append(p)
    surf = np.concatenate(surf)
    corner = []
    for a, b in _scene_edges(scene, (lo, hi)):
        L = np.linalg.norm(b - a)
        n = max(int(L / corner_leaf), 1)
        t = ((np.arange(n) + 0.5) / n)[:, None]
        corner.append(a[None] * (1 - t) + b[None] * t + g.uniform(-0.02, 0.02, (n, 3)))
    corner = np.concatenate(corner)
    per_tile = len(surf) + len(corner)
    n_tiles = max(1, int(round(n_points / per_tile)))
    side = int(math.ceil(math.sqrt(n_tiles)))
    offs = []
    for k in range(n_tiles):          # spiral-free simple grid centred on the origin tile
        i, j = k % side, k // side
        offs.append(((i - side // 2) * tile_pitch[0], (j - side // 2) * tile_pitch[1]))
    offs.sort(key=lambda o: abs(o[0]) + abs(o[1]))
    cs, ss = [], []
    for ox, oy in offs:
        d = np.array([ox, oy, 0.0])
        cs.append(corner + d)
        ss.append(surf + d)
    return _ros_to_loam4(np.concatenate(cs)), _ros_to_loam4(np.concatenate(ss))
