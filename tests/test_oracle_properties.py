"""Self-consistency of the oracle: every frozen sub-algorithm against an independent numpy
statement of the same mathematics (so the restatement is not only compared with itself)."""
import ctypes as C

import numpy as np
import pytest

from tests import scenes


def test_detmath_accuracy(orc):
    L = orc.lib()
    x = np.linspace(-7, 7, 20001).astype(np.float32)
    s = np.zeros_like(x)
    c = np.zeros_like(x)
    # exercise through transform_to_start: rotation about z only -> cos/sin appear in the output
    cfg = orc.default_config("VLP-16", deskew=0)
    for ang in (0.0, 1e-4, 0.3, -1.2, 2.9, -3.1, 6.0):
        T = np.array([0, 0, ang, 0, 0, 0], np.float32)
        out = orc.transform_to_start(cfg, T, np.array([[1, 0, 0, 0]], np.float32))[0]
        # rotZ(-ang) applied to (1,0,0): (cos, -sin, 0)
        assert abs(out[0] - np.cos(ang)) < 3e-7 and abs(out[1] + np.sin(ang)) < 3e-7
    # atan through ring assignment: elevation sweep maps to the expected ring
    cfg16 = orc.default_config("VLP-16")
    el = np.deg2rad(np.linspace(-15, 15, 16))
    raw = np.stack([np.cos(el) * 10, np.zeros(16), np.sin(el) * 10, np.ones(16)], -1).astype(np.float32)
    cloud, rs, src = orc.organise(cfg16, raw)
    assert np.array_equal(np.diff(rs), np.ones(16, int))
    assert np.array_equal(src, np.arange(16))


def test_knn_kdtree_equals_brute(orc):
    rng = np.random.default_rng(1)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-10, 10, (5000, 3))
    pts[2500:2600] = pts[:100]                  # exact duplicates: ties -> lowest index
    q = np.zeros((400, 4), np.float32)
    q[:, :3] = rng.uniform(-11, 11, (400, 3))
    q[:50] = pts[:50]
    for k in (1, 5):
        ib, db = orc.knn_brute(pts, q, k)
        ik, dk = orc.knn_kdtree(pts, q, k)
        np.testing.assert_array_equal(ib, ik)
        np.testing.assert_array_equal(db.view(np.uint32), dk.view(np.uint32))
    assert np.all(ib[:50, 0] == np.arange(50))
    # independent numpy statement
    d = ((pts[None, :, :3] - q[:, None, :3]).astype(np.float32) ** 2)
    d2 = (d[..., 0] + d[..., 1]) + d[..., 2]
    np.testing.assert_array_equal(np.argmin(d2, axis=1), orc.knn_brute(pts, q, 1)[0][:, 0])


def test_dense6_against_numpy(orc):
    rng = np.random.default_rng(0)
    for trial in range(30):
        A = rng.normal(size=(60, 6)) * np.array([30, 20, 25, 2, 3, 1.5])
        H = (A.T @ A).astype(np.float32)
        b = rng.normal(size=6).astype(np.float32) * 10
        x = orc.solve6(H, b)
        np.testing.assert_allclose(x, np.linalg.solve(H.astype(np.float64), b), rtol=2e-3, atol=1e-5)
        ev, vec = orc.eig6(H)
        w, v = np.linalg.eigh(H.astype(np.float64))
        np.testing.assert_allclose(ev, w, rtol=1e-4, atol=1e-3 * w[-1] * 1e-3)
        assert np.all(np.diff(ev) >= 0)
        # rows are eigenvectors: H v = lambda v
        for i in range(6):
            np.testing.assert_allclose(H.astype(np.float64) @ vec[i], ev[i] * vec[i], atol=2e-3 * w[-1] * 1e-2 + 1e-2)
        np.testing.assert_allclose(vec @ vec.T, np.eye(6), atol=1e-5)


def test_degeneracy_projection(orc):
    # H with two weak directions (along e3 + e4 mix): eigenvalues 1, 5 < 30
    rng = np.random.default_rng(3)
    Q, _ = np.linalg.qr(rng.normal(size=(6, 6)))
    lam = np.array([1.0, 5.0, 100.0, 400.0, 900.0, 2500.0])
    H = (Q @ np.diag(lam) @ Q.T).astype(np.float32)
    flag, ev, P = orc.degeneracy(H, 30.0)
    assert flag
    np.testing.assert_allclose(ev, lam, rtol=1e-4)
    Pexp = Q[:, 2:] @ Q[:, 2:].T
    np.testing.assert_allclose(P, Pexp, atol=2e-5)
    np.testing.assert_allclose(P @ P, P, atol=1e-5)          # a projection
    flag2, _, P2 = orc.degeneracy(H, 0.5)
    assert not flag2
    np.testing.assert_allclose(P2, np.eye(6), atol=1e-5)


def test_odometry_jacobian_finite_difference(orc):
    L = orc.lib()
    cfg = orc.default_config("VLP-16", deskew=0)
    rng = np.random.default_rng(0)
    for _ in range(5):
        T = (rng.normal(size=6) * np.array([0.03, 0.03, 0.03, 0.3, 0.3, 0.3])).astype(np.float32)
        p = np.concatenate([rng.uniform(-10, 10, 3), [0]]).astype(np.float32)[None]
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        coeff = np.array([n[0], n[1], n[2], 0.3], np.float32)
        row = np.zeros(6, np.float32)
        b = C.c_float()
        L.orc_odom_jacobian_row(T.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), coeff.ctypes.data_as(C.c_void_p),
                                row.ctypes.data_as(C.c_void_p), C.byref(b))

        def d(Tv):
            q = orc.transform_to_start(cfg, Tv.astype(np.float32), p)[0, :3].astype(np.float64)
            return n @ q
        num = np.array([(d(T + e) - d(T - e)) / 2e-3 for e in np.eye(6) * 1e-3])
        np.testing.assert_allclose(row, num, atol=3e-3)
        assert abs(b.value + 0.05 * 0.3) < 1e-7


def test_mapping_jacobian_finite_difference(orc):
    L = orc.lib()
    rng = np.random.default_rng(1)
    for _ in range(5):
        T = (rng.normal(size=6) * np.array([0.2, 0.5, 0.2, 3, 3, 3])).astype(np.float32)
        p = np.concatenate([rng.uniform(-10, 10, 3), [0]]).astype(np.float32)[None]
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        coeff = np.array([n[0], n[1], n[2], -0.2], np.float32)
        row = np.zeros(6, np.float32)
        b = C.c_float()
        L.orc_map_jacobian_row(T.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), coeff.ctypes.data_as(C.c_void_p),
                               row.ctypes.data_as(C.c_void_p), C.byref(b))

        def d(Tv):
            out = np.zeros_like(p)
            Tv = Tv.astype(np.float32)
            L.orc_point_to_map(Tv.ctypes.data_as(C.c_void_p), p.ctypes.data_as(C.c_void_p), 1, out.ctypes.data_as(C.c_void_p))
            return n @ out[0, :3].astype(np.float64)
        num = np.array([(d(T + e) - d(T - e)) / 2e-3 for e in np.eye(6) * 1e-3])
        np.testing.assert_allclose(row, num, atol=5e-3)
        assert b.value == np.float32(0.2)


def test_extraction_invariants(orc):
    cfg = orc.default_config("VLP-16")
    raw = scenes.vlp16_scan(0.0, noise=0.02, seed=4)
    c, rs, src = orc.organise(cfg, raw)
    # organise: stable per-ring order, ring id in the intensity's integer part, rel-time in [0, scanPeriod)
    ring = c[:, 3].astype(int)
    assert np.all(np.diff(ring) >= 0)
    for r in range(16):
        assert np.all(np.diff(src[rs[r]:rs[r + 1]]) > 0)
    frac = c[:, 3] - ring
    assert frac.min() >= 0 and frac.max() < 0.1001
    np.testing.assert_array_equal(c[:, 0], raw[src, 1])      # LOAM x <- ROS y
    np.testing.assert_array_equal(c[:, 1], raw[src, 2])
    np.testing.assert_array_equal(c[:, 2], raw[src, 0])
    f = orc.extract(cfg, c, rs)
    lab = f["label"]
    assert set(np.unique(lab)) <= {-1, 0, 1, 2}
    assert np.array_equal(np.sort(np.where(lab == 2)[0]), np.sort(f["sharp_idx"]))
    assert np.array_equal(np.sort(np.where(lab >= 1)[0]), np.sort(f["less_sharp_idx"]))
    assert np.array_equal(np.sort(np.where(lab == -1)[0]), np.sort(f["flat_idx"]))
    # caps per (ring, sector): 2 sharp / 20 less sharp / 4 flat  (loam_params.yaml:27-29)
    assert len(f["sharp_idx"]) <= 16 * 6 * 2 and len(f["less_sharp_idx"]) <= 16 * 6 * 20 and len(f["flat_idx"]) <= 16 * 6 * 4
    thr = cfg.surface_curvature_threshold
    assert np.all(f["curvature"][lab >= 1] > thr) and np.all(f["curvature"][lab == -1] < thr)
    # curvature against an independent numpy statement (same summation order)
    i = int(rs[3] + 400)
    w = np.float32(-10.0)
    d = w * c[i, :3]
    for m in range(1, 6):
        d = d + (c[i + m, :3] + c[i - m, :3])
    assert f["curvature"][i] == np.float32((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
    # every picked feature suppressed itself
    assert np.all(f["picked"][lab != 0] == 1)
    # less-flat: one centroid per occupied 0.2 m voxel per ring
    lf = f["less_flat"]
    lfr = f["less_flat_ring_start"]
    for r in (2, 9):
        seg = lf[lfr[r]:lfr[r + 1]]
        assert np.all(seg[:, 3].astype(int) == r)
        cand = c[rs[r]:rs[r + 1]]
        cand = cand[(lab[rs[r]:rs[r + 1]] <= 0)]
        vox = np.floor(cand[5:-5, :3] * np.float32(5.0)).astype(int)   # approx: sector range trims 5 at each end
        assert abs(len(np.unique(vox, axis=0)) - len(seg)) <= 2


def test_reduce_r1_is_blocked_sum(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    Q = 2500
    terms = rng.normal(size=(Q, 28)).astype(np.float32)
    total = np.zeros(28, np.float32)
    L.orc_reduce_r1(terms.ctypes.data_as(C.c_void_p), Q, 28, total.ctypes.data_as(C.c_void_p))
    exp = np.zeros(28, np.float32)
    for e in range(28):
        l3 = np.float32(0)
        for c2 in range(0, Q, 1024):
            l2 = np.float32(0)
            for c1 in range(c2, min(Q, c2 + 1024), 32):
                l1 = np.float32(0)
                for v in terms[c1:min(Q, c1 + 32), e]:
                    l1 = np.float32(l1 + v)
                l2 = np.float32(l2 + l1)
            l3 = np.float32(l3 + l2)
        exp[e] = l3
    np.testing.assert_array_equal(total.view(np.uint32), exp.view(np.uint32))
    np.testing.assert_allclose(total, terms.astype(np.float64).sum(0), atol=1e-3)


def test_imu_covariance_jacobian_numeric(orc):
    """A one-step preintegration: dP_dba / dV_dbg against finite differences over the bias."""
    prm = orc.imu_params()
    rng = np.random.default_rng(0)
    t = np.arange(1, 21) * 0.005
    acc = rng.normal(0, 1, (20, 3)) + [0, 0, 9.81]
    gyro = rng.normal(0, 0.2, (20, 3))
    f0 = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001)
    h = 1e-6
    for axis in range(3):
        db = np.zeros(6)
        db[axis] = h
        fp = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001, bias=db)
        fm = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001, bias=-db)
        np.testing.assert_allclose((fp["dP"] - fm["dP"]) / (2 * h), f0["dP_dba"][:, axis], atol=1e-6)
        np.testing.assert_allclose((fp["dV"] - fm["dV"]) / (2 * h), f0["dV_dba"][:, axis], atol=1e-6)
        db = np.zeros(6)
        db[3 + axis] = h
        fp = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001, bias=db)
        fm = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001, bias=-db)
        np.testing.assert_allclose((fp["dV"] - fm["dV"]) / (2 * h), f0["dV_dbg"][:, axis], atol=1e-5)
        np.testing.assert_allclose((fp["dP"] - fm["dP"]) / (2 * h), f0["dP_dbg"][:, axis], atol=1e-5)
    # rotation is orthonormal, covariance symmetric PSD and growing with the window
    np.testing.assert_allclose(f0["dR"] @ f0["dR"].T, np.eye(3), atol=1e-12)
    f_half = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.05)
    assert np.all(np.diag(f0["cov"])[:9] > np.diag(f_half["cov"])[:9])


def test_registration_recovers_motion(orc):
    """well-conditioned room: seeded like the online path, scan-to-scan lands within LOAM accuracy."""
    from vil_sensor_fusion_b200 import synth
    traj = synth.Trajectory()
    cfg = orc.default_config("VLP-16")
    fe = []
    for t0 in (0.0, 0.1):
        c, rs, _ = orc.organise(cfg, scenes.vlp16_scan(t0))
        f = orc.extract(cfg, c, rs)
        fe.append((c, f))
    gt0 = synth.loam_sweep_transform(traj.rotation(0.0), traj.position(0.0), traj.rotation(0.1), traj.position(0.1))
    gt1 = synth.loam_sweep_transform(traj.rotation(0.1), traj.position(0.1), traj.rotation(0.2), traj.position(0.2))
    (c0, f0), (c1, f1) = fe
    lc = orc.transform_to_end(cfg, gt0, c0[f0["less_sharp_idx"]])
    ls = orc.transform_to_end(cfg, gt0, f0["less_flat"])
    r = orc.odometry_register(cfg, c1[f1["sharp_idx"]], c1[f1["flat_idx"]], lc, f0["less_sharp_ring_start"], ls,
                              f0["less_flat_ring_start"], seed=gt0)
    assert r["status"] == 0 and not r["is_degenerate"]
    assert np.all(np.abs(r["transform"][:3] - gt1[:3]) < 1e-3) and np.all(np.abs(r["transform"][3:] - gt1[3:]) < 1e-2)


def test_voxel_downsample_is_a_voxel_grid(orc):
    """pcl::VoxelGrid restated (V1/V2): one output per occupied voxel, in order of first appearance, each the float64
    centroid of its members to within the 2^-20 m quantisation; leaf <= 0 passes the cloud through."""
    rng = np.random.default_rng(4)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-20, 20, (5000, 3))
    pts[:, 3] = rng.integers(0, 16, 5000) + rng.uniform(0, 0.09, 5000)
    pts[100:200, :3] = pts[0:100, :3] + 1e-3          # guaranteed shared voxels
    for leaf in (0.2, 0.4, 1.0):
        out = orc.voxel_downsample(pts, leaf)
        vox = np.floor(pts[:, :3].astype(np.float32) * np.float32(1.0 / leaf)).astype(np.int64)
        keys, first, inv = np.unique(vox, axis=0, return_index=True, return_inverse=True)
        order = np.argsort(first)                     # order of first appearance
        assert len(out) == len(keys)
        for rank, kidx in enumerate(order[:300]):
            members = pts[inv.ravel() == kidx]
            np.testing.assert_allclose(out[rank, :3], members[:, :3].astype(np.float64).mean(axis=0), atol=2e-6)
            assert int(out[rank, 3]) == int(pts[first[kidx], 3])
        assert np.all(np.floor(out[:, :3] * np.float32(1.0 / leaf)) == keys[order]) or leaf == 1.0
    np.testing.assert_array_equal(orc.voxel_downsample(pts, 0.0), pts)


def test_laser_map_maintenance_semantics(orc):
    """oracle/laser_map.c, the frozen choices M1-M6: insertion creates one point per (cube, voxel) numbered by first
    appearance; re-inserting the same cloud creates nothing and averages (c + p) / 2 = c; the FOV-valid sub-map leaves
    out the cubes straight above / below the sensor; the window shift evicts cubes for good."""
    from vil_sensor_fusion_b200 import synth
    scene = synth.scene_room(0)
    cm, sm = synth.sample_map_points(scene, 20000, seed=2)
    cfg = orc.default_config("VLP-16", deskew=0)
    ident = np.zeros(6, np.float32)
    lm = orc.LaserMap(cfg, cap=50000)
    lm.insert(cm, sm, ident)
    n0 = (lm.size(0), lm.size(1))
    p0, c0 = lm.points(1)
    # one point per voxel, every centroid inside the voxel it was created for
    vox = np.floor(p0[:, :3] * np.float32(1.0 / 0.4)).astype(np.int64)
    assert len(np.unique(np.concatenate([vox, c0[:, None].astype(np.int64)], axis=1), axis=0)) == len(p0)
    # creation order = first appearance of the voxel in the inserted cloud
    vin = np.floor(sm[:, :3] * np.float32(1.0 / 0.4)).astype(np.int64)
    _, first = np.unique(vin, axis=0, return_index=True)
    np.testing.assert_array_equal(np.floor(sm[np.sort(first), :3] * np.float32(2.5)).astype(np.int64)[:200], vox[:200])
    lm.insert(cm, sm, ident)
    assert (lm.size(0), lm.size(1)) == n0
    lm.insert(p0, p0[:0], ident)                      # nothing new from the corner side either
    p1, _ = lm.points(1)
    assert np.abs(p1[:, :3] - p0[:, :3]).max() < 0.2   # centroids moved inside their voxels only
    # FOV test: sensor level at the origin -> the cube column straight above and below fails, the ring around passes
    centre, mask = lm.select(ident)
    nb = cfg.n_neighbor_cubes
    assert mask[nb, nb, nb + 1] == 1 and mask[nb, nb + 1, nb] == 1       # (k, j, i) order: neighbours in z and in x
    assert mask[nb, nb + 3, nb] == 0 and mask[nb, nb - 3, nb] == 0       # 3 cubes straight up / down (LOAM y is up)
    lm.close()
    # window shift: 7^3 cubes of 4 m; moving 3 cubes along x shifts the window and evicts the far cubes
    small = orc.default_config("VLP-16", deskew=0, map_cube_size=4.0, n_neighbor_cubes=1)
    for a in range(3):
        small.map_dims[a] = 7
        small.map_start_cubes[a] = 3
    lm = orc.LaserMap(small, cap=50000)
    lm.insert(cm, sm, ident)
    before = lm.points(1)[1]
    assert not np.any(before & (1 << 30))
    far = np.array([0, 0, 0, 12.5, 0, 0], np.float32)
    lm.select(far)
    after = lm.points(1)[1]
    assert np.array_equal(lm.window(), [2, 3, 3]) or lm.window()[0] < 3
    assert np.any(after & (1 << 30)) and np.count_nonzero(after & (1 << 30)) < len(after)
    lm.close()


def test_organise_input_rotation_and_ring_field(orc):
    """rotateInputCloud / inputCloudRotation and useCloudIntensityandRingFields (loam_params.yaml:4-5,23) in the oracle:
    organising with the rotation option equals organising a cloud that numpy rotated by Rz(yaw) Ry(pitch) Rx(roll)
    beforehand (same rings, same order, coordinates to float rounding); with a ring field the ids come from the field,
    out-of-range and NaN entries are dropped, and the order inside a ring is still the arrival order."""
    raw = scenes.vlp16_scan(0.0, noise=0.01, seed=3, n_az=600)
    y, p, r = 0.3, -0.1, 0.05
    Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
    Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
    Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
    pre = raw.copy()
    pre[:, :3] = (Rz @ Ry @ Rx @ raw[:, :3].astype(np.float64).T).T.astype(np.float32)
    c_rot, rs_rot, src_rot = orc.organise(orc.default_config("VLP-16", rotate_input=1, input_rotation=(y, p, r)), raw)
    c_pre, rs_pre, src_pre = orc.organise(orc.default_config("VLP-16"), pre)
    # a point within float rounding of a ring boundary may change ring between the two; everything else must agree
    assert abs(len(c_rot) - len(c_pre)) <= 2
    common = np.intersect1d(src_rot, src_pre)
    assert len(common) >= len(src_rot) - 4
    pos_rot = np.full(len(raw), -1); pos_rot[src_rot] = np.arange(len(src_rot))
    pos_pre = np.full(len(raw), -1); pos_pre[src_pre] = np.arange(len(src_pre))
    a, b = c_rot[pos_rot[common]], c_pre[pos_pre[common]]
    np.testing.assert_allclose(a[:, :3], b[:, :3], atol=2e-5)
    assert np.mean(np.floor(a[:, 3]) == np.floor(b[:, 3])) > 0.999
    # identity rotation with the flag on changes nothing, bit for bit
    c_id, rs_id, src_id = orc.organise(orc.default_config("VLP-16", rotate_input=1), raw)
    c_0, rs_0, src_0 = orc.organise(orc.default_config("VLP-16"), raw)
    np.testing.assert_array_equal(src_id, src_0)
    np.testing.assert_array_equal(c_id.view(np.uint32), c_0.view(np.uint32))
    # ring field: reversed ids, some invalid
    ring_by_src = np.full(len(raw), -1.0, np.float32)
    ring_by_src[src_0] = np.floor(c_0[:, 3])
    raw5 = np.concatenate([raw, (15.0 - ring_by_src)[:, None].astype(np.float32)], axis=1)
    raw5[ring_by_src < 0, 4] = -1.0
    raw5[7::101, 4] = np.nan
    raw5[11::103, 4] = 16.0
    c_f, rs_f, src_f = orc.organise(orc.default_config("VLP-16", ring_field=4), raw5)
    np.testing.assert_array_equal(np.floor(c_f[:, 3]), raw5[src_f, 4])
    assert np.all(np.isfinite(raw5[src_f, 4])) and np.all((raw5[src_f, 4] >= 0) & (raw5[src_f, 4] < 16))
    for ring in range(16):
        seg = src_f[rs_f[ring]:rs_f[ring + 1]]
        assert np.all(np.diff(seg) > 0)                                   # arrival order inside a ring
    valid = np.isfinite(raw5[:, 4]) & (raw5[:, 4] >= 0) & (raw5[:, 4] < 16)
    assert len(src_f) == int(valid.sum())
