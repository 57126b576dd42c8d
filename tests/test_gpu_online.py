"""Online tick (vlo_process_scan) and the small host helpers of the C-ABI."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def test_online_sequence_matches_oracle_chain(orc):
    """C1-shaped: VLP-16 rolling-shutter sequence; every tick equals the oracle run the same way
    (seed = previous transform, previous features moved to their sweep end with it)."""
    from vil_sensor_fusion_b200 import api, synth
    traj = synth.Trajectory()
    ocfg = orc.default_config("VLP-16", deskew=1)
    gcfg = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=32768)
    raws = [scenes.vlp16_scan(0.1 * k) for k in range(5)]
    feats = []
    for raw in raws:
        c, rs, _ = orc.organise(ocfg, raw)
        f = orc.extract(ocfg, c, rs)
        f["cloud"] = c
        feats.append(f)
    T_prev = np.zeros(6, np.float32)
    sum_o = np.zeros(6, np.float32)
    with api.Handle(gcfg) as h:
        for k, raw in enumerate(raws):
            rc, odom, _ = h.process_scan(raw, stamp=0.1 * k)
            if k == 0:
                assert odom["status"] == 1
                continue
            f0, f1 = feats[k - 1], feats[k]
            lc = f0["cloud"][f0["less_sharp_idx"]]
            ls = f0["less_flat"]
            if k >= 2:
                lc = orc.transform_to_end(ocfg, T_prev, lc)
                ls = orc.transform_to_end(ocfg, T_prev, ls)
            ro = orc.odometry_register(ocfg, f1["cloud"][f1["sharp_idx"]], f1["cloud"][f1["flat_idx"]], lc,
                                       f0["less_sharp_ring_start"], ls, f0["less_flat_ring_start"], seed=T_prev)
            np.testing.assert_array_equal(odom["transform"].view(np.uint32), ro["transform"].view(np.uint32), err_msg="tick %d" % k)
            np.testing.assert_array_equal(odom["hessian"].view(np.uint32), ro["hessian"].view(np.uint32))
            assert odom["iterations"] == ro["iterations"]
            T_prev = ro["transform"]
            sum_o = orc.accumulate_pose(sum_o, T_prev)
        s, _ = h.online_pose()
        np.testing.assert_allclose(s, sum_o, atol=1e-6)
    # the chain tracks ground truth: accumulated pose ~ sensor pose at t = 0.5 relative to t = 0.1
    gt = synth.loam_sweep_transform(traj.rotation(0.5), traj.position(0.5), traj.rotation(0.1), traj.position(0.1))
    assert np.all(np.abs(sum_o[3:] - gt[3:]) < 0.05), (sum_o, gt)


def _euler_to_M(T):
    sx, cx, sy, cy, sz, cz = np.sin(T[0]), np.cos(T[0]), np.sin(T[1]), np.cos(T[1]), np.sin(T[2]), np.cos(T[2])
    M = np.eye(4)
    M[0, :3] = [cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx]
    M[1, :3] = [cx * sz, cx * cz, -sx]
    M[2, :3] = [-sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx]
    M[:3, 3] = T[3:6]
    return M


def _M_to_euler(M):
    return np.array([-np.arcsin(np.clip(M[1, 2], -1, 1)), np.arctan2(M[0, 2], M[2, 2]), np.arctan2(M[1, 0], M[1, 1]),
                     M[0, 3], M[1, 3], M[2, 3]], np.float32)


@pytest.mark.parametrize("io_ratio", [1, 2])
def test_online_with_maintained_map_matches_oracle_chain(orc, io_ratio):
    """The full online tick (multiScanRegistration -> laserOdometry -> laserMapping with its map maintenance ->
    transformAssociateToMap) over a rolling-shutter VLP-16 sequence equals the oracle chained the same way: odometry
    bit-exact, mapped pose within the north-star tolerance (the seed goes through float64 host trigonometry), map
    sizes equal; laserMapping runs on every io_ratio-th sweep (upstream: frameCount % ioRatio == 1)."""
    from vil_sensor_fusion_b200 import api, synth
    traj = synth.Trajectory()
    ocfg = orc.default_config("VLP-16", deskew=1, io_ratio=io_ratio)
    gcfg = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=32768, max_map_points=200000, io_ratio=io_ratio)
    raws = [scenes.vlp16_scan(0.1 * k) for k in range(6)]
    om = orc.LaserMap(ocfg, cap=200000)
    T_prev = np.zeros(6, np.float32)
    sum_o = np.zeros(6, np.float32)
    aft = np.zeros(6, np.float32)
    bef = np.zeros(6, np.float32)
    prev = None
    n_mapped = 0
    with api.Handle(gcfg) as h:
        h.map_reset()
        for k, raw in enumerate(raws):
            rc, odom, mapped = h.process_scan(raw, stamp=0.1 * k, want_map=True)
            c, rs, _ = orc.organise(ocfg, raw)
            f = orc.extract(ocfg, c, rs)
            lc, ls = c[f["less_sharp_idx"]], f["less_flat"]
            if k >= 1:
                ro = orc.odometry_register(ocfg, c[f["sharp_idx"]], c[f["flat_idx"]], prev[0], prev[1], prev[2], prev[3], seed=T_prev)
                np.testing.assert_array_equal(odom["transform"].view(np.uint32), ro["transform"].view(np.uint32), err_msg="tick %d" % k)
                T_prev = ro["transform"]
                sum_o = orc.accumulate_pose(sum_o, T_prev)
                lc = orc.transform_to_end(ocfg, T_prev, lc)
                ls = orc.transform_to_end(ocfg, T_prev, ls)
            prev = (lc, f["less_sharp_ring_start"], ls, f["less_flat_ring_start"])
            due = io_ratio < 2 or k % io_ratio == 1
            if not due:
                assert mapped["status"] == 1 and mapped["iterations"] == 0
                continue
            seed = _M_to_euler(_euler_to_M(aft) @ np.linalg.inv(_euler_to_M(bef)) @ _euler_to_M(sum_o))
            rm = om.process(lc, ls, seed)
            n_mapped += 1
            assert mapped["status"] == rm["status"], k
            dT = np.abs(mapped["transform"] - rm["transform"])
            assert np.all(dT[:3] <= 1e-5) and np.all(dT[3:] <= 1e-4), (k, dT)
            assert h.map_size() == (om.size(0), om.size(1)), k
            aft, bef = rm["transform"], sum_o.copy()
        _, m = h.online_pose()
        assert np.all(np.abs(m - aft) <= 1e-4)
    assert n_mapped == (6 if io_ratio == 1 else 3)
    # sanity against ground truth: odometry = motion since sweep 0's end (t = 0.1); the map frame is sweep 0 as it was
    # stored (not de-skewed, i.e. about mid-sweep), the mapped pose is sweep 5's end (t = 0.6)
    gt_o = synth.loam_sweep_transform(traj.rotation(0.6), traj.position(0.6), traj.rotation(0.1), traj.position(0.1))
    assert np.all(np.abs(sum_o[3:] - gt_o[3:]) < 0.06), (sum_o, gt_o)
    if io_ratio == 1:
        gt_m = synth.loam_sweep_transform(traj.rotation(0.6), traj.position(0.6), traj.rotation(0.05), traj.position(0.05))
        assert np.all(np.abs(aft[3:] - gt_m[3:]) < 0.08), (aft, gt_m)
    om.close()


def test_pose_diff_kat():
    """gtsam_fusion/test/UnitTests.cpp:183-233: identity -> (1,1,1) gives a between translation (1,1,1)."""
    from vil_sensor_fusion_b200 import api
    out = api.pose_diff([0, 0, 0, 1, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0])
    np.testing.assert_allclose(out, [1, 1, 1, 1, 0, 0, 0], atol=1e-15)


def test_host_helpers_vs_oracle(orc):
    from vil_sensor_fusion_b200 import api
    rng = np.random.default_rng(0)
    for _ in range(20):
        q1 = rng.normal(size=4); q1 /= np.linalg.norm(q1)
        q2 = rng.normal(size=4); q2 /= np.linalg.norm(q2)
        b = np.concatenate([rng.normal(size=3), q1])
        a = np.concatenate([rng.normal(size=3), q2])
        np.testing.assert_allclose(api.pose_diff(b, a), orc.pose_diff(b, a), atol=1e-14)
        A = rng.normal(size=(40, 6)) * np.array([30, 30, 30, 3, 3, 3])
        H = (A.T @ A).astype(np.float32)
        okg, lrg, ltg = api.dopt_gate(H)
        oko, lro, lto = orc.dopt_gate(H)
        assert okg == oko
        np.testing.assert_allclose([lrg, ltg], [lro, lto], rtol=1e-6)
        s = rng.normal(size=6).astype(np.float32) * 0.3
        T = rng.normal(size=6).astype(np.float32) * 0.05
        np.testing.assert_allclose(api.accumulate_pose(s, T), orc.accumulate_pose(s, T), atol=1e-6)
    # NaN log-det passes the gate exactly like the reference's comparison (degerate_odometry_filter.cpp:39)
    H = np.eye(6, dtype=np.float32)
    H[3, 3] = -1.0
    ok, lr, lt = api.dopt_gate(H, 11.5, -1.0)
    assert np.isnan(lr) and ok


def test_online_tick_from_a_pointcloud2_layout_equals_the_plain_tick():
    """vlo_process_scan_pc2: the tick fed with a PointCloud2 payload whose x / y / z sit at other byte offsets (intensity first,
    a ring column last: point_step 20) gives bit for bit what vlo_process_scan gives on the plain x y z i layout."""
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.1 * k) for k in range(3)]
    gcfg = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=65536)      # staging: 4 floats per point slot, the payload has 5
    with api.Handle(gcfg) as h:
        plain = [h.process_scan(r, 0.1 * k)[1] for k, r in enumerate(raws)]
    with api.Handle(gcfg) as h:
        for k, r in enumerate(raws):
            blob = np.zeros((r.shape[0], 5), np.float32)
            blob[:, 0] = 7.0                       # intensity
            blob[:, 1:4] = r[:, :3][:, [1, 0, 2]]  # y x z
            msg = {"data": blob.tobytes(), "point_step": 20, "fields": {"x": 8, "y": 4, "z": 12}}
            rc, o, _ = h.process_pointcloud2(msg, 0.1 * k)
            np.testing.assert_array_equal(o["transform"].view(np.uint32), plain[k]["transform"].view(np.uint32))
            np.testing.assert_array_equal(o["hessian"].view(np.uint32), plain[k]["hessian"].view(np.uint32))


def test_online_state_advances_through_a_map_full_report():
    """a one-shot capacity report of the mapping side (map full) must not freeze the odometry state: the tick
    that reports it still delivers its records and advances, so every later odometry result equals the run with a map that
    never fills."""
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.1 * k) for k in range(6)]
    big = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=32768, max_map_points=1 << 18, io_ratio=1)
    tiny = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=32768, max_map_points=2000, io_ratio=1)
    with api.Handle(big) as h:
        h.map_reset()
        ref = [h.process_scan(r, 0.1 * k, want_map=True)[1] for k, r in enumerate(raws)]
        ref_sum, _ = h.online_pose()
    reports = 0
    with api.Handle(tiny) as h:
        h.map_reset()
        got = []
        for k, r in enumerate(raws):
            # the record pointers are filled even when the call reports the capacity condition
            odom = api.Result()
            mapped = api.Result()
            r32 = np.ascontiguousarray(r, np.float32)
            import ctypes as C
            rc = h.lib.vlo_process_scan(h._h, r32.ctypes.data_as(C.c_void_p), r32.shape[0], r32.shape[1], 0.1 * k, C.byref(odom), C.byref(mapped))
            assert rc >= 0 or rc == -3, (rc, h.lib.vlo_last_error(h._h))
            reports += int(rc == -3)
            got.append(np.frombuffer(bytes(odom), api.RESULT_DTYPE)[0])
        got_sum, _ = h.online_pose()
    assert reports >= 1, "the 2000-voxel map should have filled up"
    for k in range(1, len(raws)):
        np.testing.assert_array_equal(got[k]["transform"].view(np.uint32), ref[k]["transform"].view(np.uint32), err_msg="tick %d" % k)
    np.testing.assert_array_equal(got_sum, ref_sum)
