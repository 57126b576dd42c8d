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
