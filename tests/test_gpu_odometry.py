"""Scan-to-scan registration (K2 grids + K3 association + fused GN + K4 solve/degeneracy) vs oracle.

Bar (BASELINE.json north_star): correspondence indices bit-exact; pose within 1e-4 m / 1e-5 rad;
eigenvalues within 1e-4 relative.  Because both sides freeze the same operation order the pose is in
fact compared bit-for-bit here, with the stated tolerances as the fallback assertion message.
"""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _oracle_pair(orc, ocfg, raw_last, raw_cur, seed, last_T):
    c0, rs0, _ = orc.organise(ocfg, raw_last)
    c1, rs1, _ = orc.organise(ocfg, raw_cur)
    f0 = orc.extract(ocfg, c0, rs0)
    f1 = orc.extract(ocfg, c1, rs1)
    lc = c0[f0["less_sharp_idx"]]
    ls = f0["less_flat"]
    if last_T is not None:
        lc = orc.transform_to_end(ocfg, last_T, lc)
        ls = orc.transform_to_end(ocfg, last_T, ls)
    r = orc.odometry_register(ocfg, c1[f1["sharp_idx"]], c1[f1["flat_idx"]], lc, f0["less_sharp_ring_start"], ls,
                              f0["less_flat_ring_start"], seed=seed, use_kdtree=True, trace=True)
    return r, len(f1["sharp_idx"]), len(f1["flat_idx"])


def _compare(ro, rg, tag=""):
    assert rg["iterations"] == ro["iterations"], tag
    assert rg["n_corr_edge"] == ro["n_corr_edge"] and rg["n_corr_plane"] == ro["n_corr_plane"], tag
    assert bool(rg["is_degenerate"]) == ro["is_degenerate"], tag
    dT = np.abs(rg["transform"] - ro["transform"])
    assert np.all(dT[:3] <= 1e-5) and np.all(dT[3:] <= 1e-4), (tag, dT)
    np.testing.assert_array_equal(rg["transform"].view(np.uint32), ro["transform"].view(np.uint32), err_msg=tag + " transform bits")
    np.testing.assert_allclose(rg["eig"], ro["eig"], rtol=1e-4, err_msg=tag)
    np.testing.assert_array_equal(rg["hessian"].view(np.uint32), ro["hessian"].view(np.uint32), err_msg=tag + " hessian bits")
    np.testing.assert_array_equal(rg["P"].view(np.uint32), ro["P"].view(np.uint32), err_msg=tag + " P bits")
    np.testing.assert_allclose(rg["logdet_rot"], ro["logdet_rot"], rtol=1e-5)
    np.testing.assert_allclose(rg["logdet_trans"], ro["logdet_trans"], rtol=1e-5)
    assert bool(rg["pass_dopt"]) == ro["pass_dopt"]
    np.testing.assert_allclose(rg["cov"], ro["cov"], rtol=1e-6, atol=1e-14)


@pytest.mark.parametrize("deskew", [1, 0])
def test_vlp16_pair_bit_exact(orc, deskew):
    from vil_sensor_fusion_b200 import api, synth
    traj = synth.Trajectory()
    raw0 = scenes.vlp16_scan(0.0, rolling=bool(deskew))
    raw1 = scenes.vlp16_scan(0.1, rolling=bool(deskew))
    ocfg = orc.default_config("VLP-16", deskew=deskew)
    gcfg = api.default_config("VLP-16", deskew=deskew, max_scans=2, max_points=32768)
    gt0 = synth.loam_sweep_transform(traj.rotation(0.0), traj.position(0.0), traj.rotation(0.1), traj.position(0.1)).astype(np.float32)
    for seed, last_T in ((None, None), (gt0, gt0 if deskew else None)):
        ro, n_sharp, n_flat = _oracle_pair(orc, ocfg, raw0, raw1, seed, last_T)
        with api.Handle(gcfg) as h:
            h.lib.vlo_set_trace(h._h, 1)
            h.upload([raw0, raw1])
            h.organise()
            h.extract()
            rg = h.register_pairs([0], [1], seeds=None if seed is None else [seed],
                                  last_transforms=None if last_T is None else [last_T])[0]
            n_rounds = (ro["iterations"] + 4) // 5
            per = 2 * n_sharp + 3 * n_flat
            for rnd in range(n_rounds):
                ci, si = h.pair_correspondences(0, rnd, n_sharp, n_flat)
                tr = ro["trace_idx"][rnd * per:(rnd + 1) * per]
                np.testing.assert_array_equal(ci.ravel(), tr[:2 * n_sharp], err_msg="corner idx round %d" % rnd)
                np.testing.assert_array_equal(si.ravel(), tr[2 * n_sharp:], err_msg="surf idx round %d" % rnd)
        _compare(ro, rg, "vlp16 deskew=%d seeded=%s" % (deskew, seed is not None))
        assert ro["n_corr_plane"] > 100


def test_hdl64_pairs_batch(orc):
    """three HDL-64 scans -> two independent pairs in one launch (whole-bag mode, rigid, zero seed)"""
    from vil_sensor_fusion_b200 import api
    raws = [scenes.hdl64_scan(0.0), scenes.hdl64_scan(0.1), scenes.hdl64_scan(0.2, noise=0.01, seed=7)]
    ocfg = orc.default_config("HDL-64E", deskew=0)
    gcfg = api.default_config("HDL-64E", deskew=0, max_scans=3, max_points=131072)
    with api.Handle(gcfg) as h:
        h.lib.vlo_set_trace(h._h, 1)
        h.upload(raws)
        h.organise()
        h.extract()
        res = h.register_pairs([0, 1], [1, 2])
        for p in range(2):
            ro, n_sharp, n_flat = _oracle_pair(orc, ocfg, raws[p], raws[p + 1], None, None)
            ci, si = h.pair_correspondences(p, 0, n_sharp, n_flat)
            per = 2 * n_sharp + 3 * n_flat
            np.testing.assert_array_equal(ci.ravel(), ro["trace_idx"][:2 * n_sharp])
            np.testing.assert_array_equal(si.ravel(), ro["trace_idx"][2 * n_sharp:per])
            _compare(ro, res[p], "hdl64 pair %d" % p)


@pytest.mark.parametrize("lidar,max_points", [("VLP-16", 32768), ("HDL-64E", 131072)])
def test_degenerate_corridor_and_plane(orc, lidar, max_points):
    """C3, VLP-16 and HDL-64.  Corridor along x with no end caps: the registration runs (status 0), is flagged
    degenerate with two eigenvalues of AtA under odomDegenEigVal (translation along the axis and its coupled rotation), the
    update is remapped (pose, eigenvalues, projection compared with the oracle: pose and P bit for bit, eigenvalues 1e-4),
    the along-axis motion is NOT recovered, and the D-opt gate drops the message (logdet_trans < 28.9).  Single ground plane:
    no corner features at all, so upstream's precondition (more than 10 corner / 100 surface points in the last sweep) refuses
    to optimise -- the soft status, identically on both sides; three dropped dimensions cannot arise through a registration."""
    from vil_sensor_fusion_b200 import api, synth
    ocfg = orc.default_config(lidar, deskew=0)
    gcfg = api.default_config(lidar, deskew=0, max_scans=2, max_points=max_points)
    motion = np.array([0.1, 0.02, 0.0])                      # ROS frame: 10 cm along the corridor, 2 cm across
    out = {}
    for scene in (synth.scene_corridor(), synth.scene_plane()):
        raw0 = synth.make_scan(scene, lidar, pose=(np.eye(3), np.zeros(3)), rolling=False)
        raw1 = synth.make_scan(scene, lidar, pose=(np.eye(3), motion), rolling=False)
        ro, _, _ = _oracle_pair(orc, ocfg, raw0, raw1, None, None)
        with api.Handle(gcfg) as h:
            h.upload([raw0, raw1])
            h.organise()
            h.extract()
            rg = h.register_pairs([0], [1])[0]
        out[scene.name] = (ro, rg)
    ro, rg = out["corridor"]
    assert ro["status"] == 0 and rg["status"] == 0
    _compare(ro, rg, "corridor")
    assert rg["is_degenerate"] == 1
    assert int(np.sum(rg["eig"] < ocfg.odom_degen_eig)) == 2 and int(np.sum(ro["eig"] < ocfg.odom_degen_eig)) == 2
    assert rg["pass_dopt"] == 0 and rg["logdet_trans"] < 28.9          # degerate_odometry_filter.cpp:39-42 drops it
    # LOAM axes: x = ROS y (across), z = ROS x (along).  Across-corridor motion is observed, along-axis motion is projected out
    assert abs(rg["transform"][5]) < 1e-3 and 0.005 < abs(rg["transform"][3]) < 0.025, rg["transform"]
    # the projection really removes the dropped eigen-directions: P is idempotent and of rank 4
    P = rg["P"].reshape(6, 6).astype(np.float64)
    np.testing.assert_allclose(P @ P, P, atol=1e-5)
    assert abs(np.trace(P) - 4.0) < 1e-4
    ro, rg = out["plane"]
    assert ro["status"] == 1 and rg["status"] == 1 and rg["iterations"] == 0 and ro["iterations"] == 0
    assert np.all(rg["transform"] == 0) and rg["is_degenerate"] == 0


def test_too_few_features_soft_status(orc):
    from vil_sensor_fusion_b200 import api
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=32768)
    tiny = scenes.vlp16_scan(0.0)[:200]
    with api.Handle(gcfg) as h:
        h.upload([tiny, tiny])
        h.organise()
        h.extract()
        r = h.register_pairs([0], [1])[0]
        assert r["status"] == 1 and r["iterations"] == 0
        assert np.all(r["transform"] == 0)


def test_hessian_order_option(orc):
    """vlo_config.hessian_order = 1 publishes AtA / covariance in (t; r) order: a pure permutation of the
    default record, with the D-opt gate re-read on the permuted matrix exactly as degerate_odometry_filter.cpp:32-46
    would (block(0,0) -> "trans", block(3,3) -> "rot")."""
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.0, rolling=False), scenes.vlp16_scan(0.1, rolling=False)]
    out = []
    for order in (0, 1):
        gcfg = api.default_config("VLP-16", deskew=0, max_scans=2, max_points=32768, hessian_order=order)
        with api.Handle(gcfg) as h:
            h.upload(raws)
            h.organise()
            h.extract()
            out.append(h.register_pairs([0], [1])[0])
    a, b = out
    perm = [3, 4, 5, 0, 1, 2]
    Ha, Hb = a["hessian"].reshape(6, 6), b["hessian"].reshape(6, 6)
    np.testing.assert_array_equal(Hb, Ha[np.ix_(perm, perm)])
    np.testing.assert_array_equal(b["cov"].reshape(6, 6), a["cov"].reshape(6, 6)[np.ix_(perm, perm)])
    np.testing.assert_array_equal(a["transform"], b["transform"])
    np.testing.assert_allclose(b["logdet_trans"], a["logdet_rot"], rtol=1e-5)
    np.testing.assert_allclose(b["logdet_rot"], a["logdet_trans"], rtol=1e-5)
    passed, lr, lt = api.dopt_gate(Hb)
    assert passed == bool(b["pass_dopt"])
    np.testing.assert_allclose([lr, lt], [b["logdet_rot"], b["logdet_trans"]], rtol=1e-6)


@pytest.mark.parametrize("lidar,max_points,cell", [("VLP-16", 32768, 1.0), ("HDL-64E", 131072, 0.7), ("VLP-16", 32768, 0.35)])
def test_pairs_batch_thread_search_equals_oracle_and_single_pair_path(orc, lidar, max_points, cell):
    """More than four pairs per call run the thread-per-query association (k3_assoc_thread: per-lane 27-cell search with
    the exact warp-cooperative search as fallback; the 0.35 m cells make most VLP-16 partner searches take the fallback).
    First-round correspondence indices and the converged result must equal the oracle's for two of the pairs, and every
    pair must equal the same pair registered alone (warp-per-query kernel), bit for bit."""
    from vil_sensor_fusion_b200 import api
    mk = scenes.vlp16_scan if lidar == "VLP-16" else scenes.hdl64_scan
    raws = [mk(0.1 * k, noise=0.01, seed=k) for k in range(7)]
    ocfg = orc.default_config(lidar, deskew=0)
    gcfg = api.default_config(lidar, deskew=0, max_scans=7, max_points=max_points, odom_cell_size=cell)
    last, cur = np.arange(6), np.arange(1, 7)
    with api.Handle(gcfg) as h:
        h.lib.vlo_set_trace(h._h, 1)
        h.upload(raws)
        h.organise()
        h.extract()
        res = h.register_pairs(last, cur)
        corr = {}
        for p in (0, 4):
            ro, n_sharp, n_flat = _oracle_pair(orc, ocfg, raws[p], raws[p + 1], None, None)
            corr[p] = (ro, n_sharp, n_flat, h.pair_correspondences(p, 0, n_sharp, n_flat))
        singles = [h.register_pairs([p], [p + 1])[0] for p in range(6)]
    for p, (ro, n_sharp, n_flat, (ci, si)) in corr.items():
        per = 2 * n_sharp + 3 * n_flat
        np.testing.assert_array_equal(ci.ravel(), ro["trace_idx"][:2 * n_sharp], err_msg="corner correspondences, pair %d" % p)
        np.testing.assert_array_equal(si.ravel(), ro["trace_idx"][2 * n_sharp:per], err_msg="surface correspondences, pair %d" % p)
        _compare(ro, res[p], "%s batch pair %d" % (lidar, p))
    for p in range(6):
        for f in ("transform", "hessian", "eig", "P"):
            np.testing.assert_array_equal(np.asarray(res[f][p]).view(np.uint32), np.asarray(singles[p][f]).view(np.uint32),
                                          err_msg="pair %d field %s: batch vs single-pair path" % (p, f))
        for f in ("iterations", "n_corr_edge", "n_corr_plane", "status", "is_degenerate"):
            assert res[f][p] == singles[p][f], (p, f)


def _late_first_column(raw, rings, k):
    """the sweep as a driver delivers it when the packet that opens the message is not the one with the smallest azimuth:
    column k of the (column-major) cloud first, then columns 0 .. k-1, then the rest.  startOri is taken from the first point,
    so the k moved columns get a slightly NEGATIVE relTime and their intensity ring + relTime reads as ring - 1 through
    upstream's int(intensity)"""
    cols = raw.reshape(-1, rings, raw.shape[1])
    return np.concatenate([cols[k:k + 1], cols[:k], cols[k + 1:]]).reshape(raw.shape).copy()


@pytest.mark.parametrize("lidar,rings,max_points", [("VLP-16", 16, 32768), ("HDL-64E", 64, 131072)])
def test_scan_ids_read_from_intensity_like_upstream(orc, lidar, rings, max_points):
    """Upstream's partner loops take a target point's scan id from int(intensity), not from the ring it was filed
    under; the two differ for points with negative relTime (and for less-flat centroids whose first member is one): they read as
    the ring below, move between the same-scan / other-scan classes and decide where the loops break.  Found on frame 17072 of
    the whole-bag workload (returns at ~2 cm range, random azimuth).  Every association round of the single-pair path
    (voxel-hash search) and of the batch path (box index) must equal the oracle's, and so must the converged result."""
    from vil_sensor_fusion_b200 import api
    mk = scenes.vlp16_scan if lidar == "VLP-16" else scenes.hdl64_scan
    raws = [_late_first_column(mk(0.1 * k, noise=0.01, seed=k, rolling=False), rings, 40 + 7 * k) for k in range(6)]
    ocfg = orc.default_config(lidar, deskew=0)
    c0, rs0, _ = orc.organise(ocfg, raws[0])
    f0 = orc.extract(ocfg, c0, rs0)
    slot = np.searchsorted(f0["less_flat_ring_start"], np.arange(len(f0["less_flat"])), side="right") - 1
    assert np.count_nonzero(f0["less_flat"][:, 3].astype(int) != slot) > 20, "the input does not exercise the case"
    gcfg = api.default_config(lidar, deskew=0, max_scans=6, max_points=max_points)
    check = (0, 3) if lidar == "VLP-16" else (2,)
    with api.Handle(gcfg) as h:
        h.lib.vlo_set_trace(h._h, 1)
        h.upload(raws)
        h.organise()
        h.extract()
        for p in check:
            ro, n_sharp, n_flat = _oracle_pair(orc, ocfg, raws[p], raws[p + 1], None, None)
            per = 2 * n_sharp + 3 * n_flat
            n_rounds = min(5, (ro["iterations"] + 4) // 5)
            for path in ("batch", "single"):
                res = h.register_pairs(np.arange(5), np.arange(1, 6))[p] if path == "batch" else h.register_pairs([p], [p + 1])[0]
                for rnd in range(n_rounds):
                    ci, si = h.pair_correspondences(p if path == "batch" else 0, rnd, n_sharp, n_flat)
                    tr = ro["trace_idx"][rnd * per:(rnd + 1) * per]
                    np.testing.assert_array_equal(ci.ravel(), tr[:2 * n_sharp], err_msg="%s corner idx, pair %d round %d" % (path, p, rnd))
                    np.testing.assert_array_equal(si.ravel(), tr[2 * n_sharp:], err_msg="%s surf idx, pair %d round %d" % (path, p, rnd))
                _compare(ro, res, "%s %s pair %d" % (lidar, path, p))
