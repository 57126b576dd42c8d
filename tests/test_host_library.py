"""The whole library on the CPU: every .cu of csrc/ (kernels AND host code) compiled against the SIMT emulator and the fake
CUDA runtime of tests/host/ into libvlo_emul.so, driven through the real C-ABI and the real api.py, checked against the
oracle bit for bit.  One host thread per CUDA thread, CTAs one after the other, so sizes are small -- the -m gpu suite
repeats all of this (and much more) on the B200.  TEST INFRASTRUCTURE: the product never loads this library."""
import os
import sys

import numpy as np
import pytest

from tests import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "host"))


@pytest.fixture(scope="module")
def emulated():
    import build_emul
    from vil_sensor_fusion_b200 import _lib
    path = build_emul.build()
    if path is None:
        pytest.skip("CUDA headers not found")
    saved = _lib._lib
    _lib._lib = _lib.load(path)               # api.Handle now talks to the emulated library
    try:
        from vil_sensor_fusion_b200 import api
        yield api
    finally:
        _lib._lib = saved


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_emulated_scan_to_scan_registration_equals_oracle(emulated, orc):
    """K0 + K1 + K2 (scan grids) + K3 (warp-per-query association, fused GN, on-device QR solve + Jacobi degeneracy)."""
    api = emulated
    raws = [scenes.vlp16_scan(0.1 * k, noise=0.01, seed=k, rolling=False, n_az=360) for k in range(2)]
    ocfg = orc.default_config("VLP-16", deskew=0)
    gcfg = api.default_config("VLP-16", deskew=0, max_scans=2, max_points=8192)
    with api.Handle(gcfg) as h:
        h.lib.vlo_set_trace(h._h, 1)
        h.upload(raws)
        h.organise()
        h.extract()
        rg = h.register_pairs([0], [1])[0]
        feats = [h.get_features(i) for i in range(2)]
        n_sharp, n_flat = len(feats[1]["sharp_idx"]), len(feats[1]["flat_idx"])
        ci, si = h.pair_correspondences(0, 0, n_sharp, n_flat)
    clouds = [orc.organise(ocfg, r) for r in raws]
    fo = [orc.extract(ocfg, c, rs) for c, rs, _ in clouds]
    for fg, f in zip(feats, fo):
        np.testing.assert_array_equal(fg["label"], f["label"])
        np.testing.assert_array_equal(_bits(fg["less_flat"]), _bits(f["less_flat"]))
    c0, c1 = clouds[0][0], clouds[1][0]
    ro = orc.odometry_register(ocfg, c1[fo[1]["sharp_idx"]], c1[fo[1]["flat_idx"]], c0[fo[0]["less_sharp_idx"]],
                               fo[0]["less_sharp_ring_start"], fo[0]["less_flat"], fo[0]["less_flat_ring_start"], use_kdtree=True, trace=True)
    np.testing.assert_array_equal(ci.ravel(), ro["trace_idx"][:2 * n_sharp])
    np.testing.assert_array_equal(si.ravel(), ro["trace_idx"][2 * n_sharp:2 * n_sharp + 3 * n_flat])
    assert rg["iterations"] == ro["iterations"] and rg["n_corr_edge"] == ro["n_corr_edge"] and rg["n_corr_plane"] == ro["n_corr_plane"]
    np.testing.assert_array_equal(_bits(rg["transform"]), _bits(ro["transform"]))
    np.testing.assert_array_equal(_bits(rg["hessian"]), _bits(ro["hessian"]))
    np.testing.assert_array_equal(_bits(rg["P"]), _bits(ro["P"]))
    np.testing.assert_allclose(rg["eig"], ro["eig"], rtol=1e-4)
    assert bool(rg["is_degenerate"]) == ro["is_degenerate"] and bool(rg["pass_dopt"]) == ro["pass_dopt"]


def test_emulated_scan_to_map_registration_equals_oracle(emulated, orc):
    """K7 stack VoxelGrid + K2 map grids + K5: the cooperative single-launch kernel (one scan) and the batch path
    (k5_assoc / k5_lin over the ticketed tile list with the warp-level solve, five scans) against the oracle."""
    from vil_sensor_fusion_b200 import synth
    api = emulated
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    cm, sm = synth.sample_map_points(scene, 16000, seed=1)
    raws, seeds = [], []
    for k in range(5):
        t = 0.1 * k
        raws.append(synth.make_scan(scene, "VLP-16", t0=t, traj=traj, rolling=False, noise_sigma=0.01, seed=k, n_az=300))
        seeds.append(synth.loam_map_pose(traj.rotation(t), traj.position(t)).astype(np.float32)
                     + np.array([0.004, -0.006, 0.003, 0.05, -0.04, 0.06], np.float32))
    seeds = np.stack(seeds)
    ocfg = orc.default_config("VLP-16", deskew=0)
    gcfg = api.default_config("VLP-16", deskew=0, max_scans=5, max_points=8192, max_map_points=int(max(len(cm), len(sm))))
    with api.Handle(gcfg) as h:
        h.map_build(cm, sm)
        h.upload(raws)
        h.organise()
        h.extract()
        single = h.register_map([0], [seeds[0]])[0]
        batch = h.register_map(np.arange(5), seeds)
        stacks = [h.get_stack(k) for k in (0, 3)]
    for (cq, sq), k in zip(stacks, (0, 3)):
        c, rs, _ = orc.organise(ocfg, raws[k])
        f = orc.extract(ocfg, c, rs)
        np.testing.assert_array_equal(_bits(cq), _bits(orc.voxel_downsample(c[f["less_sharp_idx"]], ocfg.corner_filter_size)))
        np.testing.assert_array_equal(_bits(sq), _bits(orc.voxel_downsample(f["less_flat"], ocfg.surface_filter_size)))
        ro = orc.mapping_register(ocfg, cq, sq, cm, sm, seeds[k], use_kdtree=True)
        for r in ([single, batch[0]] if k == 0 else [batch[k]]):
            assert r["iterations"] == ro["iterations"] and r["n_corr_plane"] == ro["n_corr_plane"] and r["n_corr_edge"] == ro["n_corr_edge"]
            np.testing.assert_array_equal(_bits(r["transform"]), _bits(ro["transform"]))
            np.testing.assert_array_equal(_bits(r["hessian"]), _bits(ro["hessian"]))
            np.testing.assert_allclose(r["eig"], ro["eig"], rtol=1e-4)
    assert np.all(batch["status"] == 0)


def test_emulated_imu_preintegration_equals_oracle(emulated, orc):
    """K6 (one warp per factor, FP64, 15x15 covariance products through shared memory) incl. the reference's KAT."""
    api = emulated
    rng = np.random.default_rng(2)
    n = 1200
    t = np.arange(n) / 200.0 + rng.uniform(-1e-4, 1e-4, n)
    tt = np.arange(n) / 200.0
    acc = np.stack([0.5 * np.sin(0.7 * tt), 0.3 * np.cos(1.3 * tt), 9.81 + 0.2 * np.sin(2.1 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    gyro = np.stack([0.2 * np.sin(0.9 * tt), 0.1 * np.cos(0.4 * tt), 0.3 * np.sin(0.5 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    t0 = 0.0317 + 0.1 * np.arange(50)
    t1 = t0 + 0.1
    bias = np.array([1e-2, -2e-2, 1.5e-2, 1e-3, -2e-3, 3e-3])
    fo = orc.imu_batch(orc.imu_params(), t, acc, gyro, t0, t1, bias)
    with api.Handle(api.default_config("VLP-16", max_scans=2, max_points=1024)) as h:
        fg = h.imu_preintegrate_batch(t, acc, gyro, t0, t1, bias)
        kat_t = np.array([0.0, 0.1, 0.2])
        kat_a = np.array([[0, 0, 0], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2]], float)
        kat = h.imu_preintegrate_batch(kat_t, kat_a, kat_a, [0.0], [0.15])[0]
    for k in ("dR", "dP", "dV", "dR_dbg", "dP_dba", "dP_dbg", "dV_dba", "dV_dbg"):
        np.testing.assert_allclose(fg[k], fo[k], rtol=0, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(fg["cov"], fo["cov"], rtol=1e-10, atol=1e-22)
    np.testing.assert_array_equal(fg["n_integrated"], fo["n_integrated"])
    np.testing.assert_allclose(kat["dV"], 0.0175, rtol=1e-6)              # gtsam_fusion/test/UnitTests.cpp:61-66
    np.testing.assert_allclose(kat["dP"], 0.0011875, rtol=1e-6)


def test_emulated_batch_pairs_thread_search_equals_single_pair_path(emulated, orc):
    """More than four pairs per call: k3_assoc_thread (thread-per-query 27-cell search with the warp-cooperative fallback;
    0.35 m cells force the fallback for most partner searches) against the warp-per-query kernel, and pair 2 against the
    oracle."""
    api = emulated
    raws = [scenes.vlp16_scan(0.1 * k, noise=0.01, seed=k, rolling=False, n_az=200) for k in range(6)]
    ocfg = orc.default_config("VLP-16", deskew=0)
    gcfg = api.default_config("VLP-16", deskew=0, max_scans=6, max_points=4096, odom_cell_size=0.35)
    with api.Handle(gcfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        batch = h.register_pairs(np.arange(5), np.arange(1, 6))
        singles = [h.register_pairs([p], [p + 1])[0] for p in (0, 2, 4)]
    for s, p in zip(singles, (0, 2, 4)):
        for f in ("transform", "hessian", "eig", "P"):
            np.testing.assert_array_equal(_bits(batch[f][p]), _bits(s[f]), err_msg="pair %d %s" % (p, f))
        assert batch["iterations"][p] == s["iterations"] and batch["n_corr_plane"][p] == s["n_corr_plane"]
    c = [orc.organise(ocfg, raws[k]) for k in (2, 3)]
    f = [orc.extract(ocfg, cc, rs) for cc, rs, _ in c]
    ro = orc.odometry_register(ocfg, c[1][0][f[1]["sharp_idx"]], c[1][0][f[1]["flat_idx"]], c[0][0][f[0]["less_sharp_idx"]],
                               f[0]["less_sharp_ring_start"], f[0]["less_flat"], f[0]["less_flat_ring_start"], use_kdtree=True)
    np.testing.assert_array_equal(_bits(batch["transform"][2]), _bits(ro["transform"]))
    np.testing.assert_array_equal(_bits(batch["hessian"][2]), _bits(ro["hessian"]))


def test_emulated_online_ticks_with_maintained_map(emulated, orc):
    """vlo_process_scan with the maintained map (K7: stack filter, cube window, sub-map, insertion + re-filtering): three
    ticks of a rolling-shutter sequence equal the oracle chained the same way (tests/test_gpu_online.py runs six full-size
    ticks on the GPU; `pytest tests/test_gpu_online.py -m gpu --emulated` runs those here, in two minutes)."""
    from tests.test_gpu_online import _euler_to_M, _M_to_euler
    api = emulated
    ocfg = orc.default_config("VLP-16", deskew=1, io_ratio=1)
    gcfg = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=8192, max_map_points=60000, io_ratio=1)
    raws = [scenes.vlp16_scan(0.1 * k, n_az=360) for k in range(3)]
    om = orc.LaserMap(ocfg, cap=60000)
    T_prev = np.zeros(6, np.float32)
    sum_o, aft, bef = np.zeros(6, np.float32), np.zeros(6, np.float32), np.zeros(6, np.float32)
    prev = None
    with api.Handle(gcfg) as h:
        h.map_reset()
        for k, raw in enumerate(raws):
            rc, odom, mapped = h.process_scan(raw, stamp=0.1 * k, want_map=True)
            c, rs, _ = orc.organise(ocfg, raw)
            f = orc.extract(ocfg, c, rs)
            lc, ls = c[f["less_sharp_idx"]], f["less_flat"]
            if k >= 1:
                ro = orc.odometry_register(ocfg, c[f["sharp_idx"]], c[f["flat_idx"]], prev[0], prev[1], prev[2], prev[3], seed=T_prev)
                np.testing.assert_array_equal(_bits(odom["transform"]), _bits(ro["transform"]), err_msg="tick %d" % k)
                T_prev = ro["transform"]
                sum_o = orc.accumulate_pose(sum_o, T_prev)
                lc = orc.transform_to_end(ocfg, T_prev, lc)
                ls = orc.transform_to_end(ocfg, T_prev, ls)
            prev = (lc, f["less_sharp_ring_start"], ls, f["less_flat_ring_start"])
            seed = _M_to_euler(_euler_to_M(aft) @ np.linalg.inv(_euler_to_M(bef)) @ _euler_to_M(sum_o))
            rm = om.process(lc, ls, seed)
            assert mapped["status"] == rm["status"], k
            dT = np.abs(mapped["transform"] - rm["transform"])
            assert np.all(dT[:3] <= 1e-5) and np.all(dT[3:] <= 1e-4), (k, dT)
            assert h.map_size() == (om.size(0), om.size(1)), k
            aft, bef = rm["transform"], sum_o.copy()
    om.close()
