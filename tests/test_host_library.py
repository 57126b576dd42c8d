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
