"""Scan-to-map registration (K5: 5-NN on the voxel-hash map, PCA line / LSQ plane, GN) vs oracle."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _setup(orc, lidar, raw, n_map, max_points):
    from vil_sensor_fusion_b200 import api, synth
    scene = synth.scene_room(0)
    cm, sm = synth.sample_map_points(scene, n_map, seed=1)
    ocfg = orc.default_config(lidar, deskew=0)
    gcfg = api.default_config(lidar, deskew=0, max_scans=2, max_points=max_points, max_map_points=max(len(cm), len(sm)))
    c, rs, _ = orc.organise(ocfg, raw)
    f = orc.extract(ocfg, c, rs)
    return ocfg, gcfg, cm, sm, c[f["less_sharp_idx"]], f["less_flat"]


def test_map_knn_exact(orc):
    from vil_sensor_fusion_b200 import api, synth
    scene = synth.scene_room(0)
    cm, sm = synth.sample_map_points(scene, 60000, seed=3)
    # duplicate some points so exact ties occur: lowest index must win
    sm[1000:1100] = sm[0:100]
    rng = np.random.default_rng(0)
    q = sm[rng.choice(len(sm), 3000, replace=False)].copy()
    q[:2000, :3] += rng.normal(0, 0.15, (2000, 3)).astype(np.float32)
    q[2990:, :3] += 50.0          # nothing within range
    q[:100] = sm[0:100]           # exact duplicates -> d2 = 0 ties between i and i+1000
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=1024, max_map_points=len(sm))
    with api.Handle(gcfg) as h:
        h.map_build(cm, sm)
        for k, md in ((1, 25.0), (5, 1.0), (5, 25.0)):
            ig, dg = h.map_knn(1, q, k, md)
            io, do = orc.knn_brute(sm, q, k)
            ok = do < md
            io = np.where(ok, io, -1)
            np.testing.assert_array_equal(ig, io)
            np.testing.assert_array_equal(dg[ok].view(np.uint32), do[ok].view(np.uint32))
        ig, _ = h.map_knn(1, q[:100], 1, 25.0)
        assert np.all(ig[:, 0] == np.arange(100))


@pytest.mark.parametrize("lidar,max_points,n_map", [("VLP-16", 32768, 200000), ("HDL-64E", 131072, 400000)])
def test_register_map_bit_exact(orc, lidar, max_points, n_map):
    from vil_sensor_fusion_b200 import api, synth
    R = synth.rot_zyx(0.1, 0.0, 0.0)
    p = np.array([-3.0, 1.0, 0.2])
    raw = synth.make_scan(synth.scene_room(0), lidar, pose=(R, p), rolling=False)
    ocfg, gcfg, cm, sm, cq, sq = _setup(orc, lidar, raw, n_map, max_points)
    gt = synth.loam_map_pose(R, p).astype(np.float32)
    seed = gt + np.array([0.005, -0.008, 0.004, 0.08, -0.05, 0.1], np.float32)
    ro = orc.mapping_register(ocfg, cq, sq, cm, sm, seed, use_kdtree=True, trace=True)
    with api.Handle(gcfg) as h:
        h.map_build(cm, sm)
        h.upload([raw])
        h.organise()
        h.extract()
        cfg1 = h.cfg
        rg = h.register_map([0], [seed])[0]
    assert rg["iterations"] == ro["iterations"]
    assert rg["n_corr_edge"] == ro["n_corr_edge"] and rg["n_corr_plane"] == ro["n_corr_plane"]
    dT = np.abs(rg["transform"] - ro["transform"])
    assert np.all(dT[:3] <= 1e-5) and np.all(dT[3:] <= 1e-4), dT
    np.testing.assert_array_equal(rg["transform"].view(np.uint32), ro["transform"].view(np.uint32))
    np.testing.assert_array_equal(rg["hessian"].view(np.uint32), ro["hessian"].view(np.uint32))
    np.testing.assert_allclose(rg["eig"], ro["eig"], rtol=1e-4)
    assert bool(rg["is_degenerate"]) == ro["is_degenerate"]
    # converged to ground truth (the map is the true scene)
    assert np.all(np.abs(rg["transform"][:3] - gt[:3]) < 2e-3) and np.all(np.abs(rg["transform"][3:] - gt[3:]) < 2e-2)


def test_map_first_association_indices(orc):
    """5-NN indices of the first association (max_iterations = 1) equal the oracle's, bit-exact."""
    from vil_sensor_fusion_b200 import api, synth
    R = synth.rot_zyx(-0.2, 0.0, 0.0)
    p = np.array([2.0, -1.5, 0.1])
    raw = synth.make_scan(synth.scene_room(0), "VLP-16", pose=(R, p), rolling=False)
    ocfg, gcfg, cm, sm, cq, sq = _setup(orc, "VLP-16", raw, 150000, 32768)
    ocfg.map_max_iterations = 1
    gcfg.map_max_iterations = 1
    seed = synth.loam_map_pose(R, p).astype(np.float32) + np.array([0, 0.01, 0, 0.05, 0.0, -0.05], np.float32)
    ro = orc.mapping_register(ocfg, cq, sq, cm, sm, seed, use_kdtree=False, trace=True)
    with api.Handle(gcfg) as h:
        h.map_build(cm, sm)
        h.upload([raw])
        h.organise()
        h.extract()
        h.register_map([0], [seed])
        ci, si = h.map_correspondences(0, len(cq), len(sq))
    tr = ro["trace_idx"][:(len(cq) + len(sq)) * 5].reshape(-1, 5)
    np.testing.assert_array_equal(ci, tr[:len(cq)])
    np.testing.assert_array_equal(si, tr[len(cq):])


def test_empty_map_soft_status(orc):
    from vil_sensor_fusion_b200 import api
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=32768, max_map_points=1000)
    with api.Handle(gcfg) as h:
        h.map_build(np.zeros((5, 4), np.float32), np.zeros((50, 4), np.float32))
        h.upload([scenes.vlp16_scan(0.0)])
        h.organise()
        h.extract()
        seed = np.arange(6, dtype=np.float32) * 0.01
        r = h.register_map([0], [seed])[0]
        assert r["status"] == 1 and r["iterations"] == 0
        np.testing.assert_array_equal(r["transform"], seed)
