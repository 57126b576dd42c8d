"""Whole-bag streaming (vlo_bag_register_*): double-buffered upload overlapped with the kernels must give
exactly the records the resident-batch calls give for the same scans (which the other suites pin to the oracle)."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _concat(scans):
    offs = np.zeros(len(scans) + 1, np.int32)
    offs[1:] = np.cumsum([s.shape[0] for s in scans])
    return np.concatenate(scans, axis=0), offs


def _same(a, b):
    for f in ("transform", "hessian", "eig", "P"):
        np.testing.assert_array_equal(a[f].view(np.uint32), b[f].view(np.uint32), err_msg=f)
    for f in ("iterations", "is_degenerate", "n_corr_edge", "n_corr_plane", "status", "pass_dopt"):
        np.testing.assert_array_equal(a[f], b[f], err_msg=f)
    np.testing.assert_allclose(a["cov"], b["cov"], rtol=1e-12, atol=0)


def test_bag_pairs_equals_resident_pairs():
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.1 * k, rolling=False, n_az=900) for k in range(7)]
    raws[3] = raws[3][: raws[3].shape[0] // 2]           # ragged: a short scan in the middle
    cfg = api.default_config("VLP-16", deskew=0, max_scans=8, max_points=16384)
    with api.Handle(cfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        ref = h.register_pairs(np.arange(6), np.arange(1, 7))
    cfg = api.default_config("VLP-16", deskew=0, max_scans=6, max_points=16384)     # halves of 3 scans
    with api.Handle(cfg) as h:
        batches = []
        for lo in (0, 2, 4):                              # batches overlap by one frame: pairs (0,1)(1,2) | (2,3)(3,4) | (4,5)(5,6)
            raw, offs = _concat(raws[lo:lo + 3])
            batches.append((raw, offs, None))
        got = h.bag_register_pairs(batches)
        assert len(got) == 6
        _same(got, ref)
        # a second pass over the same handle (events / halves reused) gives the same answer
        _same(h.bag_register_pairs(batches[::-1])[[4, 5, 2, 3, 0, 1]], ref)


def test_bag_map_equals_resident_map():
    from vil_sensor_fusion_b200 import api, synth
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    cm, sm = synth.sample_map_points(scene, 150000, seed=1)
    raws, seeds = [], []
    for k in range(5):
        t = 0.1 * k
        raws.append(synth.make_scan(scene, "VLP-16", t0=t, traj=traj, rolling=False, n_az=900))
        gt = synth.loam_map_pose(traj.rotation(t), traj.position(t)).astype(np.float32)
        seeds.append(gt + np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32))
    seeds = np.stack(seeds)
    cfg = api.default_config("VLP-16", deskew=0, max_scans=5, max_points=16384, max_map_points=int(max(len(cm), len(sm))))
    with api.Handle(cfg) as h:
        h.map_build(cm, sm)
        h.upload(raws)
        h.organise()
        h.extract()
        ref = h.register_map(np.arange(5), seeds)
        assert np.all(ref["status"] == 0)
    cfg = api.default_config("VLP-16", deskew=0, max_scans=4, max_points=16384, max_map_points=int(max(len(cm), len(sm))))
    with api.Handle(cfg) as h:
        h.map_build(cm, sm)
        batches = []
        for lo, hi in ((0, 2), (2, 4), (4, 5)):
            raw, offs = _concat(raws[lo:hi])
            batches.append((raw, offs, seeds[lo:hi]))
        got = h.bag_register_map(batches)
        _same(got, ref)
        # a second call on the same handle (staging kept between calls), clouds of 3 floats per point (CARLA's /lidar layout)
        got3 = h.bag_register_map([(np.ascontiguousarray(r[:, :3]), o, s_) for r, o, s_ in batches], stride=3)
        _same(got3, ref)
        # clouds already resident in device memory: the same call, the same records
        try:
            import torch
            on_gpu = torch.cuda.is_available()
        except ImportError:
            on_gpu = False
        if on_gpu:
            keep = [torch.from_numpy(np.ascontiguousarray(r)).cuda() for r, _, _ in batches]
            got_d = h.bag_register_map([(int(t.data_ptr()), o, s_) for t, (_, o, s_) in zip(keep, batches)])
        else:       # the CPU-emulated library: "device" memory is host memory
            keep = [np.ascontiguousarray(r, np.float32) for r, _, _ in batches]
            got_d = h.bag_register_map([(int(a.ctypes.data), o, s_) for a, (_, o, s_) in zip(keep, batches)])
        _same(got_d, ref)


def test_bag_capacity_errors():
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.0, rolling=False, n_az=450) for _ in range(3)]
    cfg = api.default_config("VLP-16", deskew=0, max_scans=4, max_points=16384)
    with api.Handle(cfg) as h:
        raw, offs = _concat(raws)                         # 3 scans > max_scans/2
        with pytest.raises(api.VloError) as e:
            h.bag_register_pairs([(raw, offs, None)])
        assert e.value.code == -3
        with pytest.raises(api.VloError) as e:            # no map resident in this handle
            h.bag_register_map([(raw[: offs[1]], offs[:2], np.zeros((1, 6), np.float32))])
        assert e.value.code == -5


def test_reprocess_pairs_from_a_ros_bag(tmp_path):
    """A bag of PointCloud2 sweeps (5 floats per point, read without ROS) reprocessed from the file equals the same
    sweeps handed over as arrays."""
    from vil_sensor_fusion_b200 import api, bag, rosbag_io as rb
    raws = [scenes.vlp16_scan(0.1 * k, rolling=False) for k in range(5)]
    msgs = []
    for k, r in enumerate(raws):
        c5 = np.concatenate([r, np.full((len(r), 1), 7.0, np.float32)], axis=1)
        msgs.append(("/lidar", "sensor_msgs/PointCloud2", rb.POINTCLOUD2_MD5, 100.0 + 0.1 * k,
                     rb.make_pointcloud2(c5, 100.0 + 0.1 * k, field_names=("x", "y", "z", "intensity", "ring"), seq=k)))
    path = str(tmp_path / "sweeps.bag")
    rb.write_bag(path, msgs, compression="bz2", chunk_messages=2)
    gcfg = api.default_config("VLP-16", deskew=0, max_scans=4, max_points=32768 + 8192)
    with api.Handle(gcfg) as h:
        res_bag, stamps = bag.reprocess_pairs_from_bag(h, path, "/lidar", batch=3)
        res_arr = bag.reprocess_pairs(h, lambda k: raws[k], 0, 5, batch=3)
    assert len(res_bag) == 4 and np.all(res_bag["status"] == 0)
    np.testing.assert_allclose(stamps, 100.0 + 0.1 * np.arange(5), atol=1e-9)
    np.testing.assert_array_equal(res_bag["transform"].view(np.uint32), res_arr["transform"].view(np.uint32))
    np.testing.assert_array_equal(res_bag["hessian"].view(np.uint32), res_arr["hessian"].view(np.uint32))


def test_register_map_enqueue_equals_register_map():
    """vlo_register_map_enqueue + vlo_synchronize + vlo_results_finish (no host synchronisation between batches) gives the records
    vlo_register_map gives, also when a second batch is enqueued before the first one's records are read"""
    from vil_sensor_fusion_b200 import api, synth
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    cm, sm = synth.sample_map_points(scene, 150000, seed=1)
    raws, seeds = [], []
    for k in range(4):
        t = 0.1 * k
        raws.append(synth.make_scan(scene, "VLP-16", t0=t, traj=traj, rolling=False, n_az=900))
        gt = synth.loam_map_pose(traj.rotation(t), traj.position(t)).astype(np.float32)
        seeds.append(gt + np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32))
    seeds = np.stack(seeds)
    cfg = api.default_config("VLP-16", deskew=0, max_scans=2, max_points=16384, max_map_points=int(max(len(cm), len(sm))))
    with api.Handle(cfg) as h:
        h.map_build(cm, sm)
        ref = []
        for lo in (0, 2):
            h.upload(raws[lo:lo + 2]); h.organise(); h.extract()
            ref.append(h.register_map(np.arange(2), seeds[lo:lo + 2]))
        ref = np.concatenate(ref)
        out = np.zeros(4, api.RESULT_DTYPE)
        for i, lo in enumerate((0, 2)):
            h.upload(raws[lo:lo + 2]); h.organise(); h.extract()
            h.register_map_enqueue(np.arange(2), seeds[lo:lo + 2], out.ctypes.data + i * 2 * api.RESULT_DTYPE.itemsize)
        h.synchronize()
        _same(h.results_finish(out), ref)


def test_register_pairs_enqueue_equals_register_pairs():
    """vlo_register_pairs_enqueue + vlo_synchronize + vlo_results_finish = vlo_register_pairs, record for record"""
    from vil_sensor_fusion_b200 import api
    raws = [scenes.vlp16_scan(0.1 * k, noise=0.01, seed=k, rolling=False, n_az=900) for k in range(6)]
    cfg = api.default_config("VLP-16", deskew=0, max_scans=6, max_points=16384)
    with api.Handle(cfg) as h:
        h.upload(raws); h.organise(); h.extract()
        ref = h.register_pairs(np.arange(5), np.arange(1, 6))
        out = np.zeros(5, api.RESULT_DTYPE)
        h.register_pairs_enqueue(np.arange(5), np.arange(1, 6), out.ctypes.data)
        h.synchronize()
        _same(h.results_finish(out), ref)
