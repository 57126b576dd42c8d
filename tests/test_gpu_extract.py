"""Parity of K0 (organise) and K1 (feature extraction) with the oracle: bit-exact."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu


def _cfgs(orc, lidar, **kw):
    from vil_sensor_fusion_b200 import api
    ocfg = orc.default_config(lidar)
    gcfg = api.default_config(lidar, **kw)
    return ocfg, gcfg


def _check_scan(orc, h, ocfg, scan_idx, raw):
    cloud_o, rs_o, src_o = orc.organise(ocfg, raw)
    cloud_g, rs_g, src_g = h.get_cloud(scan_idx)
    assert cloud_g.shape == cloud_o.shape
    np.testing.assert_array_equal(rs_g, rs_o)
    np.testing.assert_array_equal(src_g, src_o)
    np.testing.assert_array_equal(cloud_g.view(np.uint32), cloud_o.view(np.uint32))
    fo = orc.extract(ocfg, cloud_o, rs_o)
    fg = h.get_features(scan_idx)
    np.testing.assert_array_equal(fg["curvature"].view(np.uint32), fo["curvature"].view(np.uint32))
    np.testing.assert_array_equal(fg["picked"], fo["picked"])
    np.testing.assert_array_equal(fg["label"], fo["label"])
    for k in ("sharp_idx", "less_sharp_idx", "flat_idx", "less_sharp_ring_start", "less_flat_ring_start"):
        np.testing.assert_array_equal(fg[k], fo[k], err_msg=k)
    assert fg["less_flat"].shape == fo["less_flat"].shape
    np.testing.assert_array_equal(fg["less_flat"].view(np.uint32), fo["less_flat"].view(np.uint32))
    return fo


def test_vlp16_batch_bit_exact(orc):
    from vil_sensor_fusion_b200 import api
    ocfg, gcfg = _cfgs(orc, "VLP-16", max_scans=4, max_points=32768)
    raws = [scenes.vlp16_scan(0.0), scenes.vlp16_scan(0.1, noise=0.02, seed=1), scenes.ragged_scan(),
            scenes.vlp16_scan(0.3, rolling=False)]
    with api.Handle(gcfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        for i, raw in enumerate(raws):
            fo = _check_scan(orc, h, ocfg, i, raw)
            assert len(fo["sharp_idx"]) > 20 and len(fo["flat_idx"]) > 100


def test_hdl64_bit_exact(orc):
    from vil_sensor_fusion_b200 import api
    ocfg, gcfg = _cfgs(orc, "HDL-64E", max_scans=2, max_points=131072)
    raws = [scenes.hdl64_scan(0.0), scenes.hdl64_scan(0.1, noise=0.02, seed=5)]
    with api.Handle(gcfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        for i, raw in enumerate(raws):
            _check_scan(orc, h, ocfg, i, raw)


def test_edge_cases(orc):
    """empty scan, tiny scan (rings shorter than 2K+1), single ring, corridor with missing returns."""
    from vil_sensor_fusion_b200 import api, synth
    ocfg, gcfg = _cfgs(orc, "VLP-16", max_scans=4, max_points=32768)
    full = scenes.vlp16_scan(0.0)
    tiny = full[:100].copy()
    empty = np.zeros((0, 4), np.float32)
    corridor = synth.make_scan(synth.scene_corridor(), "VLP-16", pose=(np.eye(3), np.zeros(3)), rolling=False)
    short_rings = scenes.vlp16_scan(0.0, n_az=40)      # 40 points per ring: sequential sector path
    raws = [tiny, empty, corridor, short_rings]
    with api.Handle(gcfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        for i, raw in enumerate(raws):
            if raw.shape[0] == 0:
                c = h.counts()[i]
                assert c["n_valid"] == 0 and c["n_sharp"] == 0 and c["n_less_flat"] == 0
                continue
            _check_scan(orc, h, ocfg, i, raw)


def test_ring_capacity_error(orc):
    from vil_sensor_fusion_b200 import api
    gcfg = api.default_config("VLP-16", max_scans=1, max_points=32768, max_ring_points=1024)
    with api.Handle(gcfg) as h:
        h.upload([scenes.vlp16_scan(0.0)])
        h.organise()
        h.extract()
        with pytest.raises(api.VloError) as e:
            h.synchronize()
        assert e.value.code == -3
        # reported once: the handle stays usable (the reference LOAM has no per-ring cap and carries on with the next sweep)
        h.synchronize()
        short = scenes.vlp16_scan(0.1, n_az=900)                # 900 points per ring: fits
        h.upload([short])
        h.organise()
        h.extract()
        h.synchronize()
        ocfg = orc.default_config("VLP-16")
        c, rs, _ = orc.organise(ocfg, short)
        f = orc.extract(ocfg, c, rs)
        np.testing.assert_array_equal(h.get_features(0)["label"], f["label"])


@pytest.mark.parametrize("lidar", ["HDL-32", "O1-16", "O1-64", "Bperl-32"])
def test_other_lidar_presets_bit_exact(orc, lidar):
    """Every `lidar` preset of loam_params.yaml:22: ring assignment, rel-time and features equal the oracle."""
    from vil_sensor_fusion_b200 import api, synth
    ocfg, gcfg = _cfgs(orc, lidar, max_scans=2, max_points=131072)
    scene = synth.scene_room(0)
    raws = [synth.make_scan(scene, lidar, t0=0.1 * k, traj=synth.Trajectory(), rolling=bool(k)) for k in range(2)]
    with api.Handle(gcfg) as h:
        h.upload(raws)
        h.organise()
        h.extract()
        for i, raw in enumerate(raws):
            fo = _check_scan(orc, h, ocfg, i, raw)
            assert len(fo["flat_idx"]) > 50


def test_pointcloud2_layouts(orc):
    """sensor_msgs/PointCloud2 payloads with other field layouts (CARLA: 3 floats; a driver with x y z at offsets
    4 / 8 / 16 of a 32-byte point) give exactly the organised cloud of the plain xyzi payload."""
    from vil_sensor_fusion_b200 import api
    ocfg, gcfg = _cfgs(orc, "VLP-16", max_scans=2, max_points=65536)      # staging holds max_points * 16 bytes per scan
    raws = [scenes.vlp16_scan(0.0), scenes.ragged_scan()]
    rng = np.random.default_rng(0)
    with api.Handle(gcfg) as h:
        for step, (xo, yo, zo) in ((12, (0, 4, 8)), (32, (4, 8, 16)), (20, (8, 0, 16))):
            msgs = []
            for raw in raws:
                buf = rng.normal(0, 50, (raw.shape[0], step // 4)).astype(np.float32)      # junk in the other fields
                buf[:, xo // 4], buf[:, yo // 4], buf[:, zo // 4] = raw[:, 0], raw[:, 1], raw[:, 2]
                msgs.append(dict(data=buf.tobytes(), point_step=step, fields=dict(x=xo, y=yo, z=zo)))
            h.upload_pointcloud2(msgs)
            h.organise()
            h.extract()
            for i, raw in enumerate(raws):
                _check_scan(orc, h, ocfg, i, raw)
        bad = dict(data=np.zeros(30, np.uint8).tobytes(), point_step=10, fields=dict(x=0, y=4, z=8))
        with pytest.raises(api.VloError):
            h.upload_pointcloud2([bad])


@pytest.mark.parametrize("opt", ["rotate", "ring_field", "both", "uint16_ring"])
def test_input_rotation_and_ring_field_options(orc, opt):
    """rotateInputCloud / inputCloudRotation and useCloudIntensityandRingFields (loam_params.yaml:4-5,23): organise and the
    features that follow equal the oracle's bit for bit; the rotation really is Rz(yaw) Ry(pitch) Rx(roll) in the ROS frame;
    ring ids are taken from the field (reversed order here, with out-of-range / NaN entries that must be dropped)."""
    from vil_sensor_fusion_b200 import api
    base = scenes.vlp16_scan(0.0, noise=0.01, seed=3)
    ocfg0 = orc.default_config("VLP-16")
    cloud0, _, src0 = orc.organise(ocfg0, base)
    ring_by_src = np.full(len(base), -1.0, np.float32)
    ring_by_src[src0] = np.floor(cloud0[:, 3])
    raw = np.concatenate([base, (15.0 - ring_by_src)[:, None].astype(np.float32)], axis=1)     # stride 5, ring column reversed
    raw[ring_by_src < 0, 4] = -1.0
    raw[10:5000:97, 4] = np.nan
    raw[20:5000:89, 4] = 16.0
    kw = {}
    if opt == "uint16_ring":
        # velodyne-style 24-byte point: x y z intensity (float32), ring (uint16) at byte 18 (not float-aligned), padding;
        # dropped points carry an out-of-range id
        rec = np.zeros((len(base), 24), np.uint8)
        rec[:, :16] = base.view(np.uint8).reshape(len(base), 16)
        ids = np.where(ring_by_src >= 0, 15 - ring_by_src, 200).astype(np.uint16)
        rec[:, 18:20] = ids.view(np.uint8).reshape(-1, 2)
        raw = np.ascontiguousarray(rec).view(np.float32).reshape(len(base), 6)
        kw.update(ring_field=18, ring_field_type=1)
    if opt in ("rotate", "both"):
        kw.update(rotate_input=1, input_rotation=(0.3, -0.1, 0.05))
    if opt in ("ring_field", "both"):
        kw.update(ring_field=4)
    ocfg = orc.default_config("VLP-16", **kw)
    gcfg = api.default_config("VLP-16", max_scans=2, max_points=65536, **kw)      # 5 floats per point: staging is sized for 4
    with api.Handle(gcfg) as h:
        h.upload([raw, raw[::-1].copy()])
        h.organise()
        h.extract()
        fo = _check_scan(orc, h, ocfg, 0, raw)
        _check_scan(orc, h, ocfg, 1, raw[::-1].copy())
        cloud_g, rs_g, src_g = h.get_cloud(0)
    assert len(fo["sharp_idx"]) > 20 and len(fo["flat_idx"]) > 100
    if opt in ("rotate", "both"):
        y, p, r = 0.3, -0.1, 0.05
        Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(r), -np.sin(r)], [0, np.sin(r), np.cos(r)]])
        ros = (Rz @ Ry @ Rx @ raw[src_g, :3].astype(np.float64).T).T
        np.testing.assert_allclose(cloud_g[:, :3], ros[:, [1, 2, 0]], atol=2e-5)           # LOAM x y z = ROS y z x
    if opt == "uint16_ring":
        np.testing.assert_array_equal(np.floor(cloud_g[:, 3]), 15 - ring_by_src[src_g])
    if opt in ("ring_field", "both"):
        np.testing.assert_array_equal(np.floor(cloud_g[:, 3]), raw[src_g, 4])                # ring id = the field's value
        assert not np.any(np.isnan(raw[src_g, 4])) and np.all(raw[src_g, 4] < 16)
