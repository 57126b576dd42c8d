"""K6 batched IMU preintegration vs the oracle (float64) and the reference's known-answer test."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _handle():
    from vil_sensor_fusion_b200 import api
    return api.Handle(api.default_config("VLP-16", max_scans=2, max_points=1024))


def test_reference_kat():
    """gtsam_fusion/test/UnitTests.cpp:30-66: samples (0,0),(0.1,0.1),(0.2,0.2), window [0,0.15]."""
    t = np.array([0.0, 0.1, 0.2])
    a = np.array([[0, 0, 0], [0.1, 0.1, 0.1], [0.2, 0.2, 0.2]], float)
    with _handle() as h:
        f = h.imu_preintegrate_batch(t, a, a, [0.0], [0.15])[0]
    np.testing.assert_allclose(f["dV"], 0.0175, rtol=1e-6)        # EXPECT_FLOAT_EQ
    np.testing.assert_allclose(f["dP"], 0.0011875, rtol=1e-6)
    assert abs(f["dt"] - 0.15) < 1e-15 and f["n_integrated"] == 2


def test_batch_vs_oracle(orc):
    """C4-shaped: 200 Hz jittered stream, keyframe times not aligned with samples, constant bias."""
    rng = np.random.default_rng(2)
    n = 20000
    t = np.arange(n) / 200.0 + rng.uniform(-1e-4, 1e-4, n)
    tt = np.arange(n) / 200.0
    acc = np.stack([0.5 * np.sin(0.7 * tt), 0.3 * np.cos(1.3 * tt), 9.81 + 0.2 * np.sin(2.1 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    gyro = np.stack([0.2 * np.sin(0.9 * tt), 0.1 * np.cos(0.4 * tt), 0.3 * np.sin(0.5 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    nf = 900
    t0 = 0.0317 + 0.1 * np.arange(nf)
    t1 = t0 + 0.1
    bias = np.array([1e-2, -2e-2, 1.5e-2, 1e-3, -2e-3, 3e-3])
    prm = orc.imu_params()
    fo = orc.imu_batch(prm, t, acc, gyro, t0, t1, bias)
    with _handle() as h:
        fg = h.imu_preintegrate_batch(t, acc, gyro, t0, t1, bias)
    for k in ("dR", "dP", "dV", "dR_dbg", "dP_dba", "dP_dbg", "dV_dba", "dV_dbg"):
        np.testing.assert_allclose(fg[k], fo[k], rtol=0, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(fg["cov"], fo["cov"], rtol=1e-10, atol=1e-22)
    np.testing.assert_array_equal(fg["n_integrated"], fo["n_integrated"])
    np.testing.assert_allclose(fg["dt"], fo["dt"], rtol=0, atol=1e-15)
    assert np.all(fo["n_integrated"] >= 20)


def test_window_edge_cases(orc):
    """sample exactly at t0 (dropped), exactly at t1 (interpolation factor 1), empty window,
    window past the end of the stream (no interpolation step)."""
    t = np.array([0.0, 0.005, 0.010, 0.015, 0.020, 0.025])
    rng = np.random.default_rng(5)
    acc = rng.normal(0, 1, (6, 3))
    gyro = rng.normal(0, 0.1, (6, 3))
    t0 = np.array([0.005, 0.0, 0.011, 0.020, 0.026])
    t1 = np.array([0.015, 0.010, 0.014, 0.100, 0.030])
    prm = orc.imu_params()
    fo = orc.imu_batch(prm, t, acc, gyro, t0, t1)
    with _handle() as h:
        fg = h.imu_preintegrate_batch(t, acc, gyro, t0, t1)
    np.testing.assert_array_equal(fg["n_integrated"], fo["n_integrated"])
    for k in ("dR", "dP", "dV", "cov", "dt"):
        np.testing.assert_allclose(fg[k], fo[k], rtol=1e-10, atol=1e-14, err_msg=k)


def test_full_size_10k_factors(orc):
    """BASELINE config 4 at full size: 10 000 keyframe intervals over a 200 Hz stream, one warp per factor.  Every
    100th factor is compared with the oracle; all of them through properties: deltaTij = t1 - t0, deltaRij orthonormal,
    and additivity of deltaTij over consecutive windows (the windows tile the stream)."""
    rng = np.random.default_rng(2)
    nf = 10000
    n = 200 * (nf // 10) + 400
    tt = np.arange(n) / 200.0
    t = tt + rng.uniform(-1e-4, 1e-4, n)
    acc = np.stack([0.5 * np.sin(0.7 * tt), 0.3 * np.cos(1.3 * tt), 9.81 + 0.2 * np.sin(2.1 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    gyro = np.stack([0.2 * np.sin(0.9 * tt), 0.1 * np.cos(0.4 * tt), 0.3 * np.sin(0.5 * tt)], -1) + rng.normal(0, 1e-3, (n, 3))
    t0 = 0.0317 + 0.1 * np.arange(nf)
    t1 = t0 + 0.1
    bias = np.array([1e-2, -2e-2, 1.5e-2, 1e-3, -2e-3, 3e-3])
    with _handle() as h:
        fg = h.imu_preintegrate_batch(t, acc, gyro, t0, t1, bias)
    assert len(fg) == nf
    np.testing.assert_allclose(fg["dt"], t1 - t0, rtol=0, atol=1e-12)
    R = fg["dR"].reshape(nf, 3, 3)
    np.testing.assert_allclose(R @ R.transpose(0, 2, 1), np.broadcast_to(np.eye(3), (nf, 3, 3)), atol=1e-12)
    np.testing.assert_allclose(np.linalg.det(R), 1.0, atol=1e-12)
    assert abs(fg["dt"].sum() - (t1[-1] - t0[0])) < 1e-9
    assert np.all(fg["n_integrated"] >= 20) and np.all(np.isfinite(fg["cov"]))
    sel = np.arange(0, nf, 100)
    prm = orc.imu_params()
    fo = orc.imu_batch(prm, t, acc, gyro, t0[sel], t1[sel], bias)
    for k in ("dR", "dP", "dV", "dR_dbg", "dP_dba", "dP_dbg", "dV_dba", "dV_dbg"):
        np.testing.assert_allclose(fg[k][sel], fo[k], rtol=0, atol=1e-12, err_msg=k)
    np.testing.assert_allclose(fg["cov"][sel], fo["cov"], rtol=1e-10, atol=1e-22)
