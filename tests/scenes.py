"""Shared seeded inputs for the parity tests (small enough for the oracle to finish in seconds)."""
import numpy as np

from vil_sensor_fusion_b200 import synth


def vlp16_scan(t0=0.0, scene=None, noise=0.0, seed=0, rolling=True, traj=None, n_az=1800):
    scene = scene or synth.scene_room(0)
    traj = traj or synth.Trajectory()
    return synth.make_scan(scene, "VLP-16", t0=t0, traj=traj, noise_sigma=noise, seed=seed, rolling=rolling, n_az=n_az)


def hdl64_scan(t0=0.0, scene=None, noise=0.0, seed=0, rolling=False, traj=None, n_az=1800):
    scene = scene or synth.scene_room(0)
    traj = traj or synth.Trajectory()
    return synth.make_scan(scene, "HDL-64E", t0=t0, traj=traj, noise_sigma=noise, seed=seed, rolling=rolling, n_az=n_az)


def ragged_scan(seed=3):
    """VLP-16 scan with dropped returns, NaNs, zeros and out-of-FoV points mixed in."""
    rng = np.random.default_rng(seed)
    s = vlp16_scan(noise=0.01, seed=seed)
    keep = rng.uniform(size=s.shape[0]) > 0.15
    s = s[keep].copy()
    idx = rng.choice(s.shape[0], 60, replace=False)
    s[idx[:20], 0] = np.nan
    s[idx[20:40], :3] = 0.0
    s[idx[40:], 2] = 50.0     # far above the vertical FoV
    return s
