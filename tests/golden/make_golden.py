"""Generates tests/golden/*.npz.

The reference ships no LOAM / IMU golden vectors for this path ("LOAM hot path: no test
of any kind"; sample_bags/.gitignore).  These fixtures therefore pin the ORACLE against itself across
commits (regression goldens) on seeded synthetic inputs; the only values that come from the reference
are the known-answer numbers asserted in tests/test_oracle_reference_kat.py.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc                      # noqa: E402
from vil_sensor_fusion_b200 import synth              # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    cfg = orc.default_config("VLP-16", deskew=1)
    raws = [synth.make_scan(scene, "VLP-16", t0=0.1 * k, traj=traj, n_az=450, noise_sigma=0.01, seed=k) for k in range(2)]
    out = {}
    feats = []
    for k, raw in enumerate(raws):
        c, rs, src = orc.organise(cfg, raw)
        f = orc.extract(cfg, c, rs)
        feats.append((c, f))
        out["raw%d" % k] = raw
        out["cloud%d" % k] = c
        out["ring_start%d" % k] = rs
        for name in ("label", "picked", "sharp_idx", "less_sharp_idx", "flat_idx", "less_flat", "less_flat_ring_start"):
            out["%s%d" % (name, k)] = f[name]
    (c0, f0), (c1, f1) = feats
    seed = synth.loam_sweep_transform(traj.rotation(0.0), traj.position(0.0), traj.rotation(0.1), traj.position(0.1)).astype(np.float32)
    r = orc.odometry_register(cfg, c1[f1["sharp_idx"]], c1[f1["flat_idx"]], c0[f0["less_sharp_idx"]], f0["less_sharp_ring_start"],
                              f0["less_flat"], f0["less_flat_ring_start"], seed=seed, trace=True)
    out["odom_seed"] = seed
    for name in ("transform", "hessian", "eig", "P", "cov"):
        out["odom_" + name] = r[name]
    out["odom_scalars"] = np.array([r["iterations"], r["n_corr_edge"], r["n_corr_plane"], int(r["is_degenerate"]), int(r["pass_dopt"])])
    out["odom_trace_idx0"] = r["trace_idx"][:2 * len(f1["sharp_idx"]) + 3 * len(f1["flat_idx"])]
    np.savez_compressed(os.path.join(OUT, "vlp16_pair.npz"), **out)

    # scan-to-map
    cm, sm = synth.sample_map_points(scene, 30000, seed=1)
    R = synth.rot_zyx(0.1, 0.0, 0.0)
    p = np.array([-3.0, 1.0, 0.2])
    raw = synth.make_scan(scene, "VLP-16", pose=(R, p), rolling=False, n_az=450)
    cfgm = orc.default_config("VLP-16", deskew=0)
    c, rs, _ = orc.organise(cfgm, raw)
    f = orc.extract(cfgm, c, rs)
    seedm = synth.loam_map_pose(R, p).astype(np.float32) + np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32)
    rm = orc.mapping_register(cfgm, c[f["less_sharp_idx"]], f["less_flat"], cm, sm, seedm)
    np.savez_compressed(os.path.join(OUT, "vlp16_map.npz"), raw=raw, corner_map=cm, surf_map=sm, seed=seedm,
                        transform=rm["transform"], hessian=rm["hessian"], eig=rm["eig"],
                        scalars=np.array([rm["iterations"], rm["n_corr_edge"], rm["n_corr_plane"], int(rm["is_degenerate"])]))

    # IMU: the input sequence of gtsam_fusion/test/TestTest.cpp:22-29 (10 steps, dt = 0.01, all covariances 1e-4)
    t = 0.01 * np.arange(1, 11)
    i = np.arange(10)
    acc = np.stack([0.01 * i, 0.02 * i, 0.03 * i + 9.81], -1)
    gyro = np.stack([0.004 * i, 0.005 * i, 0.006 * i], -1)
    prm = orc.imu_params(1e-4, 1e-4, 1e-4, 1e-4, 1e-4, 1e-4)
    # window (0, 0.1+]: every sample integrated over dt = 0.01 (the last through the t1 > t_last rule)
    fct = orc.imu_get_factor(prm, t, acc, gyro, 0.0, 0.1000001)
    np.savez_compressed(os.path.join(OUT, "imu_testtest.npz"), t=t, acc=acc, gyro=gyro, dR=fct["dR"], dP=fct["dP"], dV=fct["dV"],
                        cov=fct["cov"], dt=fct["dt"], n=fct["n_integrated"], dP_dba=fct["dP_dba"], dV_dbg=fct["dV_dbg"])
    # LaserMapping with its map maintenance: 4 ticks of a VLP-16 sequence (450 azimuth columns), each seeded with the
    # previous mapped pose; the fixture keeps the inputs of tick 0 (to regenerate the others from the seed) and, per
    # tick, pose / sizes / a checksum of the map
    cfgl = orc.default_config("VLP-16", deskew=0)
    lm = orc.LaserMap(cfgl, cap=100000)
    seed_l = np.zeros(6, np.float32)
    poses, infos, sums = [], [], []
    for k in range(4):
        raw = synth.make_scan(scene, "VLP-16", t0=0.1 * k, traj=traj, rolling=False, n_az=450)
        c, rs, _ = orc.organise(cfgl, raw)
        f = orc.extract(cfgl, c, rs)
        r = lm.process(c[f["less_sharp_idx"]], f["less_flat"], seed_l)
        seed_l = r["transform"]
        poses.append(r["transform"])
        infos.append(list(r["info"]["n_ds"]) + list(r["info"]["n_sub"]) + list(r["info"]["n_map"]) + [r["iterations"], r["status"]])
        sums.append([float(lm.points(w)[0][:, :3].astype(np.float64).sum()) for w in range(2)])
    pc, cc = lm.points(0)
    np.savez_compressed(os.path.join(OUT, "vlp16_lasermap.npz"), poses=np.stack(poses), infos=np.array(infos), sums=np.array(sums),
                        corner_map=pc, corner_cube=cc, surf_map_head=lm.points(1)[0][:256])
    lm.close()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
