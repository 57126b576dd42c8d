"""Exploratory per-stage timing on the GPU box (not the bench): HDL-64 batch + 1M-point map."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vil_sensor_fusion_b200 import api, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
POOL = 8
scene = synth.scene_room(0)
traj = synth.Trajectory()
t0 = time.time()
pool = [synth.make_scan(scene, "HDL-64E", t0=0.1 * k, traj=traj, rolling=False, noise_sigma=0.01, seed=k) for k in range(POOL + 1)]
cm, sm = synth.sample_map_points(scene, 1000000, seed=1)
print("gen %.1fs  scan pts %d  map %d/%d" % (time.time() - t0, pool[0].shape[0], len(cm), len(sm)), flush=True)
seeds = []
for k in range(POOL + 1):
    gt = synth.loam_map_pose(traj.rotation(0.1 * k), traj.position(0.1 * k)).astype(np.float32)
    seeds.append(gt + np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32))
raws = [pool[k % POOL] for k in range(B)]
sd = np.stack([seeds[k % POOL] for k in range(B)])
cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072, max_map_points=max(len(cm), len(sm)))
with api.Handle(cfg) as h:
    h.map_build(cm, sm)
    h.set_profiling(True)
    for rep in range(3):
        t0 = time.time()
        h.upload(raws)
        h.organise(); h.extract()
        cnt = h.counts()
        tm = time.time()
        res = h.register_map(np.arange(B), sd)
        t1 = time.time()
        st = h.stage_times()
        print("rep %d: total %.1f ms (%.2f ms/scan), register_map %.1f ms; counts[0]=%s" % (rep, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / B, (t1 - tm) * 1e3, cnt[0]))
        print("   iters", np.bincount(res["iterations"]), "status", np.bincount(res["status"]), "corr", res["n_corr_edge"][:3], res["n_corr_plane"][:3])
        for k, (ms, n) in st.items():
            if n: print("   %-14s %8.3f ms  %4d launches  %.3f ms/launch  %.2f us/scan" % (k, ms, n, ms / n, ms * 1e3 / B))
    # scan-to-scan pairs
    for rep in range(2):
        t0 = time.time()
        res = h.register_pairs(np.arange(B - 1), np.arange(1, B))
        t1 = time.time()
        st = h.stage_times()
        print("pairs rep %d: %.1f ms (%.2f ms/pair) iters %s" % (rep, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / (B - 1), np.bincount(res["iterations"])))
        for k, (ms, n) in st.items():
            if n: print("   %-14s %8.3f ms  %4d launches  %.3f ms/launch  %.2f us/pair" % (k, ms, n, ms / n, ms * 1e3 / (B - 1)))
# online latency
cfg1 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, max_map_points=max(len(cm), len(sm)))
with api.Handle(cfg1) as h:
    h.map_build(cm, sm)
    h.online_set_map_pose(synth.loam_map_pose(traj.rotation(0.0), traj.position(0.0)).astype(np.float32))
    lat = []
    for k in range(POOL + 1):
        t0 = time.time()
        rc, o, m = h.process_scan(pool[k], 0.1 * k, want_map=True)
        lat.append((time.time() - t0) * 1e3)
    print("online tick ms:", np.round(lat, 2), "map iters", m["iterations"], "odom iters", o["iterations"])
    s, mp = h.online_pose()
    print("sum", s, "mapped", mp, "gt", synth.loam_map_pose(traj.rotation(0.1 * POOL), traj.position(0.1 * POOL)))
