"""Tuning helper: whole-bag scan-to-scan step time vs the cell sizes of the scan-to-scan target grids."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vil_sensor_fusion_b200 import api, synth
scene = synth.scene_room(0); traj = synth.Trajectory()
pool = [synth.make_scan(scene, "HDL-64E", t0=0.1 * k, traj=traj, rolling=False, noise_sigma=0.01, seed=k) for k in range(8)]
B = 128
order = list(range(8)) + list(range(6, 0, -1))
raws = [pool[order[k % len(order)]] for k in range(B)]
ref = None
for cc, sc in [(5.0, 0.7), (2.0, 0.7), (1.0, 0.7), (0.6, 0.7), (1.0, 0.5), (1.0, 1.0), (0.6, 0.4)]:
    cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072, odom_corner_cell_size=cc, odom_cell_size=sc)
    with api.Handle(cfg) as h:
        h.upload(raws); h.organise(); h.extract()
        for rep in range(2):
            r = h.register_pairs(np.arange(B - 1), np.arange(1, B))
        h.set_profiling(True)
        t0 = time.perf_counter()
        for rep in range(3):
            r = h.register_pairs(np.arange(B - 1), np.arange(1, B))
        dt = (time.perf_counter() - t0) / 3 * 1e3
        st = h.stage_times()
        if ref is None: ref = r["transform"].copy()
        same = np.array_equal(ref.view(np.uint32), r["transform"].view(np.uint32))
        print("corner cell %.2f surf cell %.2f: register_pairs %.3f ms  assoc %.3f gn %.3f grid %.3f  identical results %s" %
              (cc, sc, dt, st["k3_assoc"][0] / 3, st["k3_gn"][0] / 3, st["k2_grid_build"][0] / 3, same))
    # single pair latency (online shape)
    cfg1 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, odom_corner_cell_size=cc, odom_cell_size=sc)
    with api.Handle(cfg1) as h:
        h.upload(raws[:2]); h.organise(); h.extract()
        lat = []
        for rep in range(12):
            t0 = time.perf_counter(); h.register_pairs([0], [1]); lat.append((time.perf_counter() - t0) * 1e3)
        print("    single pair register_pairs p50 %.3f ms" % sorted(lat)[6])
