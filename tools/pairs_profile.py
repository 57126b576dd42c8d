"""Whole-bag scan-to-scan step (127 HDL-64 pairs) a few times -- target for ncu captures of k3_assoc / k3_gn."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vil_sensor_fusion_b200 import api, synth
scene = synth.scene_room(0); traj = synth.Trajectory()
pool = [synth.make_scan(scene, "HDL-64E", t0=0.1 * k, traj=traj, rolling=False, noise_sigma=0.01, seed=k) for k in range(8)]
B = 128
order = list(range(8)) + list(range(6, 0, -1))
raws = [pool[order[k % len(order)]] for k in range(B)]
cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072)
with api.Handle(cfg) as h:
    for rep in range(3):
        h.upload(raws); h.organise(); h.extract()
        r = h.register_pairs(np.arange(B - 1), np.arange(1, B))
    print("iterations", r["iterations"].mean(), "ok", int((r["status"] == 0).sum()))
