"""Per-stage device time of one online tick (vlo_process_scan) on HDL-64 + 1M map."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vil_sensor_fusion_b200 import api, synth
scene = synth.scene_room(0); traj = synth.Trajectory()
pool = [synth.make_scan(scene, "HDL-64E", t0=0.1 * k, traj=traj, rolling=False, noise_sigma=0.01, seed=k) for k in range(9)]
cm, sm = synth.make_voxel_map(scene, 1000000, seed=1)
cfg = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, max_map_points=int(max(len(cm), len(sm))), io_ratio=1)
with api.Handle(cfg) as h:
    h.map_build(cm, sm)
    h.online_set_map_pose(synth.loam_map_pose(traj.rotation(0.0), traj.position(0.0)).astype(np.float32))
    for k in range(4):
        h.process_scan(pool[k], 0.1 * k, want_map=True)
    h.set_profiling(True)
    lat = []
    for k in range(4, 9):
        t0 = time.perf_counter()
        rc, o, m = h.process_scan(pool[k], 0.1 * k, want_map=True)
        lat.append((time.perf_counter() - t0) * 1e3)
    st = h.stage_times()
    print("tick wall ms", np.round(lat, 3), "odom iters", o["iterations"], "map iters", m["iterations"])
    tot = 0
    for name, (ms, n) in st.items():
        if n:
            print("  %-14s %7.3f ms/tick  %5.1f launch-groups/tick" % (name, ms / 5, n / 5)); tot += ms / 5
    print("  sum of device stage time %.3f ms/tick" % tot)
