"""Per-tick wall time and stage times of the online HDL-64 tick (static 1M map), device-synthesised consecutive sweeps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vil_sensor_fusion_b200 import api, synth, synth_gpu
scene = synth_gpu.make_scene(synth.scene_room(0)); sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
npts = sensor.rings * sensor.n_az; SEQ = 24
buf = torch.empty((SEQ, npts, 4), dtype=torch.float32, device="cuda")
synth_gpu.synth_scans(scene, sensor, 0, SEQ, 99, buf.data_ptr()); torch.cuda.synchronize()
pin = torch.empty((SEQ, npts, 4), dtype=torch.float32).pin_memory(); pin.copy_(buf); seq = pin.numpy()
Rs, ps = synth_gpu.poses(sensor, 0, SEQ)
pose0 = synth.loam_map_pose(Rs[0], ps[0]).astype(np.float32)
cm, sm = synth.make_voxel_map(synth.scene_room(0), 1000000, seed=1)
maintained = os.environ.get("MAINTAINED", "0") == "1"
cfg = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, io_ratio=1, max_map_points=(1 << 20) if maintained else int(max(len(cm), len(sm))))
with api.Handle(cfg) as h:
    if maintained:
        h.map_reset(); h.map_insert(cm, sm, np.zeros(6, np.float32))
    else:
        h.map_build(cm, sm)
    for rep in range(3):
        h.lib.vlo_online_reset(h._h); h.online_set_map_pose(pose0)
        if maintained and rep: h.map_reset(); h.map_insert(cm, sm, np.zeros(6, np.float32))
        if rep == 2: h.set_profiling(True)
        rows = []
        for k in range(SEQ):
            t1 = time.perf_counter()
            rc, o, m = h.process_scan(seq[k], 0.1 * k, want_map=True)
            rows.append(((time.perf_counter() - t1) * 1e3, int(o["iterations"]), int(m["iterations"]), int(m["status"])))
    st = h.stage_times()
    print("tick ms / odom it / map it / map status:", " ".join("%.2f/%d/%d/%d" % r for r in rows))
    lat = sorted(r[0] for r in rows[1:])
    print("p50 %.3f p95 %.3f max %.3f" % (lat[len(lat) // 2], lat[int(len(lat) * 0.95)], lat[-1]))
    print(" ".join("%s %.3f/%d" % (k, v[0] / (SEQ - 1), v[1]) for k, v in st.items() if v[1]))
