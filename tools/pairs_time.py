"""Stage times of the whole-bag scan-to-scan step (127 HDL-64 pairs, device-synthesised sweeps); VLO_LIB_PATH picks a variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vil_sensor_fusion_b200 import api, synth, synth_gpu
B = int(os.environ.get("PAIRS_B", "128"))
scene = synth_gpu.make_scene(synth.scene_room(0)); sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
npts = sensor.rings * sensor.n_az
raw = torch.empty((B, npts, 4), dtype=torch.float32, device="cuda")
if os.environ.get("PAIRS_WORKLOAD") == "r01":
    # round 1's whole-bag leg: 8 numpy-synthesised sweeps of the straight trajectory in ping-pong order
    sc0 = synth.scene_room(0); traj = synth.Trajectory()
    pool = [synth.make_scan(sc0, "HDL-64E", t0=0.1 * k, traj=traj, rolling=False, noise_sigma=0.01, seed=k) for k in range(8)]
    order = list(range(8)) + list(range(6, 0, -1))
    for k in range(B):
        a = pool[order[k % len(order)]]
        raw[k].fill_(float("nan")); raw[k, :a.shape[0]] = torch.from_numpy(a).cuda()
else:
    synth_gpu.synth_scans(scene, sensor, 100, B, 1234, raw.data_ptr())
torch.cuda.synchronize()
offs = (np.arange(B + 1, dtype=np.int64) * npts).astype(np.int32)
cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=npts, odom_cell_size=float(os.environ.get("VLO_SURF_CELL", "1.0")),
                         odom_corner_cell_size=float(os.environ.get("VLO_CORNER_CELL", "5.0")))
with api.Handle(cfg) as h:
    def step():
        h.upload_raw(raw.data_ptr(), offs, 4, True); h.organise(); h.extract()
        return h.register_pairs(np.arange(B - 1), np.arange(1, B))
    for _ in range(2): r = step()
    h.set_profiling(True)
    for _ in range(5): r = step()
    st = h.stage_times()
    import zlib
    print(os.environ.get("VLO_LIB_PATH", "default"), "cells", cfg.odom_cell_size, cfg.odom_corner_cell_size, "pairs", B - 1, "iters %.1f ok %d crc %08x" % (r["iterations"].mean(), int((r["status"] == 0).sum()),
          zlib.crc32(r["transform"].tobytes())), " ".join("%s %.3f" % (k, v[0] / 5) for k, v in st.items() if v[1]))
