"""Whole-bag records of frames [F0, F0 + N) to gpurun_out/pairs_<tag>.npy (VLO_LIB_PATH picks the build)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vil_sensor_fusion_b200 import api, synth, synth_gpu
tag, F0, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
scene = synth_gpu.make_scene(synth.scene_room(0)); sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
npts = sensor.rings * sensor.n_az; PB = 256
raw = torch.empty((PB, npts, 4), dtype=torch.float32, device="cuda")
offs = (np.arange(PB + 1, dtype=np.int64) * npts).astype(np.int32)
cfg = api.default_config("HDL-64E", deskew=0, max_scans=PB, max_points=npts)
out = []
with api.Handle(cfg) as h:
    k = F0
    while k < F0 + N:
        n = min(PB, F0 + N + 1 - k)
        synth_gpu.synth_scans(scene, sensor, k, n, 1234, raw.data_ptr(), h.stream_ptr())
        h.upload_raw(raw.data_ptr(), offs[:n + 1], 4, True); h.organise(); h.extract()
        out.append(h.register_pairs(np.arange(n - 1), np.arange(1, n)))
        k += n - 1
r = np.concatenate(out)
np.save("gpurun_out/pairs_%s.npy" % tag, r)
print(tag, len(r))
