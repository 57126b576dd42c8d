"""Dump the sweeps of pair F (frames F, F+1), the record and the 5 rounds of correspondences of this build to gpurun_out/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vil_sensor_fusion_b200 import api, synth, synth_gpu
tag, F = sys.argv[1], int(sys.argv[2])
scene = synth_gpu.make_scene(synth.scene_room(0)); sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
npts = sensor.rings * sensor.n_az; NB = 8
raw = torch.empty((NB, npts, 4), dtype=torch.float32, device="cuda")
synth_gpu.synth_scans(scene, sensor, F, NB, 1234, raw.data_ptr()); torch.cuda.synchronize()
np.save("gpurun_out/pair_raw_%d.npy" % F, raw[:2].cpu().numpy())
offs = (np.arange(NB + 1, dtype=np.int64) * npts).astype(np.int32)
cfg = api.default_config("HDL-64E", deskew=0, max_scans=NB, max_points=npts)
with api.Handle(cfg) as h:
    h.lib.vlo_set_trace(h._h, 1)
    h.upload_raw(raw.data_ptr(), offs, 4, True); h.organise(); h.extract()
    r = h.register_pairs(np.arange(NB - 1), np.arange(1, NB))          # > 4 pairs: the batch (box) kernel
    c = h.counts()
    tr = [h.pair_correspondences(0, rnd, c[1]["n_sharp"], c[1]["n_flat"]) for rnd in range(5)]
np.save("gpurun_out/pair_res_%s_%d.npy" % (tag, F), r[:1])
np.savez("gpurun_out/pair_trace_%s_%d.npz" % (tag, F), **{"c%d" % i: t[0] for i, t in enumerate(tr)}, **{"s%d" % i: t[1] for i, t in enumerate(tr)})
print(tag, r["transform"][0])
