"""Experiment: the headline step on ONE handle (steps enqueued back to back on one stream) against TWO handles on two streams taking
alternate steps (their kernels may overlap).  Wall clock around one synchronisation, device-resident clouds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vil_sensor_fusion_b200 import api
B, STEPS, PF = 128, int(os.environ.get("STEPS", "20")), 3
cm, sm = bench.make_map()
raws_pool, seeds_pool = bench.make_workload(bench.POOL, 0)
pool_n = bench.POOL
n_buf = pool_n + B - 1
sizes = np.array([raws_pool[k % pool_n].shape[0] for k in range(n_buf)], np.int64)
offs_all = np.zeros(n_buf + 1, np.int64); offs_all[1:] = np.cumsum(sizes)
host = torch.empty((int(offs_all[-1]), PF), dtype=torch.float32).pin_memory()
hv = host.numpy()
for k in range(n_buf): hv[offs_all[k]:offs_all[k + 1]] = raws_pool[k % pool_n][:, :PF]
dev = host.cuda()
seeds_all = np.stack([seeds_pool[k % pool_n] for k in range(n_buf)])
scans_idx = np.arange(B, dtype=np.int32)
cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072, max_map_points=int(max(len(cm), len(sm))))
NH = int(os.environ.get('NH', '4'))
hs = [api.Handle(cfg) for _ in range(NH)]
for h in hs: h.map_build(cm, sm)
item = api.RESULT_DTYPE.itemsize
res_pin = torch.empty(STEPS * B * item, dtype=torch.uint8).pin_memory()
def window(step):
    w0 = (step * 61) % pool_n
    return w0, (offs_all[w0:w0 + B + 1] - offs_all[w0]).astype(np.int32), seeds_all[w0:w0 + B]
def run(n_handles):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(STEPS):
        h = hs[k % n_handles]
        w0, offs, seeds = window(k)
        h.upload_raw(dev.data_ptr() + int(offs_all[w0]) * 4 * PF, offs, PF, True)
        h.organise(); h.extract()
        h.register_map_enqueue(scans_idx, seeds, res_pin.data_ptr() + k * B * item)
    for h in hs: h.synchronize()
    dt = time.perf_counter() - t0
    r = np.frombuffer(res_pin.numpy(), api.RESULT_DTYPE)
    return dt / STEPS * 1e3, int((r["status"] == 0).sum()), float(r["iterations"].mean())
for n in (1, 2, 3, 4)[:NH]:
    run(n)
    print("handles %d: %.3f ms per step, ok %d, iters %.2f" % ((n,) + run(n)))
