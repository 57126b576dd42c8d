#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LiDAR-odometry hot path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
  HDL-64-shaped scans (64 x 1800 = 115 200 points) -> ring organise (K0) -> feature extraction (K1)
  -> scan-to-map registration against a 1M-point voxel-hash map incl. eigenvalue degeneracy test,
  solution remapping and the D-optimality gate (K2/K5/K4).
One "step" = one batch of `--batch` scans per GPU through that path.  `value` = whole-job scans/s with
the raw clouds already resident in HBM; `e2e` = the same through the streaming C-ABI call (vlo_bag_register_map)
with HOST (pinned) buffers: host->device copies of every step's clouds and device->host reads of its result
records inside the timed region (the library overlaps the copy of one half-batch with the previous one's kernels).
N > 1: frames are sharded by contiguous range across ranks (one process per GPU, own map replica),
no collective in the per-scan path, one gather of the result records per step ("scaling": "weak").

`--impl reference` times the CPU restatement of the reference path (oracle/, kd-tree based, all host
threads, frames in parallel) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "LiDAR scans/sec (HDL-64 scan-to-map registration + degeneracy, 1M-point map)"
UNIT = "scans/s"
POOL = 8                 # distinct synthetic scans cycled through a batch
N_MAP = 1_000_000
SEED_PERTURB = np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32)


def make_workload(pool: int, rank: int = 0):
    from vil_sensor_fusion_b200 import synth
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    raws, seeds = [], []
    for k in range(pool):
        t = 0.1 * (k + pool * rank)
        raws.append(synth.make_scan(scene, "HDL-64E", t0=t, traj=traj, rolling=False, noise_sigma=0.01, seed=k + 100 * rank))
        gt = synth.loam_map_pose(traj.rotation(t), traj.position(t)).astype(np.float32)
        seeds.append(gt + SEED_PERTURB)
    cm, sm = synth.make_voxel_map(scene, N_MAP, seed=1)
    return raws, np.stack(seeds), cm, sm


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(stage: str, counts, n_map_pts, iters_done, q_stack=None):
    """ALGORITHMIC bytes one launch of `stage` moves for the batch (SURVEY.md 8d formulas).
    q_stack = total size of the down-sampled corner + surface stacks (the scan-to-map queries)."""
    nv = sum(c["n_valid"] for c in counts)
    nfeat = sum(c["n_sharp"] + c["n_less_sharp"] + c["n_flat"] + c["n_less_flat"] for c in counts)
    q_in = sum(c["n_less_sharp"] + c["n_less_flat"] for c in counts)
    q = q_in if q_stack is None else q_stack
    if stage == "k0_organise":
        return nv * (16 + 16)                                # raw xyz(i) in, ring-major float4 out
    if stage == "k1_extract":
        return nv * 21 + 4 * nfeat                           # 16 in + 4 curvature + 1 label, + index lists
    if stage == "k1b_compact":
        return 2 * 20 * sum(c["n_sharp"] + c["n_less_sharp"] + c["n_flat"] for c in counts)
    if stage == "k5_assoc":
        # association, one Gauss-Newton iteration: map read once + query in + 5 neighbour indices out
        return n_map_pts * 16 + q * (16 + 5 * 4)
    if stage == "k5_lin":
        # linearisation, one Gauss-Newton iteration (116 B per query = point + 5 indices + 5 gathered
        # neighbours) + level-1 sums out and read back by the slot's solving warp
        return q * (16 + 5 * 4 + 5 * 16) + 2 * (q // 32 + 1) * 28 * 4
    if stage == "k7_stack_ds":
        return q_in * 16 + q * 16                            # stacks in, voxel centroids out
    return 0


def bind_to_gpu_numa(local_rank: int):
    """Pin this rank's threads to the CPUs next to its GPU before any pinned host memory is allocated (first touch puts
    the staging buffers on the GPU's NUMA node): N ranks streaming scans over PCIe must not share one socket's memory."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        dev = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(dev + "/local_cpulist").read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return open(dev + "/numa_node").read().strip()
    except Exception:
        return None


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vil_sensor_fusion_b200 import api, bag

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if rank == 0 and world > 1:
            print("warning: WORLD_SIZE %d != --gpus %d" % (world, args.gpus), file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")               # NCCL's version banner goes to stdout: keep stdout to the one JSON line
        import datetime
        # a rank that dies must not leave the others waiting for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=240))

    B = args.batch
    raws_pool, seeds_pool, cm, sm = make_workload(POOL, rank)
    raws = [raws_pool[k % POOL] for k in range(B)]
    seeds = np.stack([seeds_pool[k % POOL] for k in range(B)])
    offs = np.zeros(B + 1, np.int32)
    offs[1:] = np.cumsum([r.shape[0] for r in raws])
    n_pts = int(offs[-1])
    host = torch.empty((n_pts, 4), dtype=torch.float32).pin_memory()
    host.numpy()[:] = np.concatenate(raws, axis=0)
    dev = host.to("cuda", non_blocking=False)
    scans_idx = np.arange(B, dtype=np.int32)
    h2d_bytes = n_pts * 16
    d2h_bytes = B * api.RESULT_DTYPE.itemsize

    # odom_cell_size: 0.7 m cells for the scan-to-scan surface grids of the whole-bag leg (tools/pairs_tune.py: batches
    # prefer smaller cells than the 1 m default that minimises the single-pair latency; results are identical)
    cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072, odom_cell_size=0.7,
                             odom_corner_cell_size=float(os.environ.get("VLO_CORNER_CELL", "5.0")),
                             max_map_points=int(max(len(cm), len(sm))), device=local_rank)
    if os.environ.get("VLO_MAP_CELL"):
        cfg.map_cell_size = float(os.environ["VLO_MAP_CELL"])          # tuning experiments only
    h = api.Handle(cfg)
    h.map_build(cm, sm)
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))

    def step(on_device: bool):
        if on_device:
            h.upload_raw(dev.data_ptr(), offs, 4, True)
        else:
            h.upload_raw(host.data_ptr(), offs, 4, False)
        h.organise()
        h.extract()
        return h.register_map(scans_idx, seeds)

    def local_sync():
        torch.cuda.synchronize()
        h.synchronize()

    def barrier():
        local_sync()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- correctness guard: the timed path must do the work (converged, correspondences found)
    res = step(True)
    counts = h.counts()
    nc_ds, ns_ds = h.stack_counts(B)
    q_stack = int(nc_ds.sum() + ns_ds.sum())
    ok = int(np.sum(res["status"] == 0))
    if ok < len(res):
        print("warning: %d of %d registrations reported a soft status" % (len(res) - ok, len(res)), file=sys.stderr)

    for _ in range(max(args.warmup - 1, 0)):
        step(True)
    if world > 1:
        # warm-up of the exchange step with the timed region's shapes (NCCL connects lazily on the first all_gather)
        bag.gather_results(np.concatenate([res] * args.steps))
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    h.set_profiling(True)
    launches0 = h.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    all_res = []
    for _ in range(args.steps):
        res = step(True)
        all_res.append(res)
    if world > 1:
        # the ONE exchange step of the whole job: every rank's result records, gathered once over NVLink (NCCL
        # all_gather of (K * B) x 480-byte records); nothing collective happens per scan or per batch
        gathered = bag.gather_results(np.concatenate(all_res))
        assert len(gathered) == world * args.steps * B
    ev1.record(stream)
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = h.launch_count() - launches0
    stages = h.stage_times()
    h.set_profiling(False)

    # ---- e2e: HOST (pinned) buffers through the streaming C-ABI call (vlo_bag_register_map): every step's clouds
    # cross PCIe and every step's result records come back inside the timed region; the library overlaps the copy
    # of one half-batch with the kernels of the previous one
    HBn = B // 2
    offs_h = [offs[:HBn + 1] - offs[0], offs[HBn:] - offs[HBn]]
    ptr_h = [host.data_ptr(), host.data_ptr() + int(offs[HBn]) * 16]
    seed_h = [seeds[:HBn], seeds[HBn:]]
    one_step = [(ptr_h[0], offs_h[0], seed_h[0]), (ptr_h[1], offs_h[1], seed_h[1])]

    def e2e_pass(n_steps):
        r = h.bag_register_map(one_step * n_steps, stride=4)
        if world > 1:
            bag.gather_results(r)                       # the job's single exchange step, inside the timed region
        return r

    e2e_pass(2)
    # three passes of K steps, the MEDIAN pass is reported (host-side PCIe / scheduling hiccups on a shared box make
    # single passes noisy; every pass is listed in e2e.passes_ms_per_step)
    e2e_passes = []
    for _ in range(3):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        res_e2e = e2e_pass(args.steps)
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        e2e_passes.append((max(wall, e0.elapsed_time(e1)), wall, e0.elapsed_time(e1)))
    e2e_ms, e2e_wall_ms, e2e_dev_ms = sorted(e2e_passes)[1]
    e2e_all = [p_[0] for p_ in e2e_passes]
    if not np.array_equal(res_e2e["transform"][:B].view(np.uint32), res["transform"][:B].view(np.uint32)):
        print("warning: streaming e2e results differ from the resident-batch results", file=sys.stderr)
    clocks.stop()
    # what bounds e2e: the same pinned buffer copied host -> device on its own (plain cudaMemcpyAsync, nothing else
    # running) -- the PCIe rate of this box.  e2e.frac_of_pcie = (h2d bytes per step / e2e time per step) / that rate.
    pcie_gbs = None
    try:
        best = None
        for _ in range(4):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            c0.record()
            dev.copy_(host, non_blocking=True)
            c1.record()
            torch.cuda.synchronize()
            t_ms = c0.elapsed_time(c1)
            best = t_ms if best is None else min(best, t_ms)
        pcie_gbs = h2d_bytes / (best * 1e-3) / 1e9
    except Exception:
        pass

    # ---- whole-bag scan-to-scan leg (BASELINE config 5 shape): independent registrations of consecutive
    # sweep pairs from a zero seed + degeneracy; sweeps in ping-pong order (0 1 .. 7 6 .. 0 1 ..) so that neighbours in
    # the batch are neighbours in time.  Reported beside the headline, not instead of it.
    pairs_leg = None
    if rank == 0 and not args.no_latency:
        order = list(range(POOL)) + list(range(POOL - 2, 0, -1))
        seq = [order[k % len(order)] for k in range(B)]
        host2 = torch.empty((n_pts, 4), dtype=torch.float32).pin_memory()
        host2.numpy()[:] = np.concatenate([raws_pool[i] for i in seq], axis=0)
        offs2 = np.zeros(B + 1, np.int32)
        offs2[1:] = np.cumsum([raws_pool[i].shape[0] for i in seq])
        dev2 = host2.to("cuda")
        last_i, cur_i = np.arange(B - 1, dtype=np.int32), np.arange(1, B, dtype=np.int32)

        def pstep():
            h.upload_raw(dev2.data_ptr(), offs2, 4, True)
            h.organise()
            h.extract()
            return h.register_pairs(last_i, cur_i)

        rp = pstep()
        pcounts = h.counts()
        for _ in range(2):
            pstep()
        local_sync()                    # rank 0 only: no collective here
        h.set_profiling(True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(args.steps):
            rp = pstep()
        p1.record(stream)
        local_sync()                    # rank 0 only: no collective here
        pst = h.stage_times()
        h.set_profiling(False)
        p_ms = p0.elapsed_time(p1) / args.steps
        qe = np.array([c["n_sharp"] for c in pcounts[1:]], np.float64)
        qp = np.array([c["n_flat"] for c in pcounts[1:]], np.float64)
        gn_bytes = float(np.sum(rp["iterations"] * (56 * qe + 76 * qp + 108)))        # B_lin per GN iteration
        gn_ms = pst["k3_gn"][0] / args.steps
        m_t = np.array([c["n_less_sharp"] + c["n_less_flat"] for c in pcounts[:-1]], np.float64)
        rounds = np.ceil(rp["iterations"] / 5.0)
        as_bytes = float(np.sum(rounds * (16 * m_t + qe * 24 + qp * 28)))
        as_ms = pst["k3_assoc"][0] / args.steps
        pairs_leg = {"value": round((B - 1) / (p_ms * 1e-3), 1), "unit": "scan pairs/s", "ms_per_step": round(p_ms, 4),
                     "pairs_per_step": B - 1, "mean_gn_iterations": round(float(np.mean(rp["iterations"])), 2),
                     "ok": int(np.sum(rp["status"] == 0)), "degenerate": int(np.sum(rp["is_degenerate"] != 0)),
                     "k3_gn": {"ms_per_step": round(gn_ms, 4), "algorithmic_bytes": int(gn_bytes),
                               "achieved_gbs": round(gn_bytes / (gn_ms * 1e-3) / 1e9, 2) if gn_ms > 0 else None},
                     "k3_assoc": {"ms_per_step": round(as_ms, 4), "algorithmic_bytes": int(as_bytes),
                                  "achieved_gbs": round(as_bytes / (as_ms * 1e-3) / 1e9, 2) if as_ms > 0 else None},
                     "k2_grid_build_ms_per_step": round(pst["k2_grid_build"][0] / args.steps, 4),
                     "what": "organise + extract + scan-to-scan registration of consecutive HDL-64 sweeps (zero seed, <= 25 GN "
                             "iterations, re-association every 5) + eigen-degeneracy + D-opt gate; clouds resident in HBM"}
        del dev2, host2

    # ---- online latency (single scan per call, the online path): p50 / p95 of vlo_process_scan
    p50 = p95 = None
    extra_lat = {}
    if rank == 0 and not args.no_latency:
        from vil_sensor_fusion_b200 import synth
        cfg1 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, io_ratio=1,
                                  max_map_points=int(max(len(cm), len(sm))), device=local_rank)
        # the caller's scan buffers are pinned (as a driver's DMA buffers would be): numpy views of pinned torch tensors
        pinned_pool = []
        for r in raws_pool:
            tp = torch.empty(r.shape, dtype=torch.float32).pin_memory()
            tp.numpy()[:] = r
            pinned_pool.append(tp)
        raws_lat = [tp.numpy() for tp in pinned_pool]
        with api.Handle(cfg1) as h1:
            h1.map_build(cm, sm)
            traj = synth.Trajectory()
            h1.online_set_map_pose(synth.loam_map_pose(traj.rotation(0.0), traj.position(0.0)).astype(np.float32))
            lat = []
            for rep in range(4):
                for k in range(POOL):
                    if k == 0:
                        h1.lib.vlo_online_reset(h1._h)
                        h1.online_set_map_pose(synth.loam_map_pose(traj.rotation(0.0), traj.position(0.0)).astype(np.float32))
                    t1 = time.perf_counter()
                    h1.process_scan(raws_lat[k], 0.1 * k, want_map=True)
                    if rep > 0 and k > 0:
                        lat.append((time.perf_counter() - t1) * 1e3)
            lat.sort()
            p50 = lat[len(lat) // 2]
            p95 = lat[int(len(lat) * 0.95)]
            # scan-to-map registration + degeneracy alone (north-star target: p50 < 1 ms): the scan is resident and
            # extracted, the call uploads the seed, registers against the 1M-point map and reads the record back
            h1.upload([raws_pool[k % POOL] for k in range(2)])
            h1.organise()
            h1.extract()
            h1.register_map([0], [seeds_pool[0]])
            lat_map = []
            for rep in range(40):
                k = rep % 2
                t1 = time.perf_counter()
                h1.register_map([k], [seeds_pool[k]])
                lat_map.append((time.perf_counter() - t1) * 1e3)
            lat_map.sort()
            extra_lat["scan_to_map_p50_ms"] = round(lat_map[len(lat_map) // 2], 4)
            extra_lat["scan_to_map_p95_ms"] = round(lat_map[int(len(lat_map) * 0.95)], 4)
        # the same tick with the MAINTAINED map (BasicLaserMapping::process: sub-map selection, optimisation, insertion)
        cfg2 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, io_ratio=1,
                                  max_map_points=1 << 20, device=local_rank)
        with api.Handle(cfg2) as h2:
            lat2 = []
            for rep in range(3):
                h2.lib.vlo_online_reset(h2._h)
                h2.map_reset()
                h2.map_insert(cm, sm, np.zeros(6, np.float32))          # prior map: the same 1M points, voxel-filtered
                h2.online_set_map_pose(synth.loam_map_pose(traj.rotation(0.0), traj.position(0.0)).astype(np.float32))
                for k in range(POOL):
                    t1 = time.perf_counter()
                    h2.process_scan(raws_lat[k], 0.1 * k, want_map=True)
                    if rep > 0 and k > 0:
                        lat2.append((time.perf_counter() - t1) * 1e3)
            lat2.sort()
            extra_lat["maintained_map_tick_p50_ms"] = round(lat2[len(lat2) // 2], 4)
            extra_lat["maintained_map_points"] = list(h2.map_size())

    # max over ranks
    if world > 1:
        ta = torch.tensor(e2e_all, dtype=torch.float64, device="cuda")
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)      # a pass is as slow as its slowest rank
        e2e_all = [float(v) for v in ta]
        e2e_ms = sorted(e2e_all)[1]
        t = torch.tensor([dev_ms, e2e_ms, float(launches)], dtype=torch.float64, device="cuda")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dev_ms, e2e_ms = float(tmax[0]), float(tmax[1])
        launches = int(tsum[2])
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        if pairs_leg:
            for kname in ("k3_gn", "k3_assoc"):
                g = pairs_leg[kname]["achieved_gbs"]
                pairs_leg[kname]["frac"] = round(g / peak, 4) if g else None
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        n_map_pts = len(cm) + len(sm)
        mean_iters = float(np.mean(res["iterations"]))
        table = {}
        for name, (ms, n) in stages.items():
            if n == 0:
                continue
            # k5_* launch max_iter times but only the first `iterations` do work: average over working launches
            working = n
            if name in ("k5_assoc", "k5_lin"):
                # max_iter launches per call, only the first `iterations` do work: average over the working launches
                working = max(1, int(round(args.steps * min(mean_iters, cfg.map_max_iterations))))
            avg_ms = ms / working
            ab = algorithmic_bytes(name, counts, n_map_pts, mean_iters, q_stack)
            table[name] = {"ms_total": round(ms, 4), "launches": n, "working_launches": working, "avg_ms": round(avg_ms, 5),
                           "algorithmic_bytes": ab, "achieved_gbs": round(ab / (avg_ms * 1e-3) / 1e9, 2) if avg_ms > 0 else None,
                           "frac": round(ab / (avg_ms * 1e-3) / 1e9 / peak, 4) if avg_ms > 0 else None}
        dom = max(table, key=lambda k: table[k]["ms_total"])
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
            traffic = tr.get(dom, {}).get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": table[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": table[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                    "share_of_step": round(table[dom]["ms_total"] / dev_ms, 4),
                    "note": "algorithmic bytes over the HBM peak, as the contract asks; ncu (profiles/SUMMARY.md) shows the scan-to-map "
                            "kernels issue / latency bound on an L2-resident map (DRAM throughput ~2 %), so the fraction explains, it "
                            "does not grade the kernel"}
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "hdl64_scan_to_map_1M: HDL-64-shaped scans (64x1800) -> organise + feature extraction + "
                                   "scan-to-map registration (<=10 GN iterations, 5-NN on a 1M-point voxel-hash map) + eigen-degeneracy "
                                   "+ D-opt gate", "scans_per_step_per_gpu": B, "points_per_scan": int(n_pts // B),
                       "map_points": int(n_map_pts), "parallelism": "frame-range dp%d" % world,
                       "l2": "inputs %.0f MB per step > 126 MB L2" % (h2d_bytes / 1e6), "mean_gn_iterations": round(mean_iters, 2),
                       "pool": "%d distinct scans cycled" % POOL, "exchange": "one all_gather of the result records per job" if world > 1 else "none",
                       "numa_node": numa},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": round(e2e_ms / args.steps, 4), "wall_ms": round(e2e_wall_ms, 3), "device_ms": round(e2e_dev_ms, 3),
                    "passes_ms_per_step": [round(v / args.steps, 4) for v in e2e_all], "reported": "median of 3 passes of K steps",
                    "pcie_h2d_gbs_measured": None if pcie_gbs is None else round(pcie_gbs, 2),
                    "h2d_gbs_achieved": round(h2d_bytes / (e2e_ms / args.steps * 1e-3) / 1e9, 2),
                    "frac_of_pcie": None if not pcie_gbs else round(h2d_bytes / (e2e_ms / args.steps * 1e-3) / 1e9 / pcie_gbs, 4),
                    "bound": "PCIe host->device copy of the 16 B/point PointCloud2 payload (kernels overlap it)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "stages": table,
            "latency": {"p50_ms_per_scan": None if p50 is None else round(p50, 4), "p95_ms_per_scan": None if p95 is None else round(p95, 4),
                        "what": "vlo_process_scan: one online tick (H2D from a pinned buffer + organise + extract + scan-to-scan + scan-to-map on every sweep (ioRatio 1) + results D2H)",
                        **extra_lat},
            "whole_bag_pairs": pairs_leg,
            "ok_registrations": ok, "mean_corr": [float(np.mean(res["n_corr_edge"])), float(np.mean(res["n_corr_plane"]))],
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(raws_pool, seeds_pool, cm, sm, threads=1, n_scans=args.cpu_sample)
        print(json.dumps(out))
    h.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(raws_pool, seeds_pool, cm, sm, threads: int, n_scans: int):
    """The oracle (CPU restatement of the reference path, kd-tree based) on a bounded sample."""
    from oracle import oracle as orc
    cfg = orc.default_config("HDL-64E", deskew=0)
    m = orc.CpuMap(cfg, cm, sm)
    raws = [raws_pool[k % len(raws_pool)] for k in range(n_scans)]
    seeds = np.stack([seeds_pool[k % len(raws_pool)] for k in range(n_scans)])
    m.batch_scan_to_map(raws[:max(1, threads)], seeds[:max(1, threads)], threads)       # warm-up (page in, caches)
    t0 = time.perf_counter()
    res = m.batch_scan_to_map(raws, seeds, threads)
    dt = time.perf_counter() - t0
    m.close()
    return {"value": round(n_scans / dt, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d HDL-64 scans of the same workload (organise + extract + scan-to-map on the 1M map, kd-trees prebuilt), %.1f s"
                      % (n_scans, dt), "mean_gn_iterations": float(np.mean([r["iterations"] for r in res]))}


def run_reference(args):
    """--impl reference: the CPU restatement on all host threads; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    raws_pool, seeds_pool, cm, sm = make_workload(POOL, 0)
    from oracle import oracle as orc
    cfg = orc.default_config("HDL-64E", deskew=0)
    m = orc.CpuMap(cfg, cm, sm)
    n_per_step = max(threads, 4) * 2
    raws = [raws_pool[k % POOL] for k in range(n_per_step)]
    seeds = np.stack([seeds_pool[k % POOL] for k in range(n_per_step)])
    for _ in range(min(args.warmup, 2)):
        m.batch_scan_to_map(raws, seeds, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = m.batch_scan_to_map(raws, seeds, threads)
    dt = time.perf_counter() - t0
    m.close()
    value = n_per_step * args.steps / dt
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "hdl64_scan_to_map_1M (CPU restatement of the reference path: oracle/, kd-tree based)",
                   "scans_per_step": n_per_step, "map_points": int(len(cm) + len(sm))},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d scans per step x %d steps, frames spread over %d threads" % (n_per_step, args.steps, threads)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_gn_iterations": float(np.mean([r["iterations"] for r in res])),
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="scans per step per GPU")
    ap.add_argument("--impl", default="vlo", choices=["vlo", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true", help="skip the online-tick latency leg (profiling runs)")
    ap.add_argument("--cpu-sample", type=int, default=96, help="scans in the bounded CPU-baseline sample (~15 s on one core)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "vlo" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
