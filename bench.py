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
POOL = 200               # 200 query scans at perturbed poses, seeds 1 .. 200
N_MAP = 1_000_000
SEED_PERTURB = np.array([0.004, -0.006, 0.003, 0.06, -0.04, 0.08], np.float32)


def make_map():
    from vil_sensor_fusion_b200 import synth
    return synth.make_voxel_map(synth.scene_room(0), N_MAP, seed=1)


def make_workload_r01(pool: int, rank: int = 0):
    """Round 1's pool (8 numpy-synthesised scans, one fixed seed offset): kept behind --r01-workload for comparisons."""
    from vil_sensor_fusion_b200 import synth
    scene = synth.scene_room(0)
    traj = synth.Trajectory()
    raws, seeds = [], []
    for k in range(pool):
        t = 0.1 * (k + pool * rank)
        raws.append(synth.make_scan(scene, "HDL-64E", t0=t, traj=traj, rolling=False, noise_sigma=0.01, seed=k + 100 * rank))
        gt = synth.loam_map_pose(traj.rotation(t), traj.position(t)).astype(np.float32)
        seeds.append(gt + SEED_PERTURB)
    return raws, np.stack(seeds)


def make_workload(pool: int, rank: int = 0):
    """`pool` HDL-64 query scans at distinct poses of a loop through the room, each registered from its
    true pose perturbed by N(0, 0.1 m) / N(0, 0.5 deg) (numpy default_rng(seed), seeds 1 .. pool).  The sweeps are
    ray-cast on the device (libvlo_synth.so) and copied back, so that the CPU arm sees the very same clouds."""
    import torch
    from vil_sensor_fusion_b200 import synth, synth_gpu
    scene = synth_gpu.make_scene(synth.scene_room(0))
    sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
    npts = sensor.rings * sensor.n_az
    stride_frames = 7                                   # 0.7 s between pool poses: the pool spans several laps' worth of poses
    buf = torch.empty((pool, npts, 4), dtype=torch.float32, device="cuda")
    for k in range(pool):
        synth_gpu.synth_scans(scene, sensor, (k + pool * rank) * stride_frames, 1, 1234, buf[k].data_ptr())
    torch.cuda.synchronize()
    raws = list(buf.cpu().numpy())
    del buf
    seeds = []
    for k in range(pool):
        R, p = synth_gpu.poses(sensor, (k + pool * rank) * stride_frames, 1)
        gt = synth.loam_map_pose(R[0], p[0]).astype(np.float32)
        g = np.random.default_rng(1 + k + pool * rank)
        seeds.append(gt + np.concatenate([g.normal(0.0, np.deg2rad(0.5), 3), g.normal(0.0, 0.1, 3)]).astype(np.float32))
    return raws, np.stack(seeds)


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


RAW_BYTES = 12           # bytes per raw point of the headline workload (set from --point-floats)


def algorithmic_bytes(stage: str, counts, n_map_pts, iters_done, q_stack=None):
    """ALGORITHMIC bytes one launch of `stage` moves for the batch (SURVEY.md 8d formulas).
    q_stack = total size of the down-sampled corner + surface stacks (the scan-to-map queries)."""
    nv = sum(c["n_valid"] for c in counts)
    nfeat = sum(c["n_sharp"] + c["n_less_sharp"] + c["n_flat"] + c["n_less_flat"] for c in counts)
    q_in = sum(c["n_less_sharp"] + c["n_less_flat"] for c in counts)
    q = q_in if q_stack is None else q_stack
    if stage == "k0_organise":
        return nv * (RAW_BYTES + 16)                         # raw point in (12 B xyz / 16 B xyzi), ring-major float4 out
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


def _pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(len(v) * q))]


def crc_of(records) -> int:
    import zlib
    return zlib.crc32(np.ascontiguousarray(records["transform"]).tobytes() + np.ascontiguousarray(records["hessian"]).tobytes())


def whole_bag_pairs_leg(args, rank, world, local_rank, barrier, peak):
    """SURVEY 8d C5 / BASELINE configs[4]: offline whole-bag reprocessing of `--bag-pairs` HDL-64 scan pairs (k, k+1):
    independent scan-to-scan registrations from a zero seed + eigen-degeneracy + D-opt gate, sharded by PAIR range
    across the ranks (bag.pair_range), one gather of the result records per job, nothing collective per scan.  The
    sweeps are synthesised in device memory from (seed, frame id) inside the timed region (no input transfer).  The job
    is repeated until the timed region is at least --bag-seconds long; value = pairs / time, max over ranks."""
    import torch
    import torch.distributed as dist
    from vil_sensor_fusion_b200 import api, bag, synth, synth_gpu
    n_frames = args.bag_pairs + 1
    lo, hi = bag.pair_range(n_frames, rank, world)
    counts = [bag.pair_range(n_frames, r, world)[1] - bag.pair_range(n_frames, r, world)[0] for r in range(world)]
    scene = synth_gpu.make_scene(synth.scene_room(0))
    sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
    npts = sensor.rings * sensor.n_az
    PB = args.bag_batch
    cfg = api.default_config("HDL-64E", deskew=0, max_scans=PB, max_points=npts, max_map_points=0, device=local_rank)
    # two handles (two resident batches, two streams) take alternate batches of the job: the kernels of consecutive batches overlap;
    # every batch is only enqueued (vlo_register_pairs_enqueue), the host synchronises once per job
    hs = [api.Handle(cfg) for _ in range(int(os.environ.get("VLO_BAG_HANDLES", "2")))]
    h = hs[0]
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))
    raws = [torch.empty((PB, npts, 4), dtype=torch.float32, device="cuda") for _ in hs]
    offs = (np.arange(PB + 1, dtype=np.int64) * npts).astype(np.int32)
    synth_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    item = api.RESULT_DTYPE.itemsize
    res_pin = torch.empty(max(hi - lo, 1) * item, dtype=torch.uint8).pin_memory()

    def job(time_synth=False, n_handles=len(hs)):
        k, b, pos, synth_ms = lo, 0, 0, 0.0
        while k < hi:
            hh, raw = hs[b % n_handles], raws[b % n_handles]
            n = min(PB, hi + 1 - k)                          # frames k .. k + n - 1 -> n - 1 pairs; batches overlap by one frame
            if time_synth:
                synth_ev[0].record(stream)
            synth_gpu.synth_scans(scene, sensor, k, n, 1234, raw.data_ptr(), hh.stream_ptr())
            if time_synth:
                synth_ev[1].record(stream)
            hh.upload_raw(raw.data_ptr(), offs[:n + 1], 4, True)
            hh.organise()
            hh.extract()
            hh.register_pairs_enqueue(np.arange(n - 1), np.arange(1, n), res_pin.data_ptr() + pos * item)
            if time_synth:
                hh.synchronize()
                synth_ms += synth_ev[0].elapsed_time(synth_ev[1])
            k += n - 1
            pos += n - 1
            b += 1
        for hh in hs[:n_handles]:
            hh.synchronize()
        local = h.results_finish(np.frombuffer(res_pin.numpy(), api.RESULT_DTYPE)[:pos].copy())
        allr = bag.gather_results(local, counts=counts) if world > 1 else local
        return allr, synth_ms

    # warm-up jobs: both handles once; then one job on ONE handle with the stage timers on (every kernel alone: `stages_rank0`),
    # which also sizes the number of passes of the timed region
    job()
    h.set_profiling(True)
    barrier()
    t0 = time.perf_counter()
    allr, synth_ms = job(time_synth=True, n_handles=1)
    barrier()
    est = time.perf_counter() - t0
    st = h.stage_times()
    h.set_profiling(False)
    if world > 1:
        te = torch.tensor([est], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        est = float(te[0])
    passes = max(1, int(np.ceil(args.bag_seconds / max(est, 1e-3))))
    launches0 = sum(hh.launch_count() for hh in hs)
    barrier()
    t0 = time.perf_counter()
    for _ in range(passes):
        allr, _ = job()
    barrier()
    t_ms = (time.perf_counter() - t0) * 1e3                 # wall clock between two device-synchronised barriers (two streams)
    launches = sum(hh.launch_count() for hh in hs) - launches0
    if world > 1:
        tt = torch.tensor([t_ms, float(launches)], dtype=torch.float64, device="cuda")
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        t_ms, launches = float(tmax[0]), int(tt[1])
    assert len(allr) == args.bag_pairs
    leg = None
    if rank == 0:
        n_local = hi - lo
        pc = h.counts()
        qe = float(np.mean([c["n_sharp"] for c in pc])); qp = float(np.mean([c["n_flat"] for c in pc]))
        m_t = float(np.mean([c["n_less_sharp"] + c["n_less_flat"] for c in pc]))
        nv = float(np.mean([c["n_valid"] for c in pc]))
        it = float(np.mean(allr["iterations"]))
        per = {}                                         # per-stage: ms per pass on rank 0 + algorithmic bytes + fraction of HBM peak
        alg = {"k0_organise": n_local * nv * 32, "k1_extract": n_local * (nv * 21 + 4 * (qe + qp + m_t)),
               "k2_grid_build": n_local * 16 * m_t, "k3_assoc": n_local * np.ceil(it / 5.0) * (16 * m_t + qe * 24 + qp * 28),
               "k3_gn": n_local * it * (56 * qe + 76 * qp + 108)}
        for name, (ms, n) in st.items():
            if n == 0:
                continue
            ms_pass = ms                                 # the stage timers ran over one serial job (one handle, every kernel alone)
            ab = alg.get(name)
            per[name] = {"ms_per_pass": round(ms_pass, 3), "algorithmic_bytes": None if ab is None else int(ab),
                         "achieved_gbs": None if not ab else round(ab / (ms_pass * 1e-3) / 1e9, 2),
                         "frac": None if not ab else round(ab / (ms_pass * 1e-3) / 1e9 / peak, 4)}
        leg = {"metric": "scan pairs/s (whole-bag scan-to-scan reprocessing, sharded by pair range)", "value": round(passes * args.bag_pairs / (t_ms * 1e-3), 1),
               "unit": "scan pairs/s", "scaling": "strong", "n_gpus": world, "pairs_per_job": args.bag_pairs, "passes": passes,
               "timed_region_s": round(t_ms * 1e-3, 3), "pairs_per_rank": counts, "frames_per_batch": PB,
               "mean_gn_iterations": round(it, 2), "ok": int(np.sum(allr["status"] == 0)), "degenerate": int(np.sum(allr["is_degenerate"] != 0)),
               "pass_dopt": int(np.sum(allr["pass_dopt"] != 0)), "records_crc32": crc_of(allr),
               "inputs": "synthesised on the device from (seed, frame id) inside the timed region: %.1f %% of a pass" % (100.0 * synth_ms / max(est * 1e3, 1e-9)),
               "exchange": "one all_gather_into_tensor of the %d-byte records per job" % api.RESULT_DTYPE.itemsize if world > 1 else "none",
               "gpu_launches": int(launches), "stages_rank0": per,
               "how": "two handles (two resident batches on two streams) take alternate 256-frame batches, every batch only enqueued, one host "
                      "synchronisation per job; stages_rank0: one job on one handle with the stage timers on",
               "what": "organise + extract + box index + scan-to-scan registration of consecutive HDL-64 sweeps (zero seed, <= 25 GN iterations, "
                       "re-association every 5) + eigen-degeneracy + D-opt gate"}
    for hh in hs:
        hh.close()
    del raws
    return leg


def imu_batch_leg(args, local_rank, peak, with_cpu):
    """SURVEY 8d C4 / BASELINE configs[3]: 10 000 keyframe intervals of 0.1 s over a jittered 200 Hz stream, keyframe times
    between samples (forces the interpolated last step), one warp per factor (K6)."""
    from vil_sensor_fusion_b200 import api
    n_f = args.imu_factors
    g = np.random.default_rng(2)
    n_s = 20 * n_f + 40
    t = np.arange(n_s) / 200.0 + g.uniform(-1e-4, 1e-4, n_s)
    # smooth random signals: sums of a few sinusoids (spline-like), specific force incl. gravity reaction
    ph = g.uniform(0, 2 * np.pi, (6, 4)); fr = g.uniform(0.05, 1.5, (6, 4)); am = g.uniform(0.1, 1.0, (6, 4))
    sig = np.stack([np.sum(am[c][None] * np.sin(2 * np.pi * fr[c][None] * t[:, None] + ph[c][None]), axis=1) for c in range(6)], axis=1)
    acc = np.ascontiguousarray(sig[:, :3] + np.array([0.0, 0.0, 9.81])); gyro = np.ascontiguousarray(0.2 * sig[:, 3:])
    t0 = 0.05 + 0.1 * np.arange(n_f) + 0.0023
    t1 = t0 + 0.1
    bias = np.full(6, 1e-2)
    cfg = api.default_config("VLP-16", max_scans=2, max_points=1024, device=local_rank)
    with api.Handle(cfg) as h:
        h.imu_preintegrate_batch(t, acc, gyro, t0, t1, bias)                 # warm-up (allocations)
        h.set_profiling(True)
        reps, walls = 5, []
        for _ in range(reps):
            w0 = time.perf_counter()
            out = h.imu_preintegrate_batch(t, acc, gyro, t0, t1, bias)
            walls.append(time.perf_counter() - w0)
        st = h.stage_times()
        h.set_profiling(False)
    k_ms = st["k6_imu"][0] / max(st["k6_imu"][1], 1)
    ab = 56 * n_s + 2408 * n_f
    leg = {"metric": "IMU factors/s (batched Forster preintegration between keyframes)", "n_factors": n_f, "n_samples": n_s,
           "mean_samples_per_factor": round(float(np.mean(out["n_integrated"])), 2),
           "value": round(n_f / (k_ms * 1e-3), 1), "unit": "factors/s", "kernel_ms": round(k_ms, 4),
           "e2e": {"value": round(n_f / float(np.median(walls)), 1), "unit": "factors/s", "h2d_bytes": int(56 * n_s + 16 * n_f + 48),
                   "d2h_bytes": int(out.nbytes), "what": "vlo_imu_preintegrate_batch with host arrays: H2D of the stream, kernel, D2H of the factors"},
           "roofline": {"kernel": "k6_imu_preintegrate", "bound": "hbm", "algorithmic_bytes": int(ab), "achieved": round(ab / (k_ms * 1e-3) / 1e9, 2),
                        "peak": peak, "unit": "GB/s", "frac": round(ab / (k_ms * 1e-3) / 1e9 / peak, 4),
                        "note": "float64 15x15 covariance propagation per sample: FP64-pipe / latency bound, not HBM"}}
    if with_cpu:
        from oracle import oracle as orc
        prm = orc.imu_params()
        n_c = min(n_f, 4000)
        c0 = time.perf_counter()
        ref = orc.imu_batch(prm, t, acc, gyro, t0[:n_c], t1[:n_c], bias, n_threads=1)
        dt = time.perf_counter() - c0
        leg["cpu_baseline"] = {"value": round(n_c / dt, 1), "unit": "factors/s", "cores": 1, "kind": "port",
                               "sample": "%d of the %d factors, oracle/imu_preint.c, %.2f s" % (n_c, n_f, dt)}
        leg["max_abs_diff_vs_oracle"] = {k: float(np.max(np.abs(out[k][:n_c] - ref[k]))) for k in ("dR", "dP", "dV")}
    return leg


def vlp16_online_leg(args, local_rank, with_cpu):
    """SURVEY 8d C1 / BASELINE configs[0]: VLP-16 sequence (16 x 1800 @ 10 Hz) + 200 Hz IMU through the online path:
    vlo_process_scan per sweep (scan-to-scan odometry + degeneracy + D-opt gate, LaserMapping with the maintained map on
    every ioRatio-th sweep) and the IMU factors of the keyframe intervals."""
    import torch
    from vil_sensor_fusion_b200 import api, synth, synth_gpu
    n_scans = args.vlp16_scans
    scene = synth_gpu.make_scene(synth.scene_room(0))
    sensor = synth_gpu.make_sensor("VLP-16", noise_sigma=0.02, rolling=True)
    npts = sensor.rings * sensor.n_az
    buf = torch.empty((n_scans, npts, 4), dtype=torch.float32, device="cuda")
    synth_gpu.synth_scans(scene, sensor, 0, n_scans, 7, buf.data_ptr())
    torch.cuda.synchronize()
    host = torch.empty((n_scans, npts, 4), dtype=torch.float32).pin_memory()
    host.copy_(buf)
    del buf
    scans = host.numpy()
    cfg = api.default_config("VLP-16", deskew=1, max_scans=2, max_points=32768, max_map_points=1 << 19, device=local_rank)
    lat, n_deg, n_drop, soft = [], 0, 0, 0
    with api.Handle(cfg) as h:
        h.map_reset()
        for k in range(min(10, n_scans)):                       # warm-up ticks (allocation, first launches)
            h.process_scan(scans[k], 0.1 * k, want_map=True)
        h.lib.vlo_online_reset(h._h)
        h.map_reset()
        w0 = time.perf_counter()
        for k in range(n_scans):
            t1 = time.perf_counter()
            rc, o, m = h.process_scan(scans[k], 0.1 * k, want_map=True)
            lat.append((time.perf_counter() - t1) * 1e3)
            if k > 0:
                n_deg += int(o["is_degenerate"] != 0); n_drop += int(o["pass_dopt"] == 0); soft += int(o["status"] != 0)
        wall = time.perf_counter() - w0
        # the IMU factors of the 0.1 s keyframe intervals: one batched call (200 Hz stream of the same trajectory, analytic)
        g = np.random.default_rng(0)
        ts = np.arange(int(20 * n_scans) + 20) / 200.0
        acc = np.tile(np.array([0.0, 0.0, 9.81]), (len(ts), 1)) + g.normal(0, 1e-3, (len(ts), 3))
        gyro = g.normal(0, 1e-3, (len(ts), 3))
        i0 = time.perf_counter()
        f = h.imu_preintegrate_batch(ts, acc, gyro, 0.1 * np.arange(n_scans - 1) + 0.0017, 0.1 * np.arange(1, n_scans) + 0.0017)
        imu_s = time.perf_counter() - i0
        sum6, _ = h.online_pose()
    leg = {"metric": "LiDAR scans/s and ms/scan, online VLP-16 odometry + degeneracy (+ IMU factors)", "n_scans": n_scans,
           "value": round(n_scans / wall, 1), "unit": "scans/s", "p50_ms_per_scan": round(_pct(lat[1:], 0.5), 4), "p95_ms_per_scan": round(_pct(lat[1:], 0.95), 4),
           "imu_factors": int(len(f)), "imu_batch_ms": round(imu_s * 1e3, 3), "degenerate": n_deg, "dropped_by_dopt_gate": n_drop, "soft_status": soft,
           "final_transform_sum": [round(float(v), 4) for v in sum6],
           "what": "vlo_process_scan per sweep with pinned host clouds (H2D + organise + extract + scan-to-scan + maintained-map LaserMapping "
                   "every 2nd sweep + D2H), rolling-shutter sweeps with de-skew, range noise 0.02 m"}
    if with_cpu:
        from oracle import oracle as orc
        ocfg = orc.default_config("VLP-16", deskew=1)
        n_c = min(n_scans, 24)
        c0 = time.perf_counter()
        prev = None
        for k in range(n_c):
            c, rs, _ = orc.organise(ocfg, scans[k])
            fe = orc.extract(ocfg, c, rs)
            if prev is not None:
                pc, pf = prev
                orc.odometry_register(ocfg, c[fe["sharp_idx"]], c[fe["flat_idx"]], pc[pf["less_sharp_idx"]], pf["less_sharp_ring_start"],
                                      pf["less_flat"], pf["less_flat_ring_start"], use_kdtree=True)
            prev = (c, fe)
        dt = time.perf_counter() - c0
        leg["cpu_baseline"] = {"value": round(n_c / dt, 2), "unit": "scans/s", "cores": 1, "kind": "port",
                               "sample": "first %d sweeps: organise + extract + scan-to-scan (kd-tree) of the oracle, %.1f s" % (n_c, dt)}
    return leg


def pcie_ceiling_leg(host, dev, barrier, world):
    """What bounds e2e on N GPUs: every rank copies its pinned batch host -> device at the same time, nothing else running."""
    import torch
    import torch.distributed as dist
    best = None
    for _ in range(4):
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        dev.copy_(host, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        t_ms = c0.elapsed_time(c1)
        if world > 1:
            tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t_ms = float(tt[0])
        best = t_ms if best is None else min(best, t_ms)
    return host.numel() * 4 / (best * 1e-3) / 1e9


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from vil_sensor_fusion_b200 import api, bag, synth, synth_gpu

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
    legs = set(x for x in args.legs.split(",") if x and x != "none") if args.legs else set()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"

    B = args.batch
    cm, sm = make_map()
    if args.r01_workload:
        pool_n = 8
        raws_pool, seeds_pool = make_workload_r01(pool_n, rank)
    else:
        pool_n = POOL
        raws_pool, seeds_pool = make_workload(pool_n, rank)
    # the resident / pinned buffers hold the pool followed by its first B - 1 scans again: step s registers the B
    # consecutive scans starting at pool index (s * 61) % pool, so every pool scan (and seed) takes part
    n_buf = pool_n + B - 1
    sizes = np.array([raws_pool[k % pool_n].shape[0] for k in range(n_buf)], np.int64)
    offs_all = np.zeros(n_buf + 1, np.int64)
    offs_all[1:] = np.cumsum(sizes)
    # the PointCloud2 payload as the reference's /lidar topic carries it: CARLA publishes x, y, z float32 only (point_step 12;
    # carla_tools/src/carla_to_ros_transforms.py:69-70 reshapes the data to [-1, 3]); --point-floats 4 gives velodyne-style xyzi
    PF = args.point_floats
    global RAW_BYTES
    RAW_BYTES = 4 * PF
    host = torch.empty((int(offs_all[-1]), PF), dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for k in range(n_buf):
        hv[offs_all[k]:offs_all[k + 1]] = raws_pool[k % pool_n][:, :PF]
    dev = host.to("cuda", non_blocking=False)
    seeds_all = np.stack([seeds_pool[k % pool_n] for k in range(n_buf)])
    scans_idx = np.arange(B, dtype=np.int32)

    def window(step):
        w0 = (step * 61) % pool_n
        return w0, (offs_all[w0:w0 + B + 1] - offs_all[w0]).astype(np.int32), seeds_all[w0:w0 + B]

    n_pts = int(np.mean([offs_all[w + B] - offs_all[w] for w in range(pool_n)]))       # points per step (mean over windows)
    h2d_bytes = n_pts * 4 * PF
    d2h_bytes = B * api.RESULT_DTYPE.itemsize

    cfg = api.default_config("HDL-64E", deskew=0, max_scans=B, max_points=131072,
                             max_map_points=int(max(len(cm), len(sm))), device=local_rank)
    if os.environ.get("VLO_MAP_CELL"):
        cfg.map_cell_size = float(os.environ["VLO_MAP_CELL"])          # tuning experiments only
    h = api.Handle(cfg)
    h.map_build(cm, sm)
    stream = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))
    # further resident batches of the value loop (each its own stream and workspace, the same map): three batches in flight
    extra_handles = [api.Handle(cfg) for _ in range(2)]
    for hx in extra_handles:
        hx.map_build(cm, sm)
    extra_streams = [torch.cuda.ExternalStream(hx.stream_ptr(), device=torch.device("cuda", local_rank)) for hx in extra_handles]
    value_handles = [h] + extra_handles

    def step(on_device: bool, k: int):
        w0, offs, seeds = window(k)
        base = (dev if on_device else host).data_ptr() + int(offs_all[w0]) * 4 * PF
        h.upload_raw(base, offs, PF, on_device)
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
    res = step(True, 0)
    counts = h.counts()
    nc_ds, ns_ds = h.stack_counts(B)
    q_stack = int(nc_ds.sum() + ns_ds.sum())

    for k in range(max(args.warmup - 1, 0)):
        step(True, k + 1)
    for hx in extra_handles:                   # the other handles' warm-up
        for k in range(max(args.warmup, 1)):
            w0, offs_w, seeds_w = window(k)
            hx.upload_raw(dev.data_ptr() + int(offs_all[w0]) * 4 * PF, offs_w, PF, True)
            hx.organise()
            hx.extract()
            hx.register_map(scans_idx, seeds_w)
    res_pin = torch.empty(args.steps * B * api.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    all_counts = [B * args.steps] * world
    if world > 1:
        # warm-up of the exchange step with the timed region's shapes (NCCL connects lazily on the first collective)
        bag.gather_results(np.concatenate([res] * args.steps), counts=all_counts)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    time.sleep(0.25)
    item = api.RESULT_DTYPE.itemsize

    def enqueue_step(hh, k):
        # a step is only ENQUEUED (vlo_register_map_enqueue: its result records land in pinned memory when the stream gets there)
        w0, offs, seeds = window(k)
        hh.upload_raw(dev.data_ptr() + int(offs_all[w0]) * 4 * PF, offs, PF, True)
        hh.organise()
        hh.extract()
        hh.register_map_enqueue(scans_idx, seeds, res_pin.data_ptr() + k * B * item)

    # ---- serial pass, stage timers on: the K steps on ONE handle, one stream -- every kernel runs alone, its CUDA-event duration
    # is the kernel's own (the `stages` table and the roofline are taken here)
    h.set_profiling(True)
    evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs0.record(stream)
    for k in range(args.steps):
        enqueue_step(h, k)
    h.synchronize()
    evs1.record(stream)
    torch.cuda.synchronize()
    serial_ms = evs0.elapsed_time(evs1)
    stages = h.stage_times()
    h.set_profiling(False)
    # ---- timed region: the same K steps, dealt in turn to THREE handles (three resident batches, three streams): the kernels of
    # consecutive steps overlap -- each of them alone leaves issue slots and warp slots idle (profiles/SUMMARY.md) -- 10 steps:
    # 2.48 ms per step on one stream, 2.02 on two, 1.93 on three, 1.90 on four.  Every step is only enqueued; the host synchronises
    # once after the last one.
    launches0 = sum(hx.launch_count() for hx in value_handles)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for sx in extra_streams:
        sx.wait_event(ev0)
    for k in range(args.steps):
        enqueue_step(value_handles[k % len(value_handles)], k)
    for hx in value_handles:
        hx.synchronize()
    res_all_steps = h.results_finish(np.frombuffer(res_pin.numpy(), api.RESULT_DTYPE).copy())
    all_res = [res_all_steps[k * B:(k + 1) * B] for k in range(args.steps)]
    ev_mid = torch.cuda.Event(enable_timing=True)
    ev_mid.record(stream)
    if world > 1:
        # the ONE exchange step of the whole job: every rank's result records, gathered once over NVLink (one fixed-size
        # all_gather_into_tensor of K * B 480-byte records per rank); nothing collective happens per scan or per batch
        gathered = bag.gather_results(np.concatenate(all_res), counts=all_counts)
        assert len(gathered) == world * args.steps * B
    ev1.record(stream)
    gather_split = dict(bag.LAST_GATHER_MS) if world > 1 else {}
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    steps_only_ms = ev0.elapsed_time(ev_mid)
    launches = sum(hx.launch_count() for hx in value_handles) - launches0
    res_all = np.concatenate(all_res)
    ok = int(np.sum(res_all["status"] == 0))
    mean_iters = float(np.mean(res_all["iterations"]))
    total_iters_per_step = float(np.sum(res_all["iterations"])) / args.steps
    max_iters_per_step = float(np.mean([np.max(r["iterations"]) for r in all_res]))
    # launches of the per-iteration kernels that had work: a launch works while any slot of the batch is still iterating
    # (device-reported iteration counts)
    working_k5 = int(sum(int(np.max(r["iterations"])) for r in all_res))

    # ---- e2e: HOST (pinned) buffers through the streaming C-ABI call (vlo_bag_register_map): every step's clouds
    # cross PCIe and every step's result records come back inside the timed region; the library overlaps the copy
    # of one half-batch with the kernels of the previous one
    HBn = B // 2

    def e2e_batches(n_steps):
        out = []
        for k in range(n_steps):
            w0, offs, seeds = window(k)
            base = host.data_ptr() + int(offs_all[w0]) * 4 * PF
            out.append((base, offs[:HBn + 1], seeds[:HBn]))
            out.append((base + int(offs[HBn]) * 4 * PF, offs[HBn:] - offs[HBn], seeds[HBn:]))
        return out

    def e2e_pass(n_steps):
        r = h.bag_register_map(e2e_batches(n_steps), stride=PF)
        if world > 1:
            bag.gather_results(r, counts=[B * n_steps] * world)      # the job's single exchange step, inside the timed region
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
    if not np.array_equal(res_e2e["transform"].view(np.uint32), res_all["transform"].view(np.uint32)):
        print("warning: streaming e2e results differ from the resident-batch results", file=sys.stderr)
    clocks.stop()
    # what bounds e2e: all ranks copying their pinned batch host -> device at the same time, nothing else running
    pcie_gbs = None
    try:
        nb = int(offs_all[B]) * 4
        pcie_gbs = pcie_ceiling_leg(host[:int(offs_all[B])], dev[:int(offs_all[B])], barrier, world)
    except Exception as e:           # noqa: BLE001
        print("warning: PCIe ceiling leg failed: %r" % (e,), file=sys.stderr)

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
        ts = torch.tensor([steps_only_ms, -steps_only_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        steps_ms_ranks = (float(-ts[1]), float(ts[0]))          # fastest and slowest rank's own steps (before the exchange)
    else:
        steps_ms_ranks = (steps_only_ms, steps_only_ms)
    ms_per_step = dev_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
    h.close()
    del dev

    # ---- the other BASELINE configs, each with its own numbers
    pairs_leg = imu_leg = vlp_leg = None
    with_cpu = world == 1 and not args.no_cpu_baseline
    if "c5" in legs:
        pairs_leg = whole_bag_pairs_leg(args, rank, world, local_rank, barrier=lambda: (torch.cuda.synchronize(), world > 1 and dist.barrier()), peak=peak)
    if rank == 0 and "c4" in legs:
        imu_leg = imu_batch_leg(args, local_rank, peak, with_cpu)
    if rank == 0 and "c1" in legs:
        vlp_leg = vlp16_online_leg(args, local_rank, with_cpu)

    # ---- online latency (single HDL-64 scan per call): p50 / p95 of vlo_process_scan on consecutive sweeps
    p50 = p95 = None
    extra_lat = {}
    if rank == 0 and "latency" in legs:
        scene = synth_gpu.make_scene(synth.scene_room(0))
        sensor = synth_gpu.make_sensor("HDL-64E", noise_sigma=0.01)
        npts = sensor.rings * sensor.n_az
        SEQ = 16
        sbuf = torch.empty((SEQ, npts, 4), dtype=torch.float32, device="cuda")
        synth_gpu.synth_scans(scene, sensor, 0, SEQ, 99, sbuf.data_ptr())
        torch.cuda.synchronize()
        spin = torch.empty((SEQ, npts, 4), dtype=torch.float32).pin_memory()       # the caller's scan buffers are pinned (a driver's DMA buffers)
        spin.copy_(sbuf)
        del sbuf
        seq = spin.numpy()
        Rs, ps = synth_gpu.poses(sensor, 0, SEQ)
        pose0 = synth.loam_map_pose(Rs[0], ps[0]).astype(np.float32)
        cfg1 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, io_ratio=1,
                                  max_map_points=int(max(len(cm), len(sm))), device=local_rank)
        with api.Handle(cfg1) as h1:
            h1.map_build(cm, sm)
            lat = []
            for rep in range(4):
                for k in range(SEQ):
                    if k == 0:
                        h1.lib.vlo_online_reset(h1._h)
                        h1.online_set_map_pose(pose0)
                    t1 = time.perf_counter()
                    h1.process_scan(seq[k], 0.1 * k, want_map=True)
                    if rep > 0 and k > 0:
                        lat.append((time.perf_counter() - t1) * 1e3)
            p50, p95 = _pct(lat, 0.5), _pct(lat, 0.95)
            # scan-to-map registration + degeneracy alone (north-star target: p50 < 1 ms): the scan is resident and
            # extracted, the call uploads the seed, registers against the 1M-point map and reads the record back
            h1.upload([raws_pool[k] for k in range(2)])
            h1.organise()
            h1.extract()
            h1.register_map([0], [seeds_pool[0]])
            lat_map, it_map = [], []
            for rep in range(40):
                k = rep % 2
                t1 = time.perf_counter()
                r1 = h1.register_map([k], [seeds_pool[k]])
                lat_map.append((time.perf_counter() - t1) * 1e3)
                it_map.append(int(r1["iterations"][0]))
            extra_lat["scan_to_map_p50_ms"] = round(_pct(lat_map, 0.5), 4)
            extra_lat["scan_to_map_p95_ms"] = round(_pct(lat_map, 0.95), 4)
            extra_lat["scan_to_map_gn_iterations"] = sorted(set(it_map))
        # the same tick with the MAINTAINED map (BasicLaserMapping::process: sub-map selection, optimisation, insertion)
        cfg2 = api.default_config("HDL-64E", deskew=0, max_scans=2, max_points=131072, io_ratio=1,
                                  max_map_points=1 << 20, device=local_rank)
        with api.Handle(cfg2) as h2:
            lat2 = []
            for rep in range(3):
                h2.lib.vlo_online_reset(h2._h)
                h2.map_reset()
                h2.map_insert(cm, sm, np.zeros(6, np.float32))          # prior map: the same 1M points, voxel-filtered
                h2.online_set_map_pose(pose0)
                for k in range(SEQ):
                    t1 = time.perf_counter()
                    h2.process_scan(seq[k], 0.1 * k, want_map=True)
                    if rep > 0 and k > 0:
                        lat2.append((time.perf_counter() - t1) * 1e3)
            extra_lat["maintained_map_tick_p50_ms"] = round(_pct(lat2, 0.5), 4)
            extra_lat["maintained_map_tick_p95_ms"] = round(_pct(lat2, 0.95), 4)
            extra_lat["maintained_map_points"] = list(h2.map_size())

    if rank == 0:
        n_map_pts = len(cm) + len(sm)
        table = {}
        for name, (ms, n) in stages.items():
            if n == 0:
                continue
            working, ab = n, None
            if name in ("k5_assoc", "k5_lin"):
                # map_max_iterations launches per call; a launch works on the slots still iterating.  Algorithmic bytes are
                # counted per (slot, iteration) actually executed (the device reports every slot's iteration count); the map
                # is read once per launch that has any work (= the slowest slot's iteration count)
                working = max(1, working_k5)
                q_slot = q_stack / B
                per_q = (16 + 5 * 4) if name == "k5_assoc" else (16 + 5 * 4 + 5 * 16)
                ab_total = total_iters_per_step * args.steps * q_slot * per_q
                if name == "k5_assoc":
                    ab_total += working * n_map_pts * 16
                else:
                    ab_total += total_iters_per_step * args.steps * 2 * (q_slot / 32 + 1) * 28 * 4
                ab = ab_total / working
            else:
                ab = algorithmic_bytes(name, counts, n_map_pts, mean_iters, q_stack)
            avg_ms = ms / working
            table[name] = {"ms_total": round(ms, 4), "launches": n, "working_launches": working, "avg_ms": round(avg_ms, 5),
                           "algorithmic_bytes": int(ab), "achieved_gbs": round(ab / (avg_ms * 1e-3) / 1e9, 2) if avg_ms > 0 else None,
                           "frac": round(ab / (avg_ms * 1e-3) / 1e9 / peak, 4) if avg_ms > 0 else None}
        dom = max(table, key=lambda k: table[k]["ms_total"])
        traffic, traffic_note = dram_traffic_of(dom)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": table[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": table[dom]["frac"], "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                    "share_of_step": round(table[dom]["ms_total"] / serial_ms, 4),
                    "note": "algorithmic bytes (SURVEY 8d) over the measured HBM peak, as the contract asks; the scan-to-map kernels work on an "
                            "L2-resident map and are issue / latency bound (profiles/SUMMARY.md)"}
        out = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "hdl64_scan_to_map_1M: HDL-64-shaped scans (64x1800) -> organise + feature extraction + "
                                   "scan-to-map registration (<=10 GN iterations, 5-NN on a 1M-point voxel-hash map) + eigen-degeneracy "
                                   "+ D-opt gate", "scans_per_step_per_gpu": B, "points_per_scan": int(n_pts // B),
                       "point_step": 4 * PF, "payload": "x y z float32 (CARLA /lidar clouds, carla_to_ros_transforms.py:69-70)" if PF == 3 else "x y z intensity float32",
                       "map_points": int(n_map_pts), "parallelism": "frame-range dp%d" % world,
                       "l2": "inputs %.0f MB per step > 126 MB L2" % (h2d_bytes / 1e6), "mean_gn_iterations": round(mean_iters, 2),
                       "gn_iterations_hist": np.bincount(res_all["iterations"], minlength=cfg.map_max_iterations + 1).tolist(),
                       "pool": ("%d distinct scans at poses perturbed by N(0, 0.1 m) / N(0, 0.5 deg), seeds 1..%d; a step takes %d consecutive pool "
                                "entries from a start that moves by 61 per step" % (pool_n, pool_n, B)) if not args.r01_workload else "round-1 pool: 8 scans, one fixed offset",
                       "exchange": "one all_gather_into_tensor of the result records per job" if world > 1 else "none",
                       "value_loop": "K steps dealt in turn to three handles (three resident batches on three streams, kernels of consecutive steps "
                                     "overlap), every step only enqueued, one host synchronisation after the last",
                       "serial_pass": {"ms_per_step": round(serial_ms / args.steps, 4), "what": "the same K steps on one handle / one stream with the stage "
                                       "timers on: `stages` and `roofline` are taken there, every kernel running alone"},
                       "timed_region_ms": {"steps": round(steps_only_ms, 3), "exchange_and_wait_for_slowest_rank": round(dev_ms - steps_only_ms, 3),
                                           "steps_fastest_rank": round(steps_ms_ranks[0], 3), "steps_slowest_rank": round(steps_ms_ranks[1], 3),
                                           "exchange_rank0_wall_ms": {k: round(v, 3) for k, v in gather_split.items()}},
                       "numa_node": numa},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": round(e2e_ms / args.steps, 4), "wall_ms": round(e2e_wall_ms, 3), "device_ms": round(e2e_dev_ms, 3),
                    "passes_ms_per_step": [round(v / args.steps, 4) for v in e2e_all], "reported": "median of 3 passes of K steps",
                    "pcie_h2d_gbs_all_ranks_concurrently": None if pcie_gbs is None else round(pcie_gbs, 2),
                    "h2d_gbs_achieved": round(h2d_bytes / (e2e_ms / args.steps * 1e-3) / 1e9, 2),
                    "frac_of_pcie": None if not pcie_gbs else round(h2d_bytes / (e2e_ms / args.steps * 1e-3) / 1e9 / pcie_gbs, 4),
                    "bound": "PCIe host->device copy of the %d B/point PointCloud2 payload (kernels overlap it); the ceiling is measured with " % (4 * PF) +
                             "every rank copying at the same time"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": roofline,
            "stages": table,
            "latency": {"p50_ms_per_scan": None if p50 is None else round(p50, 4), "p95_ms_per_scan": None if p95 is None else round(p95, 4),
                        "what": "vlo_process_scan: one online tick (H2D from a pinned buffer + organise + extract + scan-to-scan + scan-to-map on every sweep (ioRatio 1) + results D2H)",
                        **extra_lat},
            "whole_bag_pairs": pairs_leg,
            "imu_batch": imu_leg,
            "vlp16_online": vlp_leg,
            "ok_registrations": ok, "registrations": int(len(res_all)),
            "mean_corr": [float(np.mean(res_all["n_corr_edge"])), float(np.mean(res_all["n_corr_plane"]))],
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(raws_pool, seeds_pool, cm, sm, threads=1, n_scans=args.cpu_sample)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dram_traffic_of(kernel: str):
    """dram__bytes_read + dram__bytes_write per launch of `kernel` from the committed ncu capture (profiles/dram_traffic.json,
    written by tools/ncu_traffic.py with the sha256 of the kernel's source file); a capture of an older source is not used."""
    import hashlib
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
        e = tr.get(kernel)
        if not e:
            return None, "no ncu capture of %s in profiles/dram_traffic.json" % kernel
        src = os.path.join(ROOT, "vil_sensor_fusion_b200", "csrc", e.get("source", ""))
        cur = hashlib.sha256(open(src, "rb").read()).hexdigest() if e.get("source") and os.path.exists(src) else None
        if cur is None or cur != e.get("source_sha256"):
            return None, "STALE: the ncu capture of %s (%s) was taken on an older %s" % (kernel, e.get("profile"), e.get("source"))
        return e.get("dram_bytes_per_launch"), "ncu --set full, %s" % e.get("profile")
    except Exception as ex:      # noqa: BLE001
        return None, "profiles/dram_traffic.json unreadable: %r" % (ex,)


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
    cm, sm = make_map()
    raws_pool, seeds_pool = make_workload_r01(8, 0) if args.r01_workload else make_workload(POOL, 0)
    from oracle import oracle as orc
    cfg = orc.default_config("HDL-64E", deskew=0)
    m = orc.CpuMap(cfg, cm, sm)
    n_per_step = max(threads, 4) * 2
    raws = [raws_pool[k % len(raws_pool)] for k in range(n_per_step)]
    seeds = np.stack([seeds_pool[k % len(raws_pool)] for k in range(n_per_step)])
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
    ap.add_argument("--legs", default=None, help="comma list of the extra legs: c5 (whole-bag pairs, every rank), c4 (IMU batch), c1 (VLP-16 online), "
                                                 "latency (HDL-64 online tick); default: all at N=1, c5 at N>1")
    ap.add_argument("--bag-pairs", type=int, default=20000, help="scan pairs of the whole-bag job (SURVEY C5)")
    ap.add_argument("--bag-batch", type=int, default=256, help="frames resident per batch of the whole-bag job")
    ap.add_argument("--bag-seconds", type=float, default=1.0, help="minimum length of the whole-bag timed region (the job is repeated)")
    ap.add_argument("--imu-factors", type=int, default=10000)
    ap.add_argument("--vlp16-scans", type=int, default=600)
    ap.add_argument("--point-floats", type=int, default=3, choices=[3, 4],
                    help="float32 fields per raw point of the headline workload: 3 = x y z (point_step 12, the CARLA clouds the reference's "
                         "/lidar topic carries), 4 = x y z intensity (point_step 16, velodyne-style; round 1's payload)")
    ap.add_argument("--r01-workload", action="store_true", help="round 1's pool (8 scans, one fixed seed offset) instead of SURVEY C2's 200 perturbed poses")
    args = ap.parse_args()
    if args.legs is None:
        args.legs = "c5" if args.gpus > 1 else "c5,c4,c1,latency"
    if args.no_latency:
        args.legs = ",".join(x for x in args.legs.split(",") if x != "latency")
    args.warmup = max(args.warmup, 3) if args.impl == "vlo" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
