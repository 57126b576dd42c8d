"""Offline whole-bag reprocessing: independent scan-pair (or scan-to-map) registrations over a frame
range, sharded by contiguous frame range across ranks (one process per GPU), with ONE exchange step:
the final gather of the per-frame result records.  No collective in the per-scan path.

The gathered `(6,6,T)` Hessian stack is what the reference's analysis tooling consumes
(vil_fusion/python/make_prettier_graphs.py:411-474 `numpify_diagnostics`, :547-576 `apply_degen_function`).
"""
from __future__ import annotations

import numpy as np

from .api import RESULT_DTYPE, Handle, hessian_stack  # noqa: F401


def frame_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: [r*N/G, (r+1)*N/G)."""
    return (n_frames * rank) // world, (n_frames * (rank + 1)) // world


def pair_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range [lo, hi) of the scan PAIRS (k, k+1), 0 <= k < n_frames - 1, rank `rank` registers.  The rank loads
    frames lo .. hi inclusive: the pair that crosses a shard boundary belongs to the lower rank, so the gathered
    records are exactly the n_frames - 1 pairs in frame order (sharding FRAMES instead would drop one pair per
    boundary)."""
    n_pairs = max(n_frames - 1, 0)
    return (n_pairs * rank) // world, (n_pairs * (rank + 1)) // world


def reprocess_pairs_sharded(h: Handle, get_scan, n_frames: int, rank: int, world: int, batch: int, seeds=None, group=None) -> np.ndarray:
    """Whole-bag scan-to-scan reprocessing on `world` ranks: this rank registers its pair_range, one gather at the end;
    every rank returns all n_frames - 1 records in frame order."""
    lo, hi = pair_range(n_frames, rank, world)
    local = reprocess_pairs(h, get_scan, lo, hi + 1, batch, None if seeds is None else seeds[lo:hi]) if hi > lo else np.zeros(0, RESULT_DTYPE)
    counts = [pair_range(n_frames, r, world)[1] - pair_range(n_frames, r, world)[0] for r in range(world)]
    return gather_results(local, group=group, counts=counts)


def reprocess_pairs(h: Handle, get_scan, first: int, last: int, batch: int, seeds=None) -> np.ndarray:
    """Registers every consecutive pair (k, k+1) with first <= k < last-1 from `seeds[k]` (or zero):
    frames are processed in resident batches of `batch` scans, overlapping by one frame.
    get_scan(k) -> float32 (n, stride) raw cloud.  Returns a RESULT_DTYPE array of length last-first-1."""
    out = []
    k = first
    while k < last - 1:
        hi = min(last, k + batch)
        scans = [get_scan(i) for i in range(k, hi)]
        h.upload(scans)
        h.organise()
        h.extract()
        n = hi - k
        sd = None if seeds is None else np.asarray(seeds[k - first:k - first + n - 1], np.float32)
        out.append(h.register_pairs(np.arange(n - 1), np.arange(1, n), seeds=sd))
        k = hi - 1
    return np.concatenate(out) if out else np.zeros(0, RESULT_DTYPE)


def reprocess_map(h: Handle, get_scan, first: int, last: int, batch: int, seeds) -> np.ndarray:
    """Scan-to-map registration of frames [first, last) against the handle's resident map."""
    out = []
    for k in range(first, last, batch):
        hi = min(last, k + batch)
        h.upload([get_scan(i) for i in range(k, hi)])
        h.organise()
        h.extract()
        out.append(h.register_map(np.arange(hi - k), np.asarray(seeds[k - first:hi - first], np.float32)))
    return np.concatenate(out) if out else np.zeros(0, RESULT_DTYPE)


def reprocess_pairs_from_bag(h: Handle, path: str, cloud_topic: str = "/lidar", first: int = 0, last: int | None = None,
                             batch: int = 64) -> tuple[np.ndarray, np.ndarray]:
    """Scan-to-scan registration of the consecutive PointCloud2 messages [first, last) of a ROS bag (v2.0, read by
    rosbag_io without ROS; payloads go to the device as they are, x / y / z picked by their field offsets).
    Returns (results for the pairs (k, k+1), message stamps of frames first .. last-1)."""
    from . import rosbag_io
    clouds, _ = rosbag_io.load_bag(path, cloud_topic)
    last = len(clouds) if last is None else min(last, len(clouds))
    out = []
    k = first
    while k < last - 1:
        hi = min(last, k + batch)
        h.upload_pointcloud2(clouds[k:hi])
        h.organise()
        h.extract()
        n = hi - k
        out.append(h.register_pairs(np.arange(n - 1), np.arange(1, n)))
        k = hi - 1
    stamps = np.array([c["stamp"] for c in clouds[first:last]])
    return (np.concatenate(out) if out else np.zeros(0, RESULT_DTYPE)), stamps


_GATHER_BUFFERS: dict = {}
LAST_GATHER_MS: dict = {}          # wall-clock split of the last gather_results call on this rank (bench.py reports it)


def gather_results(local: np.ndarray, group=None, device=None, counts=None) -> np.ndarray:
    """The single exchange step: all ranks contribute their result records, every rank receives the
    concatenation in rank (= frame) order.  Works on NCCL (device tensors over NVLink) and gloo (CPU).

    With equal shares the result is a view of a staging buffer that the next call with the same shapes overwrites (copy it to keep it).

    counts: records per rank when every rank can derive them (e.g. from pair_range) -- the exchange is then ONE
    fixed-size all_gather_into_tensor of the padded record blocks with no count round trip and no host
    synchronisation before the collective; without it one small all_gather of the counts precedes it."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.copy()
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu"))
    if counts is None:
        n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=dev)
        all_n = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_n, n_local, group=group)
        counts = all_n.tolist()
    counts = [int(c) for c in counts]
    assert counts[dist.get_rank(group)] == local.shape[0], "counts[rank] must equal the number of local records"
    nmax = max(max(counts), 1)
    item = RESULT_DTYPE.itemsize
    # staging buffers are kept between calls (a pinned allocation costs about a millisecond: more than the exchange itself)
    key = (dev.type, dev.index, world, nmax)
    bufs = _GATHER_BUFFERS.get(key)
    if bufs is None:
        pin = dev.type == "cuda"
        bufs = (torch.zeros(nmax * item, dtype=torch.uint8, pin_memory=pin), torch.empty(nmax * item, dtype=torch.uint8, device=dev),
                torch.empty(world * nmax * item, dtype=torch.uint8, device=dev), torch.empty(world * nmax * item, dtype=torch.uint8, pin_memory=pin))
        _GATHER_BUFFERS.clear()
        _GATHER_BUFFERS[key] = bufs
    import time
    mine_host, mine, everything, all_host = bufs
    t0 = time.perf_counter()
    mine_host.numpy()[:local.shape[0] * item] = np.frombuffer(np.ascontiguousarray(local).data, np.uint8)
    mine.copy_(mine_host, non_blocking=True)
    t1 = time.perf_counter()
    dist.all_gather_into_tensor(everything, mine, group=group)
    all_host.copy_(everything, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    t2 = time.perf_counter()
    flat = all_host.numpy()
    total = sum(counts)
    if all(c == nmax for c in counts):
        # equal shares: the gathered block IS the record array (a view of the pinned staging buffer, valid until the next call with
        # these shapes -- copying 8 x 1280 records into fresh pages costs more than the collective)
        out = flat[:total * item].view(RESULT_DTYPE)
    else:
        out = np.empty(total, RESULT_DTYPE)
        ob = out.view(np.uint8).reshape(-1)
        pos = 0
        for r, c in enumerate(counts):
            ob[pos * item:(pos + c) * item] = flat[r * nmax * item:r * nmax * item + c * item]
            pos += c
    LAST_GATHER_MS.update(stage_in=(t1 - t0) * 1e3, collective_and_copy_back=(t2 - t1) * 1e3, unpack=(time.perf_counter() - t2) * 1e3)
    return out
