"""World-size-2 (gloo, CPU) test of the N>1 host logic: frame-range partition + the single final
gather of result records, which is the only exchange step of whole-bag reprocessing."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from vil_sensor_fusion_b200 import bag
    from vil_sensor_fusion_b200.api import RESULT_DTYPE
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bag.frame_range(n_frames, rank, world)
    local = np.zeros(hi - lo, RESULT_DTYPE)
    local["iterations"] = np.arange(lo, hi)                     # frame id travels in the record
    local["transform"][:, 5] = np.arange(lo, hi) * 0.5
    local["hessian"][:, 0, 0] = rank + 1
    allr = bag.gather_results(local)
    dist.barrier()
    if rank == 0:
        q.put((allr["iterations"].tolist(), allr["transform"][:, 5].tolist(), allr["hessian"][:, 0, 0].tolist()))
    dist.destroy_process_group()


class _FakeHandle:
    """Stands in for api.Handle on the CPU: a pair's record carries the ids of the two frames it registered."""

    def upload(self, scans):
        self.ids = [int(s[0, 0]) for s in scans]

    def organise(self):
        pass

    def extract(self):
        pass

    def register_pairs(self, last, cur, seeds=None):
        from vil_sensor_fusion_b200.api import RESULT_DTYPE
        out = np.zeros(len(last), RESULT_DTYPE)
        out["iterations"] = [self.ids[i] for i in last]
        out["n_corr_edge"] = [self.ids[i] for i in cur]
        if seeds is not None:
            out["transform"] = seeds
        return out


def _pairs_worker(rank, world, port, n_frames, batch, q):
    import torch.distributed as dist
    from vil_sensor_fusion_b200 import bag
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    get_scan = lambda k: np.full((4, 4), k, np.float32)
    seeds = np.arange(6 * (n_frames - 1), dtype=np.float32).reshape(-1, 6)
    allr = bag.reprocess_pairs_sharded(_FakeHandle(), get_scan, n_frames, rank, world, batch, seeds=seeds)
    dist.barrier()
    if rank == 0:
        q.put((allr["iterations"].tolist(), allr["n_corr_edge"].tolist(), allr["transform"].tolist()))
    dist.destroy_process_group()


def test_pair_ranges_cover_every_pair_once():
    from vil_sensor_fusion_b200 import bag
    for n in (0, 1, 2, 7, 20001):
        for g in (1, 2, 4, 8):
            r = [bag.pair_range(n, k, g) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == max(n - 1, 0)
            assert all(r[k][1] == r[k + 1][0] for k in range(g - 1))


def test_sharded_pairs_equal_the_single_rank_job_gloo():
    """Two ranks, each registering its pair_range (the pair across the shard boundary included), one gather: the records
    are the single-rank job's, pair for pair (sharding frames dropped one pair per boundary)."""
    from vil_sensor_fusion_b200 import bag
    n_frames, batch = 23, 5
    get_scan = lambda k: np.full((4, 4), k, np.float32)
    seeds = np.arange(6 * (n_frames - 1), dtype=np.float32).reshape(-1, 6)
    single = bag.reprocess_pairs(_FakeHandle(), get_scan, 0, n_frames, batch, seeds)
    assert single["iterations"].tolist() == list(range(n_frames - 1)) and single["n_corr_edge"].tolist() == list(range(1, n_frames))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, _free_port()
    procs = [ctx.Process(target=_pairs_worker, args=(r, world, port, n_frames, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    last, cur, tr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert last == single["iterations"].tolist() and cur == single["n_corr_edge"].tolist()
    assert tr == single["transform"].tolist()


def test_frame_ranges_partition_exactly():
    from vil_sensor_fusion_b200 import bag
    for n in (0, 1, 7, 20000):
        for g in (1, 2, 4, 8):
            r = [bag.frame_range(n, k, g) for k in range(g)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(g - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 11, 2, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    it, tz, h00 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert it == list(range(n_frames))
    assert tz == [0.5 * k for k in range(n_frames)]
    assert h00 == [1.0] * 5 + [2.0] * 6


def test_gather_single_process_is_identity():
    from vil_sensor_fusion_b200 import bag
    from vil_sensor_fusion_b200.api import RESULT_DTYPE
    a = np.zeros(3, RESULT_DTYPE)
    a["iterations"] = [4, 5, 6]
    np.testing.assert_array_equal(bag.gather_results(a)["iterations"], [4, 5, 6])
