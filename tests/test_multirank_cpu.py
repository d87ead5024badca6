"""The N > 1 host logic on CPU: two gloo ranks shard game ids disjointly, reduce timings with MAX
and counts with SUM exactly as bench.py does.  (The GPU path itself has no collective.)"""
import importlib.util
import os
import socket

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_sharding():
    spec = importlib.util.spec_from_file_location("azb_sharding", os.path.join(ROOT, "alphazero-rs_b200", "sharding.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _worker(rank, world, port, out):
    import torch.distributed as dist
    sh = load_sharding()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ids = set()
    for step in range(3):
        first, n = sh.shard(step, rank, world, 4096)
        ids |= set(range(first, first + n))
    mx = sh.reduce_scalar(dist, 100.0 + rank, "MAX")     # device time: max over ranks
    sm = sh.reduce_scalar(dist, 1000 * (rank + 1), "SUM")  # simulations: summed
    lo, hi = min(ids), max(ids)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, len(ids)))
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, mx, sm, gathered))


def test_two_rank_sharding_and_reductions():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, mx, sm, gathered in res:
        assert mx == 101.0 and sm == 3000.0
        assert [g[2] for g in gathered] == [3 * 4096, 3 * 4096]
    sh = load_sharding()
    all_ids = set()
    for step in range(3):
        for rank in range(2):
            first, n = sh.shard(step, rank, 2, 4096)
            r = set(range(first, first + n))
            assert not (all_ids & r)
            all_ids |= r
    assert all_ids == set(range(3 * 2 * 4096))


def test_split_total_covers_everything():
    sh = load_sharding()
    for total, world in ((16384, 8), (65536, 8), (41, 4), (3, 8)):
        seen = []
        for r in range(world):
            first, n = sh.split_total(total, r, world)
            seen += list(range(first, first + n))
        assert seen == list(range(total))


def _grad_worker(rank, world, port, out):
    import numpy as np
    import torch.distributed as dist
    sh = load_sharding()
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = np.arange(1000, dtype=np.float32) * (rank + 1)       # rank r holds (r + 1) * base
    avg = sh.average_gradients(dist, g)
    boards = np.arange(41 * 84, dtype=np.float32).reshape(41, 2, 6, 7)
    shard = sh.shard_samples((boards, np.zeros((41, 7), np.float32), np.arange(41, dtype=np.float32)), rank, world)
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, avg[:4].tolist(), float(avg.sum()), shard[2].tolist()))


def test_two_rank_gradient_average_and_sample_shards():
    """The data-parallel training step's only collective (NNet.train(dist=...)): all-reduce + divide by the world size;
    and the sample sharding that goes with it."""
    import numpy as np
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    base = np.arange(1000, dtype=np.float32)
    for rank, head, total, vs in res:
        assert head == (1.5 * base[:4]).tolist() and total == float((1.5 * base).sum())   # mean of 1x and 2x
        assert vs == list(range(rank * 20, rank * 20 + 20))                               # 41 samples -> 20 + 20, 1 dropped
