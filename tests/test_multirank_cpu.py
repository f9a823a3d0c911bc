"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous frame sharding with no
data-path collective, then one all_gather of the per-frame metrics table."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from rpcc_b200.shard import gather_metrics, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # each rank "compresses" its own frames: row i = (frame index, bytes, valid pixels)
    local = np.stack([np.arange(lo, hi), 30000.0 + np.arange(lo, hi), 90000.0 - np.arange(lo, hi)], 1).astype(np.float64)
    table = gather_metrics(local, n_total)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_metrics_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total, world = 37, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([np.arange(n_total), 30000.0 + np.arange(n_total), 90000.0 - np.arange(n_total)], 1)
    for r in range(world):
        assert np.array_equal(got[r], want)
