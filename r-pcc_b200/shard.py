"""Frame sharding across ranks and the metrics gather -- the only collective on this path.

Frames are independent (no inter-frame prediction anywhere in R-PCC), so batch compression shards
whole frames: rank r of `world` takes a contiguous chunk of the datalist (I/O locality) and runs
its own encoder on its own GPU; the hot path has no data-path collective.  One all_gather of the
per-frame metrics table at the end (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of rank's frames; sizes differ by at most one."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_metrics(local_table, n_total, rank=None, world=None, device=None):
    """local_table: (n_local, C) float64 rows for this rank's frames (in datalist order).
    Returns the (n_total, C) table on every rank.  Without an initialised process group (single
    GPU) the local table is returned as is."""
    local = np.ascontiguousarray(local_table, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    C = local.shape[1]
    cap = (n_total + world - 1) // world
    buf = torch.zeros((cap, C), dtype=torch.float64)
    buf[:local.shape[0]] = torch.from_numpy(local)
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    rows = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        rows.append(out[r][:hi - lo].cpu().numpy())
    return np.concatenate(rows, 0)
