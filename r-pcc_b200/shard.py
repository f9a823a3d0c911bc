"""Frame sharding across ranks and the metrics gather -- the only collective on this path.

Frames are independent (no inter-frame prediction anywhere in R-PCC), so batch compression shards
whole frames: rank r of `world` takes a contiguous chunk of the datalist (I/O locality) and runs
its own encoder on its own GPU; the hot path has no data-path collective.  One all_gather of the
per-frame metrics table at the end (NCCL on GPUs, gloo in the CPU tests)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def parse_cpulist(text):
    """'0-15,32-47' (the kernel's cpulist format) -> set of CPU numbers."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(device_index, sysfs="/sys/bus/pci/devices"):
    """One process per GPU: keep this process -- its reader / entropy-coder threads and the pinned staging buffers it
    allocates from now on (first touch) -- on the CPUs next to its GPU.  On an 8-GPU box the host <-> device copies of a
    rank whose buffers sit on the other socket cross the inter-socket link (measured: 23 GB/s against 35 GB/s per GPU with
    every rank copying at once).  Returns the CPU set applied, or None when the topology is not exposed (containers,
    single-node machines) or the CPUs are not ours to choose; never raises."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        if hasattr(p, "pci_bus_id"):
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        else:                                      # older torch: ask NVML for the same device by UUID
            import pynvml
            pynvml.nvmlInit()
            info = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByUUID(("GPU-%s" % p.uuid).encode()))
            bus = info.busId.decode() if isinstance(info.busId, bytes) else info.busId
            bdf = bus.lower()[-12:]                # 00000000:1B:00.0 -> 0000:1b:00.0
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as fh:
            local = parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        mine = local & allowed
        if not mine or mine == allowed:
            return None
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


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
