"""python -m rpcc_b200.tools.compress_datalist --datalist L.txt --output_dir OUT --lidar Velodyne64E
   torchrun --nproc-per-node 8 -m rpcc_b200.tools.compress_datalist ...          (one rank per GPU)

The reference's tools/compress_datalist.py:48-206 re-plumbed: instead of a thread pool over a
per-frame closure, frames are sharded contiguously across ranks (no collective on the hot path),
each rank pushes batches of frames through the GPU chain (BatchEncoder.encode_host) and entropy
codes the sections on `--workers` host threads while the next batch is on the device.  Output paths
follow the reference: output_dir + input path with the extension text replaced by 'rpcc'
(compress_datalist.py:136-141, including its replace-everywhere quirk)."""
import concurrent.futures as futures
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from ..batch import BatchEncoder
from ..compress_utils import BasicCompressor, pack_bitstream
from ..shard import gather_metrics, shard_range
from .common import base_parser, resolve


def output_path_for(output_dir, file_name):
    file_name = file_name.strip()
    if file_name[0] == "/":
        file_name = file_name[1:]
    output_path = os.path.join(output_dir, file_name)
    return output_path.replace(output_path.split(".")[-1], "rpcc")


def load_points(path):
    ext = path.split(".")[-1]
    if ext == "bin":
        return np.fromfile(path, dtype=np.float32).reshape(-1, 4)
    if ext in ("npy", "npz"):
        a = np.load(path)
    elif ext == "txt":
        a = np.loadtxt(path)
    else:
        raise ValueError("File type not correct: " + path)
    return np.ascontiguousarray(a[:, :3], dtype=np.float32)


def compress(args):
    cfg, accuracy, segment_cfg, model_cfg, uniform, method = resolve(args)
    if model_cfg["model_method"] not in ("point", "plane") or segment_cfg["segment_method"] != "FPS":
        raise NotImplementedError("the batched driver covers FPS segmentation with point or plane modelling")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    files = [l.strip() for l in open(args.datalist) if l.strip()]
    lo, hi = shard_range(len(files), rank, world)
    mine = files[lo:hi]
    cfg = dict(cfg)
    cfg["cluster_num"] = segment_cfg["cluster_num"]
    cfg["ground_threshold"] = segment_cfg["ground_vertical_threshold"]
    cfg["plane_angle_threshold"] = model_cfg["angle_threshold"]
    enc = BatchEncoder(args.lidar, accuracy=accuracy / 2, nonuniform=not uniform, compressor_cfg=cfg,
                       max_batch=min(args.batch, max(len(mine), 1)), device=local, basic_compressor=method,
                       workers=args.workers, model_method=model_cfg["model_method"])
    bc = BasicCompressor(method_name=method)
    metrics = np.zeros((len(mine), 3), np.float64)  # bytes, valid pixels, seconds (amortised)
    pool = futures.ThreadPoolExecutor(args.workers)
    pending = []

    def entropy_and_save(i, sections, valid):
        blob = pack_bitstream({k: bc.compress(v, section=k) for k, v in sections.items()}, uniform=uniform)
        out = output_path_for(args.output_dir, mine[i])
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "wb") as f:
            f.write(blob)
        metrics[i, 0] = len(blob)
        metrics[i, 1] = valid

    t0 = time.time()
    B = enc.max_batch
    for b0 in range(0, len(mine), B):
        names = mine[b0:b0 + B]
        with futures.ThreadPoolExecutor(args.workers) as io:
            clouds = list(io.map(load_points, names))
        strides = {c.shape[1] for c in clouds}
        if len(strides) != 1:
            clouds = [np.ascontiguousarray(c[:, :3]) for c in clouds]
        pts = np.concatenate(clouds, 0)
        off = np.cumsum([0] + [c.shape[0] for c in clouds]).astype(np.int64)
        out = enc.encode_host(pts, off, None)
        for j in range(len(names)):
            sec = BatchEncoder.frame_sections(out, j)   # copies out of the pinned buffers
            pending.append(pool.submit(entropy_and_save, b0 + j, sec, int(out["results"]["sym_count"][j])))
    for p in pending:
        p.result()
    pool.shutdown()
    dt = time.time() - t0
    metrics[:, 2] = dt / max(len(mine), 1)
    table = gather_metrics(metrics, len(files), device=torch.device("cuda", local) if world > 1 else None)
    enc.close()
    if rank == 0:
        total_bytes, total_valid = table[:, 0].sum(), table[:, 1].sum()
        print("\nCompressed %d frames on %d GPU(s) in %.2f s (%.1f frames/s incl. file I/O and %s)." %
              (len(files), world, dt, len(files) / dt if dt > 0 else 0.0, method))
        print("    mean BPP: ", 8.0 * total_bytes / max(total_valid, 1))
        print("    mean compression ratio: ", (total_valid * 96.0) / max(8.0 * total_bytes, 1))
    if world > 1:
        dist.barrier()
    return table


def main(argv=None):
    args = base_parser(single=False).parse_args(argv)
    return compress(args)


if __name__ == "__main__":
    main()
