"""python -m rpcc_b200.tools.decompress_datalist --datalist L.txt --output_dir OUT --lidar Velodyne64E
Batched mirror of the reference's tools/decompress_datalist.py:48-134; datalist lines are .rpcc files,
outputs are .bin files under output_dir (the extension text replaced as the reference does, :129)."""
import concurrent.futures as futures
import os
import time

import numpy as np
import torch

from ..batch import BatchDecoder
from ..shard import shard_range
from .common import base_parser, resolve


def decompress(args):
    cfg, accuracy, segment_cfg, model_cfg, uniform, method = resolve(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    files = [l.strip() for l in open(args.datalist) if l.strip()]
    lo, hi = shard_range(len(files), rank, world)
    mine = files[lo:hi]
    dec = BatchDecoder(args.lidar, accuracy=accuracy / 2, nonuniform=not uniform, compressor_cfg=dict(cfg),
                       basic_compressor=method, workers=args.workers)
    t0 = time.time()
    pool = futures.ThreadPoolExecutor(max(1, args.workers))
    for b0 in range(0, len(mine), args.batch):
        names = mine[b0:b0 + args.batch]
        blobs = list(pool.map(lambda n: open(n, "rb").read(), names))
        out = dec.decode(blobs, want_xyz=True)
        xyz = out["xyz"].cpu().numpy()

        def save(j):
            # dataset/dataset.py:72-81 (save_point_cloud_to_file): drop x + y + z == 0, append a zero intensity
            n = names[j]
            fn = n[1:] if n[0] == "/" else n
            path = os.path.join(args.output_dir, fn)
            path = path.replace(path.split(".")[-1], "bin")
            os.makedirs(os.path.dirname(path), exist_ok=True)
            pc = xyz[j].reshape(-1, 3)
            pc = pc[np.where(np.sum(pc, -1) != 0)]
            np.concatenate((pc, np.zeros((pc.shape[0], 1), np.float32)), -1).astype(np.float32).tofile(path)

        list(pool.map(save, range(len(names))))     # numpy releases the GIL in these array passes
    pool.shutdown()
    if rank == 0:
        dt = time.time() - t0
        print("Decompressed %d frames in %.2f s" % (len(mine), dt))


def main(argv=None):
    args = base_parser(single=False).parse_args(argv)
    return decompress(args)


if __name__ == "__main__":
    main()
