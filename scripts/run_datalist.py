#!/usr/bin/env python
"""BASELINE configs[4] at a reduced count: N synthetic 64E .bin files on local disk through
rpcc_b200.tools.compress_datalist and decompress_datalist (file I/O + GPU chain + host bz2 threads), with a
byte check of a few outputs against the oracle.   python scripts/run_datalist.py [N] [workers]"""
import io
import json
import os
import shutil
import sys
import time
from contextlib import redirect_stdout

import numpy as np

sys.path.insert(0, ".")
from rpcc_b200 import synthetic  # noqa: E402
from rpcc_b200.tools import compress_datalist, decompress_datalist  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
workers = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, (os.cpu_count() or 2) - 1)
root = "/tmp/pcc_dl"    # no "bin" / "rpcc" in the path: the tools replace the extension text everywhere, as the reference does
shutil.rmtree(root, ignore_errors=True)
os.makedirs(root + "/in")
nd = 32
per = [synthetic.frame(5000 + i, "Velodyne64E") for i in range(nd)]
names = []
for i in range(N):
    p = "%s/in/%06d.bin" % (root, i)
    per[i % nd][0].tofile(p)
    names.append(p)
open(root + "/list.txt", "w").write("\n".join(names) + "\n")
argv = ["--datalist", root + "/list.txt", "--output_dir", root + "/out", "--lidar", "Velodyne64E", "--workers", str(workers),
        "--batch", "592"]
buf = io.StringIO()
t0 = time.time()
with redirect_stdout(buf):
    table = compress_datalist.main(argv)
t_c = time.time() - t0
outs = [compress_datalist.output_path_for(root + "/out", n) for n in names]
open(root + "/list_rpcc.txt", "w").write("\n".join(outs) + "\n")
t0 = time.time()
with redirect_stdout(buf):
    decompress_datalist.main(["--datalist", root + "/list_rpcc.txt", "--output_dir", root + "/dec", "--lidar", "Velodyne64E",
                              "--workers", str(workers), "--batch", "296"])
t_d = time.time() - t0
# the device-fitted ground is part of the stream: decode one frame with the oracle and compare the written .bin
import oracle  # noqa: E402  (checker)
from rpcc_b200.compress_utils import BasicCompressor, parse_bitstream  # noqa: E402
blob = open(outs[3], "rb").read()
sec = BasicCompressor(method_name="bzip2").decompress_dict(parse_bitstream(blob, uniform=True))
rec, xyz, seg = oracle.decompress_sections(sec, "Velodyne64E", 0.02)
pc = xyz.reshape(-1, 3)
pc = pc[pc.sum(-1) != 0]
got = np.fromfile(os.path.join(root + "/dec", outs[3][1:]).replace("rpcc", "bin"), np.float32).reshape(-1, 4)
same = bool(got.shape[0] == pc.shape[0] and np.array_equal(got[:, :3].view(np.uint32), pc.view(np.uint32)))
ri = oracle.project(per[3][0], *oracle.lidar_params("Velodyne64E"))
err = float(np.abs(rec.reshape(ri.shape) - ri)[ri > 0].max())
print(json.dumps({"frames": N, "host_threads": workers, "compress_s": t_c, "compress_frames_per_s": N / t_c,
                  "decompress_s": t_d, "decompress_frames_per_s": N / t_d, "mean_rpcc_bytes": float(table[:, 0].mean()),
                  "decoded_bin_equals_oracle_decode": same, "max_abs_range_error": err,
                  "tool_output": buf.getvalue().strip().splitlines()[-4:]}))
shutil.rmtree(root, ignore_errors=True)
