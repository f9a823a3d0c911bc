"""GPU tests of the batch tools (BASELINE configs[4]): rpcc_b200.tools.compress_datalist / decompress_datalist against the
oracle, and the invariance the tool promises -- a frame's `.rpcc` bytes depend on its points alone (not on --batch, the
position in the datalist, the number of ranks, or which tool wrote the file).  Mirrors what
reference tools/compress_datalist.py:91-206 and tools/decompress_datalist.py:93-131 do per frame."""
import argparse
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIDAR = "Velodyne64E"
N_FILES = 23


def _args(datalist, outdir, **kw):
    from rpcc_b200.tools.common import base_parser
    a = base_parser(single=False).parse_args(["--datalist", datalist, "--output_dir", outdir, "--lidar", kw.pop("lidar", LIDAR),
                                              "--workers", "4", "--batch", str(kw.pop("batch", 64))])
    for k, v in kw.items():
        setattr(a, k, v)
    return a


@pytest.fixture(scope="module")
def corpus(tmp_path_factory):
    """N_FILES synthetic 64E frames as KITTI .bin files (no 'bin' / 'rpcc' in the directory names: the tools replace the
    extension text everywhere in the path, like the reference)."""
    from rpcc_b200 import synthetic
    root = str(tmp_path_factory.mktemp("dl"))
    os.makedirs(root + "/in")
    names, frames = [], []
    for i in range(N_FILES):
        p, g = synthetic.frame(7000 + (i % 9), LIDAR)
        if i >= 9:                                    # repeated content further down the list: same bytes expected
            p = p.copy()
        path = "%s/in/%06d.bin" % (root, i)
        p.tofile(path)
        names.append(path)
        frames.append(p)
    lst = root + "/list.txt"
    open(lst, "w").write("\n".join(names) + "\n")
    return dict(root=root, names=names, frames=frames, list=lst)


def _outputs(corpus, outdir):
    from rpcc_b200.tools.compress_datalist import output_path_for
    return [open(output_path_for(outdir, n), "rb").read() for n in corpus["names"]]


def _oracle_rpcc(points, nonuniform=False, model_method="point"):
    H, W, hf, vmax, vmin = oracle.lidar_params(LIDAR)
    lut = oracle.transform_map(H, W, hf, vmax, vmin)
    g = oracle.ground_fit(oracle.project(points, H, W, hf, vmax, vmin), lut)      # frame key 0: content only
    out = oracle.compress_frame(points, LIDAR, g, nonuniform=nonuniform, model_method=model_method, plane_impl="device")
    return oracle.write_rpcc(out["sections"]), out


def test_compress_datalist_bytes_do_not_depend_on_batching_or_sharding(corpus):
    from rpcc_b200.tools import compress_datalist
    root = corpus["root"]
    t7 = compress_datalist.compress(_args(corpus["list"], root + "/o7", batch=7))
    t64 = compress_datalist.compress(_args(corpus["list"], root + "/o64", batch=64))
    b7, b64 = _outputs(corpus, root + "/o7"), _outputs(corpus, root + "/o64")
    assert b7 == b64
    assert np.array_equal(t7[:, :2], t64[:, :2])
    assert [len(b) for b in b7] == list(t7[:, 0].astype(int))
    # world 2: each rank's contiguous shard on its own (no process group on a one-GPU box); the union is the same set
    for r in range(2):
        compress_datalist.compress(_args(corpus["list"], root + "/ow2", batch=5), rank=r, world=2, collective=False)
    assert _outputs(corpus, root + "/ow2") == b7
    # the same content at another position of the list gives the same bytes
    for i in range(9, N_FILES):
        assert b7[i] == b7[i % 9], i
    # and they are the oracle's bytes (its restatement of the deterministic ground RANSAC, key 0)
    for i in (0, 4, 8):
        want, _ = _oracle_rpcc(corpus["frames"][i])
        assert b7[i] == want, i


def test_compress_datalist_equals_the_single_file_tool(corpus, capsys):
    from rpcc_b200.tools import compress, compress_datalist
    from rpcc_b200.tools.common import base_parser
    root = corpus["root"]
    compress_datalist.compress(_args(corpus["list"], root + "/o1", batch=3))
    many = _outputs(corpus, root + "/o1")
    for i in (1, 5):
        out = "%s/single_%d.rpcc" % (root, i)
        compress.compress(base_parser(single=True).parse_args(["--input", corpus["names"][i], "--output", out, "--lidar", LIDAR]))
        assert open(out, "rb").read() == many[i], i
    capsys.readouterr()


@pytest.mark.parametrize("mode", ["nonuniform", "plane", "deflate"])
def test_compress_datalist_other_configurations(corpus, mode):
    """non-uniform framework, plane modelling and a coder that runs on the Python pool -- each against the oracle."""
    import gzip
    from rpcc_b200.compress_utils import parse_bitstream
    from rpcc_b200.tools import compress_datalist
    root = corpus["root"]
    small = root + "/small.txt"
    open(small, "w").write("\n".join(corpus["names"][:4]) + "\n")
    kw = dict(batch=3)
    if mode == "nonuniform":
        kw["nonuniform"] = True
    if mode == "plane":
        kw["model_method"] = "plane"
    if mode == "deflate":
        kw["basic_compressor"] = "deflate"
    out = root + "/o_" + mode
    compress_datalist.compress(_args(small, out, **kw))
    for i in range(4):
        got = open(compress_datalist.output_path_for(out, corpus["names"][i]), "rb").read()
        want, o = _oracle_rpcc(corpus["frames"][i], nonuniform=(mode == "nonuniform"),
                               model_method="plane" if mode == "plane" else "point")
        if mode == "deflate":
            sec = parse_bitstream(got, uniform=True)
            for k, v in o["sections"].items():
                assert gzip.decompress(sec[k]) == v, (i, k)
        else:
            assert got == want, (mode, i)


def test_eval_columns_and_decompress_datalist(corpus, capsys):
    """--eval figures against a brute-force restatement on the CPU/GPU, and the decode tool's .bin files against the
    oracle's decode of the same streams (tools/decompress_datalist.py:93-131, dataset/dataset.py:72-81)."""
    import torch
    from rpcc_b200.evaluate_metrics import calc_chamfer_distance
    from rpcc_b200.tools import compress_datalist, decompress_datalist
    from rpcc_b200.tools.compress_datalist import M_CD_MEAN, M_DEPTH_MAX, M_DEPTH_MEAN, M_FSCORE, output_path_for
    root = corpus["root"]
    small = root + "/small6.txt"
    open(small, "w").write("\n".join(corpus["names"][:6]) + "\n")
    table = compress_datalist.compress(_args(small, root + "/oe", batch=4, eval=True, output=True))
    text = capsys.readouterr().out
    assert "Chamfer Distance (mean)" in text and "F1 score" in text and "BPP" in text
    rp = [output_path_for(root + "/oe", n) for n in corpus["names"][:6]]
    open(root + "/rp.txt", "w").write("\n".join(rp) + "\n")
    decompress_datalist.decompress(_args(root + "/rp.txt", root + "/dec", batch=4))
    H, W, hf, vmax, vmin = oracle.lidar_params(LIDAR)
    lut = oracle.transform_map(H, W, hf, vmax, vmin)
    for i in range(6):
        _, o = _oracle_rpcc(corpus["frames"][i])
        rec, xyz, seg = oracle.decompress_sections(o["sections"], LIDAR, 0.02)
        ri = o["range_image"]
        dif = np.abs(rec.reshape(H, W) - ri)
        assert table[i, M_DEPTH_MAX] == float(dif.max())
        assert abs(table[i, M_DEPTH_MEAN] - float(dif.astype(np.float64).mean())) <= 1e-9
        assert table[i, M_DEPTH_MAX] <= 0.02 + 1e-5
        # decode tool: exactly the rows save_point_cloud_to_file would write
        pc = xyz.reshape(-1, 3)
        pc = pc[np.where(np.sum(pc, -1) != 0)]
        got = np.fromfile(decompress_datalist.output_path_for(root + "/dec", rp[i]), np.float32).reshape(-1, 4)
        assert got.shape[0] == pc.shape[0] and np.array_equal(got[:, :3].view(np.uint32), pc.view(np.uint32)), i
        assert not got[:, 3].any()
        if i < 2:        # chamfer + F-score against the exact brute-force kernel (itself pinned to the reference's)
            ref = calc_chamfer_distance(ri[..., None] * lut, xyz, out=False)
            assert abs(table[i, M_CD_MEAN] - ref["mean"]) <= 1e-5 * ref["mean"]
            assert abs(table[i, M_FSCORE] - ref["f_score"]) <= 1e-6
    torch.cuda.synchronize()


def test_two_rank_torchrun_matches_single_rank(corpus):
    """Real torchrun, two ranks over NCCL (runs where the box has two GPUs): same files, and the gathered table."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rpcc_b200.tools import compress_datalist
    root = corpus["root"]
    compress_datalist.compress(_args(corpus["list"], root + "/s1", batch=8))
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "datalist_entry.py"), "compress", "--datalist",
           corpus["list"], "--output_dir", root + "/s2", "--lidar", LIDAR, "--workers", "2", "--batch", "8", "--eval"]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert _outputs(corpus, root + "/s2") == _outputs(corpus, root + "/s1")
    assert "Compressed %d frames on 2 GPU(s)" % N_FILES in res.stdout
