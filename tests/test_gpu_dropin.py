"""The drop-in claim of INTEGRATION.md, run: the reference's UNMODIFIED Python (staged byte for byte under the git-ignored
baseline/_ref/R-PCC) on librpcc_b200.so through the `ops` shims of baseline/ops_b200, against the goldens that the same
Python produced on the reference's own C++ (tests/golden/make_golden.py).  BASELINE configs[0] (tools/compress.py on
example.bin) and configs[1] (non-uniform encode -> .rpcc -> decode -> chamfer) each in one piece."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import EXAMPLE_GROUND

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "example_golden.npz")
EXAMPLE = os.path.join(ROOT, "tests", "golden", "example.bin")
STAGED = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "R-PCC", "tools"))
needs_ref = pytest.mark.skipif(not STAGED, reason="baseline/_ref/R-PCC not staged (python baseline/stage_reference.py where /root/reference exists)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _env():
    return dict(os.environ, RPCC_STUB_GROUND=",".join(repr(x) for x in EXAMPLE_GROUND), PYTHONPATH=ROOT)


@needs_ref
@pytest.mark.parametrize("mode", ["uniform", "nonuniform"])
def test_reference_l3_python_unmodified_on_the_shims(mode):
    gold = np.load(GOLD)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "dropin_check.py"), EXAMPLE, mode, "b200"],
                       env=_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert "baseline/_ref/R-PCC/utils/segment_utils.py" in got["segment_utils"]      # the reference's file ran, not the mirror
    t = "u_" if mode == "uniform" else "n_"
    assert got["range_sha"] == str(gold[t + "range_sha"])
    assert got["seg_sha"] == sha(gold[t + "seg_u8"])                    # the reference's torch block + the FPS plugin
    assert got["model_sha"] == sha(gold[t + "model_param"])
    assert got["pred_sha"] == str(gold[t + "pred_sha"])
    assert got["symbols_sha"] == sha(gold[t + "symbols"])
    assert got["rpcc_sha"] == sha(gold[t + "rpcc"]) and got["rpcc_bytes"] == gold[t + "rpcc"].size
    assert got["rec_sha"] == str(gold[t + "rec_sha"]) and got["xyz_sha"] == str(gold[t + "xyz_sha"])
    assert got["max_err"] == float(gold[t + "max_err"])
    if mode == "nonuniform":
        assert got["salience_sha"] == sha(gold["n_salience"]) and got["key_points_sha"] == sha(gold["n_key_points_u8"])
    # chamfer of the round trip through the reference's calc_chamfer_distance on the chamfer_3D shim: against the oracle
    import oracle
    H, W, hf, vmax, vmin = oracle.lidar_params("Velodyne64E")
    pts = np.fromfile(EXAMPLE, np.float32).reshape(-1, 4)
    o = oracle.compress_frame(pts, "Velodyne64E", np.array(EXAMPLE_GROUND), nonuniform=(mode == "nonuniform"))
    rec, xyz, _ = oracle.decompress_sections(o["sections"], "Velodyne64E", 0.02)
    from rpcc_b200.evaluate_metrics import calc_chamfer_distance        # exact kernel, pinned to the reference's chamfer3D.cu
    want = calc_chamfer_distance(o["range_image"][..., None] * o["lut"], xyz, out=False)
    assert abs(got["chamfer_mean"] - want["mean"]) <= 1e-5 * want["mean"] and abs(got["f_score"] - want["f_score"]) <= 1e-6
    assert got["dist1_sha"] == sha(want["chamfer_dist_info"]["dist1"])


@needs_ref
def test_reference_compress_tool_unmodified_on_the_shims(tmp_path):
    """`python tools/compress.py --input example.bin ...` -- the reference's own file, executed as __main__ -- writes the
    golden .rpcc when its `ops.*` imports resolve to librpcc_b200.so (BASELINE configs[0])."""
    gold = np.load(GOLD)
    for mode, key in ((None, "u_rpcc"), ("--nonuniform", "n_rpcc")):
        out = str(tmp_path / ("x%s.rpcc" % (mode or "")))
        cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_reference_tool.py"), "--ops", "b200", "--tool", "compress", "--",
               "--input", EXAMPLE, "--output", out, "--lidar", "Velodyne64E"] + ([mode] if mode else [])
        r = subprocess.run(cmd, env=_env(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        assert open(out, "rb").read() == gold[key].tobytes(), mode
        assert "BPP" in r.stderr
