#!/usr/bin/env python
"""Run one of the reference's OWN tools, unmodified, and time it.

  python baseline/run_reference_tool.py --ops reference|b200 [--tool compress_datalist] [--sync DIR --nprocs P]
         -- <the tool's own flags>

The tool (baseline/_ref/R-PCC/tools/<tool>.py, a byte-for-byte copy of the reference's file) is executed with runpy
as `__main__`, exactly as `python tools/compress_datalist.py ...` would, after everything it imports has been imported
once and the CUDA context exists -- so the time printed is the tool's work, not interpreter start-up.  `--ops` picks
what the reference's `ops.*` imports bind to: its own C++ / CUDA (`reference`) or librpcc_b200.so (`b200`, the
drop-in).  With `--sync DIR --nprocs P` the process first waits until P runners have checked in (files in DIR), so
that several processes sharing the host start together.  Prints one JSON line: {"seconds": wall time of the tool}."""
import argparse
import importlib
import importlib.util
import json
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref", "R-PCC")


def setup_paths(flavour):
    """sys.path for importing the staged reference with the chosen native binding; returns the reference root."""
    if not os.path.isdir(os.path.join(REF, "tools")):
        raise SystemExit("baseline/_ref/R-PCC is missing: run `python baseline/stage_reference.py` where /root/reference exists")
    sys.path[:0] = [os.path.join(HERE, "stubs"), os.path.join(HERE, "ops_" + flavour), REF, ROOT]
    if not hasattr(importlib, "find_loader"):          # removed in Python 3.12; dist_chamfer_3D.py:6 still calls it
        importlib.find_loader = lambda name: importlib.util.find_spec(name)
    if flavour == "b200":
        import rpcc_b200.plugin.chamfer_3D as chamfer
        sys.modules["chamfer_3D"] = chamfer            # dist_chamfer_3D.py:15-24 finds it instead of JIT-building
    return REF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="reference", choices=["reference", "b200"])
    ap.add_argument("--tool", default="compress_datalist")
    ap.add_argument("--sync", default=None)
    ap.add_argument("--nprocs", type=int, default=1)
    ap.add_argument("--tag", default="0")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    ref = setup_paths(a.ops)
    os.chdir(ref)
    import torch
    torch.zeros(1).cuda()
    # everything the tool imports, imported once (its top-level argparse runs under runpy below)
    import dataset  # noqa: F401
    import utils.compress_utils  # noqa: F401
    import utils.evaluate_metrics  # noqa: F401
    import utils.segment_utils  # noqa: F401
    if a.sync:
        open(os.path.join(a.sync, "ready.%s" % a.tag), "w").close()
        while len([f for f in os.listdir(a.sync) if f.startswith("ready.")]) < a.nprocs:
            time.sleep(0.005)
    tool = os.path.join(ref, "tools", a.tool + ".py")
    sys.argv = [tool] + rest
    out_fd = os.dup(1)
    os.dup2(2, 1)                 # the tool's own prints go to stderr; stdout carries the one JSON line
    t0 = time.time()
    runpy.run_path(tool, run_name="__main__")
    torch.cuda.synchronize()
    dt = time.time() - t0
    os.write(out_fd, (json.dumps({"seconds": dt, "t0": t0, "t1": t0 + dt, "ops": a.ops, "tool": a.tool}) + "\n").encode())


if __name__ == "__main__":
    main()
