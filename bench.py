#!/usr/bin/env python
"""bench.py -- frames/s of R-PCC's per-frame compression hot path on synthetic KITTI-like 64E frames.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

b200 arm.  A "step" is `--passes` passes of project -> ground fit -> FPS -> labels -> point models -> quantise + pack
over the `--frames` frames resident on each GPU (BASELINE.json configs[1]/[4]: Velodyne64E, FPS + point modelling,
accuracy 0.02, ~120k points per frame).
  value          device-resident: points already in HBM, CUDA events on the encoder's streams, max over ranks
  e2e            rpcc_encoder_encode_host: pinned host buffers in (xyz rows, 12 B/point -- the array the reference's
                 projection op receives), sections out, copies inside the timed region
  e2e.datalist   BASELINE configs[4]: rpcc_b200.tools.compress_datalist over `--datalist-frames` .bin files sharded over
                 the ranks -- file reads, GPU chain, host bzip2 and .rpcc writes -- frames/s of the whole tool
  e2e.decode     BatchDecoder on the .rpcc streams of the e2e frames (host bytes in, .bin rows out in pinned memory)
  roofline       per-stage CUDA-event times of the same steps on one stream slot; algorithmic bytes as DESIGN.md section 3

reference arm (`--impl reference`, rank 0 only).  The UNMODIFIED reference (baseline/_ref/R-PCC, staged by
baseline/stage_reference.py) driven through its own tools/compress_datalist.py on a bounded sample of the same
workload: default path (torch eager ops + its own FPS kernel on one GPU + its own C++), `--cpu`, one process and one
process per core.  The arm of round 1 (reference C++ + the oracle's C port of segment(), one process per core) is kept
and labelled as what it is.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "64E frames/sec (project->FPS->model->quantize)"
UNIT = "frames/s"
LIDAR = "Velodyne64E"
HW = 64 * 2000
WORKLOAD = "Velodyne64E uniform, FPS(100) + point modelling, accuracy 0.02 (BASELINE configs[1]/[4] frame shape)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=2368, help="frames resident on each GPU (default 2 x 1184)")
    ap.add_argument("--passes", type=int, default=10,
                    help="passes over the resident frames per step (10 x 2368 frames = ~120 ms per step: a 20-step run "
                         "is a 2.4 s timed region)")
    ap.add_argument("--distinct", type=int, default=64, help="distinct synthetic frames generated per GPU")
    ap.add_argument("--max-batch", type=int, default=1184,
                    help="frames per kernel launch (8 x 148 SMs: the one-CTA-per-frame kernels draw frames from a queue, "
                         "so a long launch evens out the frame-to-frame spread of the FPS work)")
    ap.add_argument("--e2e-frames", type=int, default=1184)
    ap.add_argument("--host-chunk", type=int, default=0,
                    help="frames per upload/kernels/download pipeline stage of encode_host (0 = three quarters of a frame per SM)")
    ap.add_argument("--datalist-frames", type=int, default=8192, help="files of the config-5 run, all ranks together (0 = skip)")
    ap.add_argument("--datalist-batch", type=int, default=592)
    ap.add_argument("--no-numa-bind", action="store_true", help="leave the process on whatever CPUs the launcher gave it")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--inject-ground", action="store_true",
                    help="feed the scene's true ground plane instead of fitting it on the device")
    ap.add_argument("--ref-seconds", type=float, default=10.0, help="reference arm: target wall time of each leg's sample")
    ap.add_argument("--ref-legs", default="all", help="reference arm: comma list of legs (tool,tool_cpu,multi,port) or all")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------- workload
def make_workload(distinct, frames, rank):
    """`distinct` generated frames tiled to `frames` slots (each slot is its own copy in memory, so
    every frame of a pass is read from HBM; the resident input is ~4 GB, far larger than L2).  Every rank generates
    the same frames (weak scaling with identical per-GPU work; `rank` only documents the call)."""
    from rpcc_b200 import synthetic
    per = [synthetic.frame(i, LIDAR) for i in range(distinct)]
    pts, grounds, off = [], [], [0]
    for i in range(frames):
        p, g = per[i % distinct]
        pts.append(p)
        grounds.append(g)
        off.append(off[-1] + p.shape[0])
    return np.concatenate(pts, 0), np.asarray(off, np.int64), np.stack(grounds).astype(np.float32)


def write_corpus(root, count, distinct=32, seed0=900000, first=0, last=None, make_sources=True):
    """`count` KITTI .bin paths under root/in (hard links to `distinct` generated files: every path is a file of its own
    to the tools, the page cache holds the bytes once -- a warm local disk).  Entries first..last are created by this
    caller (ranks share the work); returns all `count` paths."""
    from rpcc_b200 import synthetic
    os.makedirs(os.path.join(root, "in"), exist_ok=True)
    names = [os.path.join(root, "in", "%06d.bin" % i) for i in range(count)]
    last = count if last is None else last
    srcs = []
    for d in range(min(distinct, count)):
        src = os.path.join(root, "in", "src%03d.bin" % d)
        if make_sources and not os.path.exists(src):
            synthetic.frame(seed0 + d, LIDAR)[0].tofile(src + ".tmp")
            os.replace(src + ".tmp", src)
        srcs.append(src)
    for i in range(first, last):
        src = srcs[i % len(srcs)]
        while not os.path.exists(src):        # another rank is still writing the sources
            time.sleep(0.01)
        if not os.path.exists(names[i]):
            try:
                os.link(src, names[i])
            except OSError:
                shutil.copyfile(src, names[i])
    return names


def scratch_dir(tag):
    base = os.environ.get("RPCC_BENCH_TMP") or tempfile.gettempdir()
    # no "bin" / "rpcc" in the path: the tools replace the extension text everywhere in it, as the reference does
    return os.path.join(base, "pcc_%s_%d" % (tag, os.getuid()))


# ------------------------------------------------------------------------------------------- reference arm
def cpu_oracle_frame(args):
    """One frame through the round-1 CPU arm: the reference's compiled C++ (oracle/_ref) for the stages it has on the
    CPU, the oracle's C restatement for segment() (GPU-only in the reference)."""
    import oracle
    from oracle import ref
    pts, ground, use_ref = args
    H, W, hf, vmax, vmin = oracle.lidar_params(LIDAR)
    lut = cpu_oracle_frame.lut
    t = [time.perf_counter()]
    if use_ref:
        d, s, q, c = (ref.cpp("dataset_utils_cpp"), ref.cpp("segment_utils_cpp"), ref.cpp("quantization_utils_cpp"),
                      ref.cpp("contour_utils_cpp"))
        ri = d.point_cloud_to_range_image_even(np.ascontiguousarray(pts[:, :3]), H, W, hf, vmax, vmin)
        t.append(time.perf_counter())
        seg, _, _ = oracle.segment(ri, lut, ground, 100)
        t.append(time.perf_counter())
        pm = s.point_modeling(ri[..., None], seg)
        cm = np.concatenate((np.zeros((pm.shape[0], 3)), pm[:, None]), -1)[1:]
        mp = np.concatenate((np.asarray(ground, np.float64).reshape(1, 4), cm), 0)
        pred = s.intra_predict(seg, mp, lut)
        res = ri[..., None] - pred
        sym = q.uniform_quantize(seg, res, 0.04)
        c.extract_contour(seg)
        t.append(time.perf_counter())
        return int(sym.size), t[1] - t[0], t[2] - t[1], t[3] - t[2]
    out = oracle.compress_frame(pts, LIDAR, ground)
    return int(out["symbols"].size), 0.0, time.perf_counter() - t[0], 0.0


def _cpu_init():
    import oracle
    H, W, hf, vmax, vmin = oracle.lidar_params(LIDAR)
    cpu_oracle_frame.lut = oracle.transform_map(H, W, hf, vmax, vmin)


def cpu_throughput(n_frames, cores, use_ref):
    """frames/s of the round-1 CPU arm on `cores` processes over a bounded sample of the workload, and its stage split."""
    import multiprocessing as mp
    from rpcc_b200 import synthetic
    frames = [synthetic.frame(900000 + i, LIDAR) for i in range(min(n_frames, 32))]
    jobs = [(frames[i % len(frames)][0], frames[i % len(frames)][1], use_ref) for i in range(n_frames)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(cpu_oracle_frame, jobs[:cores])  # warm the workers
        t0 = time.perf_counter()
        rows = pool.map(cpu_oracle_frame, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    split = {"project_ms": 1e3 * float(np.mean([r[1] for r in rows])), "segment_port_ms": 1e3 * float(np.mean([r[2] for r in rows])),
             "model_predict_quantize_contour_ms": 1e3 * float(np.mean([r[3] for r in rows]))}
    return n_frames / dt, dt, split


def run_reference_tool(names, workers, extra, nprocs=1, timeout=900):
    """The unmodified reference tools/compress_datalist.py over `names`, as `nprocs` processes (each on a contiguous
    shard, started together).  -> (frames/s over the span first start .. last end, seconds, stderr tail)"""
    root = os.path.dirname(os.path.dirname(names[0]))
    sync = tempfile.mkdtemp(prefix="sync", dir=root)
    per = (len(names) + nprocs - 1) // nprocs
    shards = [names[p * per:(p + 1) * per] for p in range(nprocs)]
    shards = [s for s in shards if s]
    procs = []
    for p, mine in enumerate(shards):
        lst = os.path.join(sync, "list%d.txt" % p)
        open(lst, "w").write("\n".join(mine) + "\n")
        procs.append([sys.executable, os.path.join(ROOT, "baseline", "run_reference_tool.py"), "--ops", "reference", "--sync", sync,
                      "--nprocs", str(len(shards)), "--tag", str(p), "--", "--datalist", lst, "--output_dir",
                      os.path.join(root, "ref_out"), "--lidar", LIDAR, "--workers", str(workers)] + list(extra))
    running = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for c in procs]
    outs = []
    for r in running:
        try:
            o, e = r.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            r.kill()
            o, e = r.communicate()
        outs.append((r.returncode, o, e))
    shutil.rmtree(sync, ignore_errors=True)
    shutil.rmtree(os.path.join(root, "ref_out"), ignore_errors=True)
    bad = [(rc, e[-600:]) for rc, o, e in outs if rc != 0]
    if bad:
        return None, None, "exit %s: %s" % (bad[0][0], bad[0][1])
    spans = [json.loads(o.strip().splitlines()[-1]) for _, o, _ in outs]
    dt = max(s["t1"] for s in spans) - min(s["t0"] for s in spans)
    return len(names) / dt, dt, ""


def stage_split_of_reference_tool(names, extra):
    """The reference's own per-stage time table (tools/compress_datalist.py:145-158, printed under --output), one
    worker, averaged over the frames."""
    root = os.path.dirname(os.path.dirname(names[0]))
    lst = os.path.join(root, "split.txt")
    open(lst, "w").write("\n".join(names) + "\n")
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_reference_tool.py"), "--ops", "reference", "--", "--datalist", lst,
           "--output_dir", os.path.join(root, "ref_out"), "--lidar", LIDAR, "--workers", "1", "--output"] + list(extra)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    shutil.rmtree(os.path.join(root, "ref_out"), ignore_errors=True)
    if r.returncode != 0:
        return {"error": r.stderr[-400:]}
    keys = {"Load data": "load_ms", "Segmentation module": "segment_ms", "Modeling module": "model_ms",
            "Intra-prediction module": "predict_ms", "Quantization module": "quantize_ms", "Basic compressor module": "entropy_ms",
            "Save binary file": "save_ms", "Total time cost": "total_ms"}
    acc = {}
    for line in r.stderr.splitlines():
        s = line.strip()
        for k, name in keys.items():
            if s.startswith(k + ":") or s.startswith(k + " ("):
                try:
                    acc.setdefault(name, []).append(float(s.split(":")[-1]))
                except ValueError:
                    pass
    return {k: 1e3 * float(np.mean(v[1:] if len(v) > 2 else v)) for k, v in acc.items()}


def reference_arm(a):
    """--impl reference: the reference's own implementation of the path, timed on this box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from oracle import ref
    oracle.lib()
    cores = os.cpu_count() or 1
    legs_wanted = ("tool", "tool_cpu", "multi", "port") if a.ref_legs == "all" else tuple(a.ref_legs.split(","))
    staged = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "R-PCC", "tools")) and ref.have_cpp() and ref.have_cuda()
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    legs, notes = {}, []
    root = scratch_dir("ref")
    shutil.rmtree(root, ignore_errors=True)
    t_start = time.time()
    if staged and have_gpu:
        names_all = write_corpus(root, 2048)

        def leg(tag, workers, extra, nprocs, what, est=None):
            """Calibrate on a short run (the W warm-up steps) unless an estimate is at hand, then time one sample sized
            for ~ref_seconds."""
            n0 = max(8, 2 * nprocs)
            fps0 = est
            if fps0 is None:
                fps0, dt0, err = run_reference_tool(names_all[:n0], workers, extra, nprocs)
                if fps0 is None:
                    notes.append("%s failed: %s" % (tag, err))
                    return
            n = int(min(len(names_all), max(n0, fps0 * a.ref_seconds)))
            fps, dt, err = run_reference_tool(names_all[:n], workers, extra, nprocs)
            if fps is not None and dt < 0.5 * a.ref_seconds and n < len(names_all):
                # the short calibration run is dominated by first-call costs and undersizes the sample: once more
                n = int(min(len(names_all), max(n, fps * a.ref_seconds)))
                fps, dt, err = run_reference_tool(names_all[:n], workers, extra, nprocs)
            if fps is None:
                notes.append("%s failed: %s" % (tag, err))
                return
            legs[tag] = {"value": fps, "unit": UNIT, "frames": n, "seconds": dt, "processes": nprocs, "workers_per_process": workers,
                         "what": what}

        nop = ["--basic_compressor", "lz4"]       # baseline/stubs/lz4.py is a pass-through: no entropy coder
        if "tool" in legs_wanted:
            leg("tool_gpu_path", cores, nop, 1, "unmodified tools/compress_datalist.py --workers %d, default path (torch eager "
                "ops + the reference's FPS kernel on one GPU + its C++), entropy coder off (lz4 pass-through stub)" % cores)
            leg("tool_gpu_path_bzip2", cores, [], 1, "the same with bzip2: BASELINE configs[4]'s comparator, file I/O and coder included",
                est=0.7 * legs["tool_gpu_path"]["value"] if "tool_gpu_path" in legs else None)
            legs["tool_gpu_path_stage_split_ms"] = stage_split_of_reference_tool(names_all[:12], [])
        if "tool_cpu" in legs_wanted:
            leg("tool_cpu_flag_bzip2", cores, ["--cpu"], 1, "unmodified tool with --cpu (numpy label assignment; FPS still on the "
                "GPU, utils/segment_utils.py:121), bzip2")
        if "multi" in legs_wanted:
            P = max(1, min(cores, 16))
            leg("multi_process_gpu_path", 2, nop, P, "%d concurrent processes of the unmodified tool (--workers 2 each) sharing "
                "one GPU, entropy coder off: every host core the reference can use" % P)
            leg("multi_process_gpu_path_bzip2", 2, [], P, "the same with bzip2",
                est=0.7 * legs["multi_process_gpu_path"]["value"] if "multi_process_gpu_path" in legs else None)
    else:
        notes.append("reference tool legs skipped: %s" % ("baseline/_ref or oracle/_ref missing" if not staged else "no CUDA device"))
    port = None
    if "port" in legs_wanted:
        use_ref = ref.have_cpp()
        n = max(cores * 4, 32)
        cpu_throughput(cores, cores, use_ref)
        fps, dt, split = cpu_throughput(n, cores, use_ref)
        port = {"value": fps, "unit": UNIT, "frames": n, "seconds": dt, "processes": cores, "stage_split_ms": split,
                "what": "round-1 arm: %s for projection / point_modeling / intra_predict / uniform_quantize / extract_contour + the "
                        "ORACLE'S C PORT of segment() (mask + FPS + label assignment; not reference code -- the reference runs "
                        "that stage on the GPU), one process per core, ground injected, no entropy coder" %
                        ("reference C++ (oracle/_ref)" if use_ref else "oracle C restatement")}
        legs["port_cpp_plus_segment_port"] = port
    shutil.rmtree(root, ignore_errors=True)
    ref_only = {k: v for k, v in legs.items() if isinstance(v, dict) and "value" in v and k != "port_cpp_plus_segment_port"}
    hot = {k: v for k, v in ref_only.items() if not k.endswith("bzip2")}
    coded = {k: v for k, v in ref_only.items() if k.endswith("bzip2")}
    if hot:
        best = max(hot, key=lambda k: hot[k]["value"])
        value, kind, sample = hot[best]["value"], "reference", "%s: %s; %d frames in %.1f s" % (best, hot[best]["what"], hot[best]["frames"], hot[best]["seconds"])
        frames_used = hot[best]["frames"]
    elif port:
        value, kind = port["value"], "reference C++ + port(segment)" if ref.have_cpp() else "port"
        sample, frames_used = port["what"], port["frames"]
    else:
        emit({"impl": "reference", "unavailable": "; ".join(notes) or "no leg selected"})
        return
    datalist = None
    if coded:
        bestc = max(coded, key=lambda k: coded[k]["value"])
        datalist = {"value": coded[bestc]["value"], "unit": UNIT, "leg": bestc, "frames": coded[bestc]["frames"],
                    "what": coded[bestc]["what"]}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * frames_used / value / max(a.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": frames_used / max(a.steps, 1),
                       "sample": "each leg: a calibration run (the warm-up), then one timed run of the reference's own tool over "
                                 "enough frames for ~%.0f s; value = the fastest leg made of reference code only" % a.ref_seconds},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "datalist": datalist, "legs": legs, "notes": notes, "gpu_launches": 0, "arm_seconds": time.time() - t_start}
    emit(line)


# ------------------------------------------------------------------------------------------- main arm
_JSON_FD = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (NCCL banners, warnings)
    was diverted to stderr at start-up."""
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _JSON_FD
    a = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if a.impl == "reference":
        reference_arm(a)
        return
    import torch
    import torch.distributed as dist

    import rpcc_b200
    from rpcc_b200.batch import BatchDecoder, BatchEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    # one process per GPU: stay on the CPUs (and so the memory) next to it; undone before the cpu_baseline leg, which is
    # entitled to every host core
    from rpcc_b200.shard import bind_to_gpu_numa
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if a.no_numa_bind else bind_to_gpu_numa(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(o.item()) for o in out]
        return [float(x)]

    F, MB = a.frames, min(a.max_batch, a.frames)
    pts_np, off_np, g_np = make_workload(a.distinct, F, rank)
    npts = int(off_np[-1])
    max_chunk_pts = int(max(off_np[min(i + MB, F)] - off_np[i] for i in range(0, F, MB)))
    enc = BatchEncoder(LIDAR, accuracy=0.02, max_batch=MB, max_points=max_chunk_pts, device=local,
                       host_chunk=a.host_chunk)
    d_pts = torch.from_numpy(pts_np).cuda()
    d_off = torch.from_numpy(off_np).cuda()
    d_g = torch.from_numpy(g_np).cuda()
    chunks = [(i, min(MB, F - i)) for i in range(0, F, MB)]
    nslots = enc.slots
    streams = [torch.cuda.ExternalStream(enc.stream(s)) for s in range(nslots)]

    def step():
        ci = 0
        for _ in range(a.passes):
            for (f0, nb) in chunks:
                # un-rebased offsets: the kernel indexes `points` with absolute row numbers
                enc.encode_device(ci % nslots, d_pts, d_off[f0:f0 + nb + 1], nb, d_g[f0:f0 + nb] if a.inject_ground else None)
                ci += 1

    for _ in range(max(a.warmup, 3)):
        step()
    enc.sync()
    torch.cuda.synchronize()
    barrier()

    launches0 = rpcc_b200.launch_count()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(nslots)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(nslots)]
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        barrier()
        for s in range(nslots):
            starts[s].record(streams[s])
        for _ in range(a.steps):
            step()
        for s in range(nslots):
            ends[s].record(streams[s])
        enc.sync()
        torch.cuda.synchronize()
        barrier()
    my_ms = max(starts[i].elapsed_time(ends[j]) for i in range(nslots) for j in range(nslots))
    launches = rpcc_b200.launch_count() - launches0
    # per-kernel durations: the same frames again on ONE stream slot (no overlap between chunks), with
    # CUDA events recorded on that stream between the stages of every launch
    enc.profile(True)
    for _ in range(max(2, min(a.steps, 8))):
        for (f0, nb) in chunks:
            enc.encode_device(0, d_pts, d_off[f0:f0 + nb + 1], nb, d_g[f0:f0 + nb] if a.inject_ground else None)
    stage_ms, stage_frames, stage_calls = enc.stage_times()
    enc.profile(False)
    res0 = enc.device_buffer(0, "results", (MB, 4), torch.int32).cpu().numpy()      # the last launch of slot 0
    valid_mean = float(res0[:, 0].astype(np.int64).mean())
    runs_mean = float(res0[:, 1].astype(np.int64).mean())
    per_rank_ms = all_ranks(my_ms)
    per_rank_stage = {k: all_ranks(v / max(stage_calls, 1)) for k, v in stage_ms.items()}
    elapsed_ms = max(per_rank_ms)
    frames_per_step = F * a.passes
    value = world * frames_per_step * a.steps / (elapsed_ms / 1000.0)

    # ---- e2e: host buffers through the C ABI (upload + kernels + download inside the timed region)
    e2e = None
    blobs = None
    if not a.no_e2e:
        EF = min(a.e2e_frames, F)
        h_xyz = torch.from_numpy(np.ascontiguousarray(pts_np[:off_np[EF], :3])).pin_memory()
        h_off = off_np[:EF + 1].copy()
        h_g = g_np[:EF].copy() if a.inject_ground else None
        out = enc.encode_host(h_xyz, h_off, h_g)  # warm-up (allocates the pinned output buffers)
        d2h = int(out["symbols"].nbytes + out["seq"].nbytes + out["model"].nbytes + out["contour"].nbytes + 16 * EF)
        h2d = int(h_xyz.numel() * 4 + h_off.nbytes + (h_g.nbytes if h_g is not None else 0))
        ref_sym = out["symbols"].copy()
        enc.encode_host(h_xyz, h_off, h_g)
        torch.cuda.synchronize()
        # the PCIe link on its own: the same pinned buffer uploaded by plain copies (every rank at the same moment:
        # under torchrun this is the contended figure, the ceiling of this rank's e2e)
        d_tmp = torch.empty((h_xyz.shape[0], 3), dtype=torch.float32, device="cuda")
        d_tmp.copy_(h_xyz, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(3):
            d_tmp.copy_(h_xyz, non_blocking=True)
        l1.record()
        torch.cuda.synchronize()
        link_gbs = 3 * h_xyz.numel() * 4 / (l0.elapsed_time(l1) / 1000.0) / 1e9
        del d_tmp
        barrier()

        def timed_host(buf):
            reps = max(2, a.steps)
            torch.cuda.synchronize()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            t0 = time.perf_counter()
            for _ in range(reps):
                enc.encode_host(buf, h_off, h_g)
            e1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            return world * EF * reps / max_over_ranks(max(e0.elapsed_time(e1) / 1000.0, wall)), reps

        e2e_fps, reps = timed_host(h_xyz)
        # the same call fed the raw KITTI rows (x, y, z, intensity): 16 B/point over the link
        h_pts = torch.from_numpy(pts_np[:off_np[EF]]).pin_memory()
        same = bool(np.array_equal(ref_sym, enc.encode_host(h_pts, h_off, h_g)["symbols"]))
        kitti_fps, _ = timed_host(h_pts)
        kitti_rows = {"value": kitti_fps, "unit": UNIT, "h2d_bytes_per_step": int(h_pts.numel() * 4 + h_off.nbytes),
                      "symbols_identical_to_the_xyz_rows": same}
        del h_pts
        link_ceiling = link_gbs * 1e9 / (h2d / EF)     # frames/s this rank's link could carry if nothing else cost time
        e2e = {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "frames_per_step": EF, "rows": "xyz (12 B/point): what the reference's projection op receives after "
                                              "dataset/dataset.py:62 dropped the intensity column",
               "h2d_link_gbs_plain_copy_per_gpu": all_ranks(link_gbs), "h2d_achieved_gbs_per_gpu": h2d * e2e_fps / (world * EF) / 1e9,
               "link_ceiling_frames_per_s_all_gpus": float(sum(all_ranks(link_ceiling))),
               "note": "bound by the upload of the points over PCIe: link_ceiling = what the ranks' links, probed at the same "
                       "moment with plain copies, could carry",
               "kitti_rows_16B": kitti_rows,
               "cpus_per_rank_after_numa_binding": all_ranks(float(len(numa_cpus)) if numa_cpus else 0.0)}

        # ---- e2e.decode: .rpcc streams (host bytes) -> rows of the output .bin files in pinned host memory
        ND = min(EF, 592)
        blobs = enc.compress(h_xyz[:off_np[ND]], off_np[:ND + 1].copy(), g_np[:ND].copy() if a.inject_ground else None)
        dec = BatchDecoder(LIDAR, accuracy=0.02, device=local)
        dec.decode(blobs, want_xyz=False, want_points=True)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        dreps = 3
        for _ in range(dreps):
            dout = dec.decode(blobs, want_xyz=False, want_points=True)
        torch.cuda.synchronize()
        t_dec = max_over_ranks(time.perf_counter() - t0)
        in_bytes = int(sum(len(b) for b in blobs))
        out_bytes = int(sum(p.nbytes for p in dout["points"]))
        e2e["decode"] = {"value": world * ND * dreps / t_dec, "unit": UNIT, "frames_per_step": ND, "host_threads": dec.workers,
                         "h2d_bytes_per_step": None, "rpcc_bytes_in_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                         "note": "host: bzip2 decoding of the sections on threads (the bound); device: recover_map / "
                                 "dequantise / predict / range x LUT / row compaction; output = the .bin rows in pinned memory"}
        dec.close()
        e2e["mean_rpcc_bytes"] = float(np.mean([len(b) for b in blobs]))

        # ---- e2e.datalist: BASELINE configs[4], the tool end to end
        if a.datalist_frames > 0:
            from rpcc_b200.shard import shard_range
            from rpcc_b200.tools import compress_datalist
            from rpcc_b200.tools.common import base_parser
            root = scratch_dir("dl")
            if local == 0:
                shutil.rmtree(root, ignore_errors=True)
            barrier()
            lo, hi = shard_range(a.datalist_frames, rank, world)
            names = write_corpus(root, a.datalist_frames, distinct=32, first=lo, last=hi, make_sources=rank == 0)
            if rank == 0:
                open(os.path.join(root, "list.txt"), "w").write("\n".join(names) + "\n")
            barrier()
            workers = enc.workers
            targs = base_parser(single=False).parse_args(["--datalist", os.path.join(root, "list.txt"), "--output_dir",
                                                          os.path.join(root, "out"), "--lidar", LIDAR, "--workers", str(workers),
                                                          "--batch", str(a.datalist_batch)])
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            table = compress_datalist.compress(targs, rank=rank, world=world, collective=world > 1)
            torch.cuda.synchronize()
            barrier()
            t_dl = max_over_ranks(time.perf_counter() - t0)
            e2e["datalist"] = {"value": a.datalist_frames / t_dl, "unit": UNIT, "frames": a.datalist_frames, "seconds": t_dl,
                               "host_threads_per_rank": workers, "host_cores": os.cpu_count(), "frames_per_launch": a.datalist_batch,
                               "mean_rpcc_bytes": float(table[:, 0].mean()) if len(table) else None,
                               "what": "rpcc_b200.tools.compress_datalist over %d KITTI .bin files (hard links to 32 distinct "
                                       "frames, page-cache warm) sharded over %d rank(s): file reads -> GPU chain -> bzip2 on "
                                       "the native pool -> .rpcc files; bound by libbz2 on the host cores" % (a.datalist_frames, world)}
            # full-size consistency (outside the timed region): the corpus is `distinct` frames under many names, and a
            # file's bytes may depend on nothing but its points -- every copy of a frame, on whichever rank and in whichever
            # batch it landed, must have produced the same .rpcc
            import hashlib
            import json as _json
            digests = {}
            for i in range(lo, hi):
                with open(compress_datalist.output_path_for(os.path.join(root, "out"), names[i]), "rb") as fh:
                    digests.setdefault(i % 32, set()).add(hashlib.sha256(fh.read()).hexdigest())
            with open(os.path.join(root, "digests_%d.json" % rank), "w") as fh:
                _json.dump({str(k): sorted(v) for k, v in digests.items()}, fh)
            barrier()
            if rank == 0:
                merged = {}
                for r in range(world):
                    for k, v in _json.load(open(os.path.join(root, "digests_%d.json" % r))).items():
                        merged.setdefault(k, set()).update(v)
                bad = sorted(k for k, v in merged.items() if len(v) != 1)
                e2e["datalist"]["consistency"] = {"files_hashed": a.datalist_frames, "distinct_frames": len(merged),
                                                  "distinct_outputs": sum(len(v) for v in merged.values()),
                                                  "ok": not bad}
                if bad:
                    raise SystemExit("bench: copies of frames %s produced different .rpcc bytes" % bad[:8])
            barrier()
            if local == 0:
                shutil.rmtree(root, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline: algorithmic bytes (SURVEY 8d / DESIGN.md) over the live per-stage event times
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_mean = npts / F
    alg = {  # bytes per frame
        "project": 16.0 * n_mean + 4.0 * HW,
        "quantize": 4.0 * HW + 1.0 * HW + 2.0 * valid_mean + HW / 8.0 + 2.0 * runs_mean,
        "assign": 4.0 * HW + 1.0 * HW,
        "fps": 4.0 * HW,
    }
    total_stage = sum(stage_ms.values()) or 1.0
    kernels = {}
    for k, ms in stage_ms.items():
        if ms <= 0:
            continue
        ent = {"ms_per_launch": ms / max(stage_calls, 1), "share_of_step": ms / total_stage,
               "frames_per_launch": stage_frames / max(stage_calls, 1)}
        if k in alg:
            gbs = alg[k] * stage_frames / (ms / 1000.0) / 1e9
            ent.update({"algorithmic_bytes_per_frame": alg[k], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak})
        if world > 1:
            ent["ms_per_launch_per_rank"] = per_rank_stage.get(k)
        kernels[k] = ent
    pj = kernels.get("project", {})
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))["project_kernel"]
            if int(tj["frames_per_launch"]) == int(round(pj.get("frames_per_launch", 0))):
                traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("capture")
        except Exception:
            pass
    roofline = {"kernel": "project_kernel", "bound": "hbm", "achieved": pj.get("achieved_gbs"), "peak": peak,
                "unit": "GB/s", "frac": pj.get("frac_of_hbm_peak"), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": pj.get("algorithmic_bytes_per_frame", 0.0) * pj.get("frames_per_launch", 0.0),
                "peak_source": peak_src,
                "note": "dominant HBM mover of the chain (16 B/point in + 4 B/pixel out); the chain's time is dominated by "
                        "the latency/issue-bound FPS and label kernels, see kernels{}", "kernels": kernels}

    cpu_baseline = None
    os.sched_setaffinity(0, all_cpus)
    if not a.no_cpu_baseline:
        import oracle  # noqa: F401  (checker; the one CPU leg of this arm)
        from oracle import ref
        use_ref = ref.have_cpp()
        cores = os.cpu_count() or 1
        n = a.cpu_frames or max(10 * cores, 64)        # ~0.1 s of CPU per frame: 10-30 s of CPU work, ~1 s of wall clock
        fps, dt, split = cpu_throughput(n, cores, use_ref)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference C++ + port(segment)" if use_ref else "port",
                        "sample": "%d frames of the same synthetic 64E workload in %.1f s on %d host processes (%s for the CPU "
                                  "stages; the oracle's C port of the GPU-only segment(), %.0f of %.0f ms per frame; no entropy "
                                  "coder).  The reference's own tool is timed by `--impl reference`." %
                                  (n, dt, cores, "reference C++ from oracle/_ref" if use_ref else "oracle C restatement",
                                   split["segment_port_ms"], sum(split.values())),
                        "stage_split_ms": split}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + ", %.0f points/frame" % n_mean,
                       "frames_per_step_per_gpu": frames_per_step, "frames_resident_per_gpu": F, "passes_per_step": a.passes,
                       "frames_per_launch": MB, "distinct_frames": a.distinct,
                       "l2": "inputs larger than L2 (%.2f GB of points resident, every pass reads all of it)" % (npts * 16 / 1e9),
                       "ground_model": "injected (true plane of the synthetic scene)" if a.inject_ground else
                                       "fitted on the device (deterministic RANSAC, inside the timed region)",
                       "stage_timing": "roofline.kernels: the same frames re-run on one stream slot with CUDA events between stages",
                       "per_rank": "every rank runs the same frames; value uses the slowest rank's elapsed time"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "per_rank_elapsed_ms": per_rank_ms}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
