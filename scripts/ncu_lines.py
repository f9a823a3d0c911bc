#!/usr/bin/env python
"""Per-CUDA-source-line totals of an ncu report captured with --import-source on:
   python scripts/ncu_lines.py file.ncu-rep [top]
Prints warp instructions executed and stall samples per (file, line), sorted by samples."""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, lines = None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0].isdigit() and len(r) > 7 and r[7].replace(".", "").isdigit():
        lines.append((fname, int(r[0]), r[1].strip(), int(r[6]) if r[6].isdigit() else 0, int(r[7])))
tot_i = sum(x[4] for x in lines) or 1
tot_s = sum(x[3] for x in lines) or 1
print("total warp-inst %d, samples %d" % (tot_i, tot_s))
for f, ln, src, smp, ins in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * smp / tot_s, 100.0 * ins / tot_i, f, ln, src[:110]))
