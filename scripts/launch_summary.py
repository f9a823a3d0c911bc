"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
python scripts/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    unit, val = r[-2], float(r[-1].replace(",", ""))
    us = val / 1000.0 if unit in ("ns", "nsecond") else val * (1000.0 if unit in ("ms", "msecond") else 1.0)
    a = agg.setdefault(name, {"n": 0, "us": 0.0, "grid": r[8], "block": r[7]})
    a["n"] += 1
    a["us"] += us
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches | grid | block | total us | mean us | share |")
print("|---|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print("| %s | %d | %s | %s | %.1f | %.1f | %.1f%% |" % (k, a["n"], a["grid"], a["block"], a["us"], a["us"] / a["n"], 100 * a["us"] / tot))
print("\ntotal %.1f us over %d launches (per-launch times under ncu are serialised and cold-cache; compare SHARES with bench.py's roofline.kernels)" % (tot, len(rows)))
