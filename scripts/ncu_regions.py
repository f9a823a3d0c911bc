"""Summarise an `ncu --page source --csv` dump: opcode mix and contiguous SASS regions by executed count.
usage: python scripts/ncu_regions.py dump.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ie, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) > ie and r[ie].isdigit():
        data.append((r[si].strip(), int(r[ie]), int(r[ss]) if r[ss].isdigit() else 0))
tot = sum(d[1] for d in data)
print("total warp-inst", tot, "sass", len(data))
ops = collections.Counter()
for s, n, _ in data:
    ops[s.split()[0].split(".")[0] if not s.startswith("@") else s.split()[1].split(".")[0]] += n
print(ops.most_common(22))
run, cur = [], None
for i, (s, n, sm) in enumerate(data):
    if cur is None or abs(n - cur[2]) > 0.3 * max(cur[2], 1) + 100:
        if cur:
            run.append(cur)
        cur = [i, i, n, n, sm]
    else:
        cur[1] = i
        cur[3] += n
        cur[4] += sm
run.append(cur)
for a, b, lv, t, sm in run:
    if t > tot * 0.01:
        print("sass %5d-%5d len %4d level %9d sum %11d %5.1f%% samples %d" % (a, b, b - a + 1, lv, t, 100 * t / tot, sm))
if len(sys.argv) > 2:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, b):
        print(i, data[i][1], data[i][2], data[i][0])
