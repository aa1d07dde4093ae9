#!/usr/bin/env python3
"""tools/ncu_launch_sum.py -- add up an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv ...`):
kernel-only time of the command, per kernel name and in total.    python tools/ncu_launch_sum.py x.csv [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = {}
for r in rows:
    if r is hdr or len(r) <= iv or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(r[iu], 1e-6)
    name = r[ik].split("(")[0][-80:]
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
total = sum(v for _, v in acc.values())
print("%s: %d launches, %.3f ms of kernel time" % (path, sum(c for c, _ in acc.values()), total))
for name, (c, v) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print("   %8.3f ms  %5.1f %%  %4d x  %s" % (v, 100 * v / total, c, name))
