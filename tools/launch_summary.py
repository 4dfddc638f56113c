#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): the launches of the LAST forward, grouped by
kernel and grid.   tools/launch_summary.py file.csv [launches_per_step] [top]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))][1:]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if per == 0:
    # the forward starts with the alpha pyramid: take everything from its second-to-last occurrence pair on
    idx = [i for i, r in enumerate(rows) if "alpha_pyramid" in r[4]]
    per = len(rows) - idx[-3] if len(idx) >= 3 else len(rows)
last = rows[-per:]
agg = collections.OrderedDict()
for r in last:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("b200::<unnamed>::", "")[:58]
    a = agg.setdefault((name, r[8]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[-1]) / 1e3
tot = sum(a[1] for a in agg.values())
print(f"{per} launches, {tot / 1e3:.3f} ms of kernel time")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print(f"{a[1]:9.1f} us {100 * a[1] / tot:5.1f}% n={a[0]:3d} avg={a[1] / a[0]:8.1f}  {k[0]} {k[1]}")
