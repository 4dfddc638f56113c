#!/usr/bin/env python
"""Stall-reason breakdown of an ncu source-page CSV (ncu -i rep --page source --csv > f.csv):
   tools/ncu_stalls.py f.csv [addr_lo addr_hi]   (hex offsets relative to the kernel start; default whole kernel)
prints totals per stall reason and the top instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
iA, iS, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
idx = {r: hdr.index(r) for r in reasons}
base = int(data[0][iA], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 40
tot = {r: 0 for r in reasons}; insts = []
for r in data:
    off = int(r[iA], 16) - base
    if not (lo <= off < hi): continue
    n = int(r[iN] or 0)
    st = {k: int(r[i] or 0) for k, i in idx.items()}
    for k, v in st.items(): tot[k] += v
    insts.append((n, off, r[iS], st))
alls = sum(n for n, *_ in insts) or 1
print(f"samples {alls} in [{lo:#x},{hi:#x})")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v: print(f"  {k:24s} {v:6d} {100*v/alls:5.1f}%")
for n, off, src, st in sorted(insts, key=lambda t: -t[0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {off:#7x} {n:5d} {100*n/alls:4.1f}%  {src[:70]:70s} {top}")
