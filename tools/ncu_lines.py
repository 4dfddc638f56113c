#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction samples with nvdisasm line info -> samples per CUDA source line.
usage: tools/ncu_lines.py <report.ncu-rep> <object.o|.so> <kernel-substring> [topN]"""
import csv, io, re, subprocess, sys, tempfile, os, glob
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
iA, iS, iSm, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(data[0][iA], 16)
samples = {int(r[iA], 16) - base: (int(r[iSm]), int(r[iE]), r[iS]) for r in data}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
txt = ""
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt += subprocess.run(["nvdisasm", "--print-line-info-inline", cub], capture_output=True, text=True).stdout
# find the kernel's .text section
m = re.search(r"\.section\s+\.text\.[^\n]*" + re.escape(kern) + r"[^\n]*\n(.*?)(?=\n\s*\.section|\Z)", txt, re.S)
body = m.group(1)
line = None; per_line = {}; 
for l in body.splitlines():
    lm = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if lm:
        line = (os.path.basename(lm.group(1)), int(lm.group(2))); continue
    am = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
    if am and line:
        off = int(am.group(1), 16)
        if off in samples:
            s, e, _ = samples[off]
            d = per_line.setdefault(line, [0, 0]); d[0] += s; d[1] += e
tot = sum(v[0] for v in per_line.values()) or 1
srcs = {}
def getsrc(f, n):
    if f not in srcs:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)) + glob.glob("**/" + f, recursive=True)
        srcs[f] = open(cands[0]).read().splitlines() if cands else []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
print(f"total samples {tot}")
for (f, n), (s, e) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*s/tot:5.1f}%  {s:6d} smp {e:9d} inst  {f}:{n}  {getsrc(f, n)}")
