#!/usr/bin/env python
"""DRAM traffic of the hot-path kernels from `ncu --set full` raw pages -> profiles/r02_ncu_traffic.json, the table that
bench.py reports as `roofline.traffic` (per launch, like `achieved`).  Every entry carries the sha256 of the CSV it was
read from and the digest of the library that was profiled, so a stale table is detectable.

    python tools/ncu_traffic.py attention_8x8=profiles/r02_attn_sp_ncu_raw.csv:mwa_sp_kernel \
                                gdn=profiles/r02_gdn_tc_ncu_raw.csv:gdn_tc
"""
import csv
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = importlib.import_module("deep-learning-based-rgba-image-compression-with-masked-window-based-attention_b200.build")


def _num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def launches(path, kernel_substr):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    ki = names.index("Kernel Name")
    ir, iw, it = names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum"), names.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out = []
    for r in data:
        if len(r) > max(ir, iw) and kernel_substr in r[ki]:
            out.append(dict(read=_num(r[ir]) * scale[units[ir]], write=_num(r[iw]) * scale[units[iw]],
                            time_us=_num(r[it]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}[units[it]]))
    return out


def main():
    table = {}
    for spec in sys.argv[1:]:
        key, rest = spec.split("=", 1)
        path, kern = rest.rsplit(":", 1)
        ls = launches(path, kern)
        if not ls:
            raise SystemExit(f"no launch of {kern} in {path}")
        with open(path, "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        per = [l["read"] + l["write"] for l in ls]
        entry = {"kernel": kern, "launches_in_capture": len(ls), "dram_bytes_per_launch": sum(per) / len(per),
                 "dram_bytes_read": [l["read"] for l in ls], "dram_bytes_written": [l["write"] for l in ls],
                 "time_us_under_ncu": [l["time_us"] for l in ls],
                 "source": f"{os.path.relpath(path, ROOT)} (sha256 {sha[:16]}, library digest {B._digest()[:16]})"}
        if key == "gdn":
            entry["dram_bytes_per_step"] = sum(per)
        table[key] = entry
    out = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    with open(out, "w") as f:
        json.dump(table, f, indent=1)
    print(json.dumps(table, indent=1)[:1500])


if __name__ == "__main__":
    main()
