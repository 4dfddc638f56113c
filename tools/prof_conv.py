#!/usr/bin/env python
"""Launch one convolution of the codec a few times (profiling target for ncu):
   python tools/prof_conv.py ru1x1b|ru3x3|ru1x1a|dse3x3|cc0|cc2|cc4|x2|dx3|dx4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg  # noqa: E402

if os.environ.get("MWA_LIB"):              # A/B against a variant library
    pkg._abi.LIB_PATH = os.path.abspath(os.environ["MWA_LIB"])

B, H, W = 16, 512, 768
# kind, cin, cout, k, stride, h, w, act, residual
CASES = {
    "ru1x1a": ("conv", 192, 96, 1, 1, H // 4, W // 4, 1, False),
    "ru3x3": ("conv", 96, 96, 3, 1, H // 4, W // 4, 1, False),
    "ru1x1b": ("conv", 96, 192, 1, 1, H // 4, W // 4, 1, True),
    "dse3x3": ("conv", 32, 32, 3, 1, H, W, 2, False),
    "dseout": ("conv", 32, 3, 1, 1, H, W, 0, True),
    "x4": ("conv", 192, 80, 1, 1, H // 8, W // 8, 0, False),
    "cc0": ("conv", 120, 224, 3, 1, H // 8, W // 8, 1, False),
    "cc2": ("conv", 224, 128, 3, 1, H // 8, W // 8, 1, False),
    "cc4": ("conv", 128, 8, 3, 1, H // 8, W // 8, 0, False),
    "x2": ("conv", 192, 192, 5, 2, H // 2, W // 2, 0, False),
    "dx3": ("deconv", 192, 192, 5, 2, H // 4, W // 4, 0, False),
    "dx4": ("deconv", 192, 3, 5, 2, H // 2, W // 2, 0, False),
}
what = sys.argv[1]
kind, cin, cout, k, s, h, w, act, res = CASES[what]
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.no_grad():
    x = torch.randn(B, cin, h, w, device=dev)
    if kind == "conv":
        m = pkg.conv.Conv2d(cin, cout, k, stride=s, padding=k // 2).to(dev)
    else:
        m = pkg.conv.ConvTranspose2d(cin, cout, k, stride=s, padding=k // 2, output_padding=1).to(dev)
    y = m(x, act=act)
    r = torch.randn_like(y) if (res and "nores" not in sys.argv) else None
    kw = dict(emit_ps=1)
    if "nodense" in sys.argv:              # attribution runs: planes only / dense only / no residual
        kw["want_dense"] = False
    if "noplanes" in sys.argv:
        kw = {}
    planes = "planes" in sys.argv          # as inside a chain: planes in, dense + planes out
    if planes:
        xs = pkg.conv.split_into(x, pkg.conv.SplitAct.empty(B, cin, h, w, 2 if (kind == "conv" and s == 2) else 1, dev))
    for _ in range(4):
        y = m(xs, act=act, residual=r, **kw) if planes else m(x, act=act, residual=r)
    if "time" in sys.argv:                 # CUDA-event time per launch (not under ncu)
        big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(10):
            big.zero_()                    # flush L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = m(xs, act=act, residual=r, **kw) if planes else m(x, act=act, residual=r)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{what:8s} {'planes' if planes else 'dense '} median {ts[len(ts) // 2] * 1e3:8.1f} us   min {ts[0] * 1e3:8.1f} us")
torch.cuda.synchronize()
y = y.dense if hasattr(y, "dense") else y
print("done", float(y.abs().mean()) if y is not None else "(planes only)")
