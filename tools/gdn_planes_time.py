#!/usr/bin/env python
"""development (B200 box): GDN at the transform's call sites, dense result + conv_act_split against gdn_forward_planes"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg  # noqa: E402

dev = torch.device("cuda:0")
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


with torch.no_grad():
    for (H, W, ps, inv) in [(256, 384, 2, False), (256, 384, 1, True), (64, 96, 1, False)]:
        m = pkg.GDN(192, inverse=inv).to(dev)
        x = torch.randn(16, 192, H, W, device=dev)
        sp = pkg.conv.SplitAct.empty(16, 192, H, W, ps, dev)
        m(x)
        m.request_planes(ps)(x)
        t_dense = timed(lambda: m(x))
        t_split = timed(lambda: pkg.conv.split_into(m(x), sp))
        t_planes = timed(lambda: m.request_planes(ps)(x))
        gb = 16 * H * W * 1536 / 1e9
        print(f"{H}x{W} ps={ps} inverse={inv}: dense {t_dense:7.1f} us ({gb / t_dense * 1e6:6.0f} GB/s)   dense + split {t_split:7.1f} us   "
              f"planes {t_planes:7.1f} us ({gb / t_planes * 1e6:6.0f} GB/s)")
