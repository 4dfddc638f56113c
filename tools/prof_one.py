#!/usr/bin/env python
"""Launch one hot-path kernel a few times (profiling target for ncu): python tools/prof_one.py gdn|igdn|attn8|attn4"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg  # noqa: E402
if os.environ.get("MWA_B200_LIB"):          # development: a variant library (tools/build_variants.py)
    pkg._abi.LIB_PATH = os.path.abspath(os.environ["MWA_B200_LIB"])

what = sys.argv[1] if len(sys.argv) > 1 else "gdn"
algo = {"auto": 0, "simt": 1, "tc": 2}[sys.argv[2] if len(sys.argv) > 2 else "auto"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.no_grad():
    if what in ("gdn", "igdn"):
        m = pkg.GDN(192, inverse=(what == "igdn"))
        m.gamma.add_(torch.rand(192, 192) * 0.02)
        m = m.to(dev)
        m.algo = algo
        x = torch.randn(16, 192, 256, 384, device=dev)
        for _ in range(5):
            y = m(x)
    else:
        C, h, ws, s, H, W = (192, 8, 8, 4, 128, 192) if what == "attn8" else (80, 8, 4, 2, 64, 96)
        m = pkg.MaskedWinBasedAttention(C, h, ws, s).to(dev)
        m.algo = algo
        x = torch.randn(16, C, H, W, device=dev)
        if len(sys.argv) > 3 and sys.argv[3] == "blob":      # 4x4-window blobs, 50 % kept
            a = (torch.rand(16, 1, H // ws // 4, W // ws // 4, device=dev) < 0.5).float().repeat_interleave(4 * ws, 2).repeat_interleave(4 * ws, 3)
            a = torch.roll(a, (s, s), (2, 3))
        elif len(sys.argv) > 3 and sys.argv[3] == "ones":
            a = torch.ones(16, 1, H, W, device=dev)
        else:
            a = (torch.rand(16, 1, H // ws, W // ws, device=dev) > 0.4).float().repeat_interleave(ws, 2).repeat_interleave(ws, 3)
        for _ in range(5):
            y = m(x, a)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
