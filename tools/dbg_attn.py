#!/usr/bin/env python
"""debug: tcgen05 attention vs fp32 SIMT kernel, error broken down by window / channel"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
from oracle import ref_ops as R
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(C, heads, ws, s, B, H, W, alpha_mode):
    m = pkg.MaskedWinBasedAttention(C, heads, ws, s).to(dev)
    x = torch.randn(B, C, H, W, device=dev)
    if alpha_mode == "ones":
        a = torch.ones(B, 1, H, W, device=dev)
    else:
        a = (torch.rand(B, 1, H // ws, W // ws, device=dev) > 0.4).float().repeat_interleave(ws, 2).repeat_interleave(ws, 3)
        a = torch.roll(a, (s, s), (2, 3))
    with torch.no_grad():
        m.algo = pkg.ALGO_SIMT; y0 = m(x, a)
        m.algo = pkg.ALGO_TCGEN05; y1 = m(x, a)
    torch.cuda.synchronize()
    d = (y1 - y0).abs()
    # per window (shifted frame) max error
    dw = R.to_windows(torch.roll(d.permute(0, 2, 3, 1), (-s, -s), (1, 2)).cpu(), ws).reshape(-1, ws * ws, C)
    per_win = dw.amax(dim=(1, 2))
    per_ch = dw.amax(dim=(0, 1))
    per_tok = dw.amax(dim=(0, 2))
    bad = (per_win > 2e-3).nonzero().flatten().tolist()
    print(f"C={C} h={heads} ws={ws} s={s} B={B} {H}x{W} alpha={alpha_mode}: max {d.max().item():.3e} mean {d.mean().item():.3e} "
          f"bad windows {len(bad)}/{per_win.numel()} first {bad[:24]}")
    if bad:
        print("   per-channel max (first 80):", [f"{v:.1e}" for v in per_ch[:80:4].tolist()])
        print("   per-token max:", [f"{v:.1e}" for v in per_tok.tolist()[:16]])
for cfg in [(192, 8, 8, 0, 1, 8, 16, "ones"), (192, 8, 8, 0, 1, 16, 32, "ones"), (192, 8, 8, 4, 2, 64, 96, "blob"),
            (192, 8, 8, 4, 16, 128, 192, "blob"), (192, 6, 8, 4, 4, 128, 192, "blob"), (80, 8, 4, 2, 16, 64, 96, "blob"),
            (192, 8, 8, 4, 3, 24, 40, "blob"), (80, 8, 4, 2, 3, 12, 20, "blob"),
            (80, 8, 4, 0, 1, 8, 16, "ones"), (80, 8, 4, 0, 1, 16, 32, "ones"), (80, 8, 4, 2, 1, 16, 32, "ones"), (80, 8, 4, 2, 2, 16, 32, "blob"),
            (192, 8, 8, 0, 1, 8, 16, "ones"), (192, 8, 8, 4, 1, 32, 48, "blob"), (192, 6, 8, 4, 1, 32, 48, "blob")]:
    run(*cfg)
