#!/usr/bin/env python
"""debug: hunt rare corruption in the split-precision attention kernel.  usage: dbg_sp_race.py [lib.so] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
if len(sys.argv) > 1 and sys.argv[1].endswith(".so"):
    pkg._abi.LIB_PATH = os.path.abspath(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
from oracle import ref_ops as R
dev = torch.device("cuda:0")
torch.manual_seed(0)
C, heads, ws, s, B, H, W = 192, 8, 8, 4, 16, 128, 192
m = pkg.MaskedWinBasedAttention(C, heads, ws, s).to(dev)
x = torch.randn(B, C, H, W, device=dev)
a = torch.ones(B, 1, H, W, device=dev)
with torch.no_grad():
    m.algo = pkg.ALGO_SIMT; y0 = m(x, a)
    m.algo = pkg.ALGO_AUTO
    for rep in range(reps):
        y1 = m(x, a)
        torch.cuda.synchronize()
        d = (y1 - y0).abs()
        dw = R.to_windows(torch.roll(d.permute(0, 2, 3, 1), (-s, -s), (1, 2)).cpu(), ws).reshape(-1, ws * ws, C)
        per_win = dw.amax(dim=(1, 2))
        bad = (per_win > 2e-3).nonzero().flatten().tolist()
        print(f"rep {rep}: max {d.max().item():.3e} bad windows {len(bad)}", flush=True)
        for wi in bad[:6]:
            b, r = divmod(wi, (H // ws) * (W // ws)); wy, wx = divmod(r, W // ws)
            e = dw[wi]                                  # (64 tok, C)
            big_ch = (e.amax(0) > 0.1).nonzero().flatten().tolist()
            big_tok = (e.amax(1) > 0.1).nonzero().flatten().tolist()
            print(f"   win {wi} (b {b} wy {wy} wx {wx}): channels with err>0.1: {big_ch[:20]} (n={len(big_ch)}), tokens: n={len(big_tok)} "
                  f"first {big_tok[:8]}; median small err {e.median().item():.1e}")
