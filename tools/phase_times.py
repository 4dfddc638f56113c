#!/usr/bin/env python
"""per-phase clock64 totals of the tcgen05 attention kernel (CTA 0): python tools/phase_times.py [attn8|attn4]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
what = sys.argv[1] if len(sys.argv) > 1 else "attn8"
dev = torch.device("cuda:0")
C, h, ws, s, H, W = (192, 8, 8, 4, 128, 192) if what == "attn8" else (80, 8, 4, 2, 64, 96)
m = pkg.MaskedWinBasedAttention(C, h, ws, s).to(dev); m.algo = pkg.ALGO_TCGEN05
x = torch.randn(16, C, H, W, device=dev); a = torch.ones(16, 1, H, W, device=dev)
lib = pkg._abi.load()
with torch.no_grad():
    for _ in range(3): m(x, a)
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    lib.mwa_debug_set_timing_buffer(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
    lib.mwa_debug_set_timing_buffer(None)
t = buf.cpu().tolist(); tiles = max(t[15], 1)
names = ["prologue", "x load+convert", "QKV issue group 0 (slab waits)", "QKV MMA wait", "qkv drain", "QKV issue next group", "attention core (HMMA)",
         "-", "-", "-", "-", "proj issue (slab wait)", "residual loads issue", "last proj wait", "epilogue stores"]
tot = sum(t[:15])
print("  acquire_q wait cycles per tile by K block:", [round(v / tiles) for v in t[16:20]])
print(f"{what}: kernel+scan {e0.elapsed_time(e1)*1e3:.0f} us, CTA0 tiles {tiles}, cycles/tile {sum(t[1:15])/tiles:.0f}")
for n, v in zip(names, t):
    print(f"  {n:28s} {v:10d} cyc  {100*v/tot:5.1f}%  per tile {v/tiles:8.0f}")
