#!/usr/bin/env python
"""per-stage clock64 totals of the warp-specialised attention kernel (CTA 0, first warp of each role):
python tools/phase_times.py [attn8|attn4]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
what = sys.argv[1] if len(sys.argv) > 1 else "attn8"
dev = torch.device("cuda:0")
C, h, ws, s, H, W = (192, 8, 8, 4, 128, 192) if what == "attn8" else (80, 8, 4, 2, 64, 96)
m = pkg.MaskedWinBasedAttention(C, h, ws, s).to(dev); m.algo = pkg.ALGO_TCGEN05
x = torch.randn(16, C, H, W, device=dev); a = torch.ones(16, 1, H, W, device=dev)
keep_frac = 1.0
if len(sys.argv) > 2:      # fraction of kept windows (blobs of 4x4 windows in the shifted frame)
    keep_frac = float(sys.argv[2])
    torch.manual_seed(1)
    blob = (torch.rand(16, 1, H // ws // 4, W // ws // 4, device=dev) < keep_frac).float()
    a = blob.repeat_interleave(4 * ws, 2).repeat_interleave(4 * ws, 3)
    a = torch.roll(a, (s, s), (2, 3))
lib = pkg._abi.load()
with torch.no_grad():
    for _ in range(3): m(x, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)
    buf = torch.zeros(4096, dtype=torch.int64, device=dev)
    lib.mwa_debug_set_timing_buffer(buf.data_ptr())
    e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
    lib.mwa_debug_set_timing_buffer(None)
t = buf.cpu().tolist()
nwin = int(16 * (H // ws) * (W // ws) * keep_frac); tiles_cta = max(1, -(-(nwin * ws * ws // 128) // 148))
print(f"{what}: scan+compact+kernel {plain*1e3:.0f} us (timing build {e0.elapsed_time(e1)*1e3:.0f} us), ~{tiles_cta} tiles per CTA")
roles = [("MMA issuer", 0, ["wait X full", "wait D_qkv drained", "QKV issue + slab waits", "wait O_g", "wait proj acc free", "proj issue + slab wait"]),
         ("x producer / epilogue", 8, ["prologue", "wait X free", "store X", "load x (tile+2)", "wait proj complete", "epilogue"]),
         ("attention", 16, ["wait D_qkv", "drain", "wait O buffer", "attention core", "tile set-up"])]
# event trace of CTA 0: [1024 + role * 1024 + i] = stage << 48 | cycles since kernel start at the END of the stage
ev, stage_tot = [], {}
for role_i in range(3):
    prev = 0
    for v in t[1024 + role_i * 1024: 2048 + role_i * 1024]:
        if v == 0:
            break
        slot, when = v >> 48, v & ((1 << 48) - 1)
        ev.append((when, role_i, slot, when - prev))
        stage_tot[slot] = stage_tot.get(slot, 0) + when - prev
        prev = when
for role, off, names in roles:
    tot = sum(stage_tot.get(off + i, 0) for i in range(len(names)))
    print(f"  {role}: {tot / tiles_cta:.0f} cycles per tile")
    for i, n in enumerate(names):
        v = stage_tot.get(off + i, 0)
        print(f"    {n:28s} {v:10d} cyc  {100 * v / max(tot, 1):5.1f}%  per tile {v / tiles_cta:8.0f}")
cyc = torch.tensor(t[64:64 + 148], dtype=torch.float64); til = torch.tensor(t[320:320 + 148], dtype=torch.float64)
print(f"  per-CTA kernel cycles: min {cyc.min():.0f} median {cyc.median():.0f} max {cyc.max():.0f}; tiles per CTA min {til.min():.0f} max {til.max():.0f}; "
      f"cycles/tile median {(cyc / til.clamp(min=1)).median():.0f} max {(cyc / til.clamp(min=1)).max():.0f}")

names = {}
for role, off, ns in roles:
    for i, n in enumerate(ns):
        names[off + i] = (role.split()[0][:3].upper(), n)
ev.sort()
per_tile = (cyc[0] / max(til[0], 1)).item()
lo, hi = 6 * per_tile, 8.2 * per_tile
print(f"  timeline of CTA 0, cycles {lo:.0f} .. {hi:.0f} (about tiles 6-8; stage END time, [duration]):")
for when, role_i, slot, dur in ev:
    if lo <= when <= hi:
        tag, n = names[slot]
        print(f"    {when:8d}  {'':{role_i * 34}s}{tag} {n} [{dur}]")
