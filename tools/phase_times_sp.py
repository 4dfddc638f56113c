#!/usr/bin/env python
"""per-stage clock64 totals + timeline of the split-precision attention kernel (CTA 0, first warp of each role):
python tools/phase_times_sp.py [heads] [keep_frac]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
heads = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
C, ws, s, H, W = 192, 8, 4, 128, 192
m = pkg.MaskedWinBasedAttention(C, heads, ws, s).to(dev); m.algo = pkg.ALGO_AUTO
x = torch.randn(16, C, H, W, device=dev); a = torch.ones(16, 1, H, W, device=dev)
keep_frac = 1.0
if len(sys.argv) > 2:
    keep_frac = float(sys.argv[2]); torch.manual_seed(1)
    blob = (torch.rand(16, 1, H // ws // 4, W // ws // 4, device=dev) < keep_frac).float()
    a = torch.roll(blob.repeat_interleave(4 * ws, 2).repeat_interleave(4 * ws, 3), (s, s), (2, 3))
lib = pkg._abi.load()
with torch.no_grad():
    for _ in range(3): m(x, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    lib.mwa_debug_set_timing_buffer(buf.data_ptr())
    e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
    lib.mwa_debug_set_timing_buffer(None)
t = buf.cpu().tolist()
cyc = torch.tensor(t[64:64 + 148], dtype=torch.float64); til = torch.tensor(t[320:320 + 148], dtype=torch.float64)
tiles0 = max(til[0].item(), 1)
print(f"h={heads}: all launches {plain*1e3:.0f} us (timing build {e0.elapsed_time(e1)*1e3:.0f} us); per-CTA cycles min {cyc.min():.0f} median {cyc.median():.0f} "
      f"max {cyc.max():.0f}; tiles/CTA {til.min():.0f}-{til.max():.0f}; cycles/tile median {(cyc / til.clamp(min=1)).median():.0f}")
names = {0: "wait P,V + PV issue", 1: "wait Q,K + S issue", 2: "QKV(G+2) issue+waits", 3: "wait OA/acc + proj issue",
         8: "set-up + wait X_hi free", 9: "wait proj complete", 10: "epilogue+conversion", 11: "wait X_lo free", 12: "X_lo->TMEM", 13: "(acc released)",
         16: "wait S", 17: "softmax",
         24: "wait D_qkv", 25: "q->TMEM (+wait QK free)", 26: "k->smem", 27: "ld v + wait V free", 28: "v->smem", 29: "wait O", 30: "normalise"}
roles = ["MMA issuer", "converter/epilogue", "softmax", "drain+normalise"]
ev, tot = [], {}
for role_i in range(4):
    prev = 0
    for v in t[1024 + role_i * 1024: 2048 + role_i * 1024]:
        if v == 0: break
        slot, when = v >> 48, v & ((1 << 48) - 1)
        ev.append((when, role_i, slot, when - prev)); tot[slot] = tot.get(slot, 0) + when - prev; prev = when
for role_i, role in enumerate(roles):
    slots = [k for k in names if k // 8 == role_i]
    rt = sum(tot.get(k, 0) for k in slots)
    print(f"  {role}: {rt / tiles0:.0f} cycles per tile (traced part)")
    for k in slots:
        print(f"    {names[k]:28s} per tile {tot.get(k, 0) / tiles0:8.0f}  {100 * tot.get(k, 0) / max(rt, 1):5.1f}%")
ev.sort()
per_tile = cyc[0].item() / tiles0
lo, hi = 3 * per_tile, 3.4 * per_tile
print(f"  timeline of CTA 0, cycles {lo:.0f} .. {hi:.0f} (stage END time, [duration]):")
for when, role_i, slot, dur in ev:
    if lo <= when <= hi:
        print(f"    {when:8d}  {'':{role_i * 30}s}{names[slot]} [{dur}]")
