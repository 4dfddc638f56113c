#!/usr/bin/env python
"""debug: split-precision tcgen05 attention (ALGO_AUTO) vs the fp32 SIMT kernel and the fp64 oracle; error / tolerance"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
from oracle import ref_ops as R
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(C, heads, ws, s, B, H, W, alpha_mode, scale=1.0, oracle=False):
    m = pkg.MaskedWinBasedAttention(C, heads, ws, s).to(dev)
    x = torch.randn(B, C, H, W, device=dev) * scale
    if alpha_mode == "ones":
        a = torch.ones(B, 1, H, W, device=dev)
    else:
        a = (torch.rand(B, 1, H // ws, W // ws, device=dev) > 0.4).float().repeat_interleave(ws, 2).repeat_interleave(ws, 3)
        a = torch.roll(a, (s, s), (2, 3))
    with torch.no_grad():
        m.algo = pkg.ALGO_SIMT; y0 = m(x, a)
        m.algo = pkg.ALGO_AUTO; y1 = m(x, a)
    torch.cuda.synchronize()
    ref = y0
    if oracle:
        p = {k: v.detach().double().cpu() for k, v in dict(qkv_w=m.attn.qkv.weight, qkv_b=m.attn.qkv.bias, proj_w=m.attn.proj.weight,
                                                          proj_b=m.attn.proj.bias, table=m.attn.relative_position_bias_table).items()}
        ref = R.masked_window_attention(x.double().cpu(), a.double().cpu(), p["qkv_w"], p["qkv_b"], p["proj_w"], p["proj_b"], p["table"],
                                        heads, ws, s).float().to(dev)
    d = (y1 - ref).abs()
    tol = 1e-4 + 1e-3 * ref.abs()
    dropped_exact = True
    dw = R.to_windows(torch.roll(d.permute(0, 2, 3, 1), (-s, -s), (1, 2)).cpu(), ws).reshape(-1, ws * ws, C)
    per_win = dw.amax(dim=(1, 2))
    bad = (per_win > 2e-3).nonzero().flatten().tolist()
    nan = int(torch.isnan(y1).sum())
    print(f"C={C} h={heads} s={s} B={B} {H}x{W} alpha={alpha_mode} x*{scale}: max {d.max().item():.3e} rms {d.pow(2).mean().sqrt().item():.2e} "
          f"worst err/tol {(d / tol).max().item():.2f} nan {nan} bad windows {len(bad)}/{per_win.numel()} first {bad[:12]}", flush=True)
    if bad:
        per_ch = dw.amax(dim=(0, 1)); per_tok = dw[bad[0]].amax(dim=1)
        print("   per-channel max (every 8th):", [f"{v:.1e}" for v in per_ch[::8].tolist()])
        print("   per-token max of first bad window:", [f"{v:.1e}" for v in per_tok.tolist()[:64:4]])
cfgs = [(192, 8, 8, 0, 1, 8, 16, "ones"), (192, 8, 8, 0, 1, 16, 32, "ones"), (192, 8, 8, 0, 1, 8, 8, "ones"), (192, 8, 8, 4, 1, 16, 32, "ones"),
        (192, 8, 8, 4, 2, 64, 96, "blob"), (192, 8, 8, 4, 3, 24, 40, "blob"), (192, 6, 8, 4, 1, 32, 48, "blob"),
        (192, 8, 8, 4, 16, 128, 192, "blob"), (192, 6, 8, 4, 4, 128, 192, "blob"), (192, 8, 8, 4, 16, 128, 192, "ones")]
for cfg in cfgs:
    run(*cfg)
run(192, 8, 8, 4, 2, 64, 96, "blob", 1.0, True)
run(192, 8, 8, 4, 2, 64, 96, "blob", 3.0, True)
run(192, 6, 8, 4, 2, 64, 96, "blob", 1.0, True)
run(192, 6, 8, 4, 2, 64, 96, "blob", 3.0, True)
run(192, 8, 8, 4, 2, 64, 96, "blob", 10.0, True)
run(192, 8, 8, 4, 1, 8, 8, "ones", 1.0, True)
run(192, 6, 8, 0, 1, 8, 24, "ones", 1.0, True)
# quick timing
m = pkg.MaskedWinBasedAttention(192, 8, 8, 4).to(dev)
x = torch.randn(16, 192, 128, 192, device=dev); a = torch.ones(16, 1, 128, 192, device=dev)
for algo, name in ((pkg.ALGO_AUTO, "split"), (4, "fp16-ws")):
    m.algo = algo
    with torch.no_grad():
        for _ in range(3): m(x, a)
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): m(x, a)
        e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per call (6144 windows kept)")
