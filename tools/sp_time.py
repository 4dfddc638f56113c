#!/usr/bin/env python
"""development: time the default (split-precision) 8x8 attention alone, L2 flushed between launches, and check it against
the fp32 SIMT kernel on a small case:  python tools/sp_time.py [--lib build/variants/x.so] [--phase]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
if "--lib" in sys.argv:
    pkg._abi.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
dev = torch.device("cuda:0")
torch.manual_seed(0)
flush = torch.zeros(128 * 1024 * 1024, device=dev)
def t(m, x, a, iters=15):
    with torch.no_grad():
        for _ in range(3): m(x, a)
        ts = []
        for _ in range(iters):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); m(x, a); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
for heads in (8, 6):
    m = pkg.MaskedWinBasedAttention(192, heads, 8, 4).to(dev)
    xs = torch.randn(2, 192, 64, 96, device=dev); as_ = torch.ones(2, 1, 64, 96, device=dev)
    with torch.no_grad():
        m.algo = pkg.ALGO_SIMT; y0 = m(xs, as_); m.algo = pkg.ALGO_AUTO; y1 = m(xs, as_)
    err = ((y1 - y0).abs() / (1e-4 + 1e-3 * y0.abs())).max().item()
    x = torch.randn(16, 192, 128, 192, device=dev)
    res = []
    for keep in (1.0, 0.5):
        a = torch.ones(16, 1, 128, 192, device=dev)
        if keep < 1:
            a = (torch.rand(16, 1, 16, 24, device=dev) < keep).float().repeat_interleave(8, 2).repeat_interleave(8, 3)
            a = torch.roll(a, (4, 4), (2, 3))
        ms = t(m, x, a)
        kept = 6144 * keep
        res.append(f"keep {keep:.0%}: {ms:.3f} ms ({kept * 22.020096e6 / ms / 1e9:.0f} TFLOP/s algorithmic)")
    print(f"h={heads}: worst err/tol {err:.3f}; " + "; ".join(res), flush=True)
