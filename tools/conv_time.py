#!/usr/bin/env python
"""development: every distinct convolution of one AutoEncoderRGB_Journal forward at BASELINE config-2 shapes (batch 16,
768x512), timed alone: this repo's tcgen05 kernel vs torch (cuDNN) fp32 and TF32.  python tools/conv_time.py [--only name]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
dev = torch.device("cuda:0")
B, H, W = 16, 512, 768
# name, kind, cin, cout, k, stride, h, w, count per forward
L = [("enc.x1", "conv", 3, 192, 5, 2, H, W, 1), ("enc.x2", "conv", 192, 192, 5, 2, H // 2, W // 2, 1),
     ("enc.x3", "conv", 192, 192, 5, 2, H // 4, W // 4, 1), ("enc.x4", "conv", 192, 80, 1, 1, H // 8, W // 8, 1),
     ("ru192.1x1a", "conv", 192, 96, 1, 1, H // 4, W // 4, 12), ("ru192.3x3", "conv", 96, 96, 3, 1, H // 4, W // 4, 12),
     ("ru192.1x1b", "conv", 96, 192, 1, 1, H // 4, W // 4, 12), ("wrap192.1x1", "conv", 192, 192, 1, 1, H // 4, W // 4, 2),
     ("ru80.1x1a", "conv", 80, 40, 1, 1, H // 8, W // 8, 12), ("ru80.3x3", "conv", 40, 40, 3, 1, H // 8, W // 8, 12),
     ("ru80.1x1b", "conv", 40, 80, 1, 1, H // 8, W // 8, 12), ("wrap80.1x1", "conv", 80, 80, 1, 1, H // 8, W // 8, 2),
     ("h_a.0", "conv", 80, 320, 3, 2, H // 8, W // 8, 1), ("h_a.2", "conv", 320, 288, 3, 1, H // 16, W // 16, 1),
     ("h_a.4", "conv", 288, 256, 3, 2, H // 16, W // 16, 1), ("h_s.4", "conv", 224, 1024, 3, 1, H // 32, W // 32, 2),
     ("h_s.8", "conv", 288, 320, 3, 1, H // 16, W // 16, 2),
     ("cc.0 (120->224)", "conv", 120, 224, 3, 1, H // 8, W // 8, 30), ("cc.2 (224->128)", "conv", 224, 128, 3, 1, H // 8, W // 8, 30),
     ("cc.4 (128->8)", "conv", 128, 8, 3, 1, H // 8, W // 8, 30),
     ("dec.x1", "conv", 80, 192, 1, 1, H // 8, W // 8, 1), ("dec.x2", "deconv", 192, 192, 5, 2, H // 8, W // 8, 1),
     ("dec.x3", "deconv", 192, 192, 5, 2, H // 4, W // 4, 1), ("dec.x4", "deconv", 192, 3, 5, 2, H // 2, W // 2, 1),
     ("dse.in", "conv", 3, 32, 1, 1, H, W, 1), ("dse.3x3", "conv", 32, 32, 3, 1, H, W, 6), ("dse.out", "conv", 32, 3, 1, 1, H, W, 1)]
only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tot = {"ours": 0.0, "fp32": 0.0, "tf32": 0.0}
with torch.no_grad():
    for name, kind, cin, cout, k, s, h, w, cnt in L:
        if only and only not in name: continue
        x = torch.randn(B, cin, h, w, device=dev)
        if kind == "conv":
            m = pkg.conv.Conv2d(cin, cout, k, stride=s, padding=k // 2).to(dev)
            ref = lambda: F.conv2d(x, m.weight, m.bias, stride=s, padding=k // 2)
            fl = 2.0 * B * cout * cin * k * k * (h // s) * (w // s)
        else:
            m = pkg.conv.ConvTranspose2d(cin, cout, k, stride=s, padding=k // 2, output_padding=1).to(dev)
            ref = lambda: F.conv_transpose2d(x, m.weight, m.bias, stride=s, padding=k // 2, output_padding=1)
            fl = 2.0 * B * cout * cin * k * k * h * w
        ours = t(lambda: m(x))
        torch.backends.cudnn.allow_tf32 = False
        fp32 = t(ref)
        torch.backends.cudnn.allow_tf32 = True
        tf32 = t(ref)
        torch.backends.cudnn.allow_tf32 = False
        err = float((m(x) - ref()).abs().max())
        tot["ours"] += ours * cnt; tot["fp32"] += fp32 * cnt; tot["tf32"] += tf32 * cnt
        print(f"{name:18s} x{cnt:2d} {fl/1e9:7.1f} GF  ours {ours:7.3f} ms {fl/ours/1e9:6.0f} TF/s | cudnn fp32 {fp32:7.3f} ms {fl/fp32/1e9:5.0f} | tf32 {tf32:7.3f} ms {fl/tf32/1e9:5.0f} | max diff {err:.1e}", flush=True)
print("per forward (ms):", {k: round(v, 2) for k, v in tot.items()})
