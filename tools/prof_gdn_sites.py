#!/usr/bin/env python
"""profiling target: the six GDN / IGDN launches of one forward at BASELINE config-2 shapes (batch 16), once each, as the
forward runs them: gdn1 / gdn3 / igdn1 / igdn3 write their consumer's fp16 hi / lo planes (ps 2, 1, 1, 1), gdn2 / igdn2 the
dense tensor the attention reads"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mwa_b200 as pkg
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.no_grad():
    for warm in (True, False):
        for d, inv, ps in ((2, False, 2), (4, False, 0), (8, False, 1), (8, True, 1), (4, True, 0), (2, True, 1)):
            m = pkg.GDN(192, inverse=inv)
            m.gamma.add_(torch.rand(192, 192) * 0.02)
            m = m.to(dev)
            x = torch.randn(16, 192, 512 // d, 768 // d, device=dev)
            y = m.request_planes(ps)(x) if ps else m(x)
            if warm:
                break
torch.cuda.synchronize()
print("done")
