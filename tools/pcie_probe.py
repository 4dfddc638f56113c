#!/usr/bin/env python
"""Host<->device copy ceiling of the box (pinned memory, CUDA events): H2D alone, D2H alone, both at once on two
streams.  bench.py's `e2e` leg moves every operand of the hot path over this link inside the timed region, so these
numbers are its roofline (development / evidence tool)."""
import torch

dev = torch.device("cuda:0")
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, a, b in (("H2D", True, False), ("D2H", False, True), ("H2D + D2H concurrently", True, True)):
    ms = run(a, b)
    print(f"{name:24s} 1 GiB each: {ms:7.2f} ms  = {n / ms / 1e6:6.1f} GB/s per direction", flush=True)
