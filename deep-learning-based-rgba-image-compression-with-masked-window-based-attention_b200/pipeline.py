"""Host <-> device pipeline around the codec's forward (serving: images arrive in host memory, reconstructions leave to it).

    pipe = HostPipeline(net)                       # net: RGBACodec on a CUDA device
    pipe.run(host_rgba_batches, host_x_hat_outs)   # pinned (B, 4, H, W) in, pinned (B, 3, H, W) out, one pair per step

The H2D copy of batch i + 1 and the D2H copy of batch i - 1's reconstruction run on two copy streams while batch i computes
(two staging slots, events between the streams, no host synchronisation inside the loop); the forward itself is the
reference's call sequence (trainRGB.py:283-289: alpha pyramid, then AutoEncoder.forward with reconmask = the alpha)."""
from __future__ import annotations

import torch


class HostPipeline:
    def __init__(self, net, device=None):
        self.net = net
        self.device = device if device is not None else next(net.parameters()).device
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.stage = [None, None]
        self.loaded = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def forward(self, rgba):
        image, alpha = rgba[:, :3], rgba[:, 3:4]
        me = self.net.EncMakeMask(alpha)
        return self.net(image, alpha, alpha, me[0], me[1], me[2], me[3])

    @torch.no_grad()
    def run(self, host_batches, host_outs=None):
        """returns the list of per-batch (mse, bpp) device scalars; host_outs[i] receives x_hat of batch i"""
        main = torch.cuda.current_stream(self.device)
        results = []
        for i, hb in enumerate(host_batches):
            s = i & 1
            if self.stage[s] is None or self.stage[s].shape != hb.shape:
                self.stage[s] = torch.empty(hb.shape, dtype=hb.dtype, device=self.device)
            with torch.cuda.stream(self.h2d):
                if i >= 2:
                    self.h2d.wait_event(self.consumed[s])         # the forward of batch i - 2 has read this slot
                self.stage[s].copy_(hb, non_blocking=True)
                self.loaded[s].record(self.h2d)
            main.wait_event(self.loaded[s])
            x_hat, mse, bpp = self.forward(self.stage[s])[:3]
            self.consumed[s].record(main)
            results.append((mse, bpp))
            if host_outs is not None:
                done = torch.cuda.Event()
                done.record(main)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    host_outs[i].copy_(x_hat, non_blocking=True)
                x_hat.record_stream(self.d2h)
        main.wait_stream(self.d2h)
        main.wait_stream(self.h2d)
        return results
