#!/usr/bin/env python
"""The reference's hot path as it runs TODAY on a GPU -- PyTorch eager ops dispatched to ATen / cuBLAS / cuDNN, same
B200 -- timed beside the hand-written kernels at the BASELINE config-2 shapes (batch 16, 768x512).

SURVEY.md section 2.2: the reference ships no CUDA of its own; "the bar is the reference's PyTorch eager path on the
same B200".  The eager side is a re-statement of the reference's op sequence in plain torch (cat / roll / partition /
boolean-index gather with its host syncs / Linear / bmm / softmax / index_put scatter for the attention block,
layers/masked_win_attention.py:169-251; pow / conv2d / sqrt / div for GDN, layers/GDN.py:74-90; round-sub-add for the
rounding, models/AutoEncoderRGB_Journal.py:31-32) with the reference's default precision flags (cuDNN TF32 on, matmul TF32
off).  Development / evidence tool: one line per op, CUDA events, L2 flushed between iterations.
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import mwa_b200 as pkg  # noqa: E402


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(w, ws, H, W):
    B = int(w.shape[0] / (H * W / ws / ws))
    x = w.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def eager_attention(m, x, alpha):
    """op-for-op what layers/masked_win_attention.py:169-251 + :96-131 launch (incl. the three nonzero() host syncs)"""
    a, ws, s, h = m.attn, m.window_size, m.shift_size, m.num_heads
    B, C, H, W = x.shape
    shortcut = x
    xs, al = x.permute(0, 2, 3, 1), alpha.permute(0, 2, 3, 1)
    if s > 0:
        cat = torch.roll(torch.cat([xs, al], dim=-1), shifts=(-s, -s), dims=(1, 2))
        xs, al = cat[..., :C], cat[..., C:]
    cat = window_partition(torch.cat([xs, al], dim=-1), ws)
    xw, aw = cat[..., :C], cat[..., C:]
    keep = aw.sum(dim=(1, 2, 3)) != 0
    mask = None
    if s > 0:
        img = torch.zeros((B, H, W, 1), device=x.device)
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
            for wsl in (slice(0, -ws), slice(-ws, -s), slice(-s, None)):
                img[:, hs, wsl, :] = cnt
                cnt += 1
        mw = window_partition(img, ws)[keep].view(-1, ws * ws)
        mask = mw.unsqueeze(1) - mw.unsqueeze(2)
        mask = mask.masked_fill(mask != 0, -100.0).masked_fill(mask == 0, 0.0)
    xk = xw[keep].view(-1, ws * ws, C)
    K, N = xk.shape[0], ws * ws
    qkv = a.qkv(xk).reshape(K, N, 3, h, C // h).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * a.scale, qkv[1], qkv[2]
    att = q @ k.transpose(-2, -1)
    att = att + a.relative_position_bias_table[a.relative_position_index.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        att = att + mask.unsqueeze(1)
    att = torch.softmax(att, dim=-1)
    y = a.proj((att @ v).transpose(1, 2).reshape(K, N, C))
    res = torch.zeros_like(xw)
    res[keep] = y.view(-1, ws, ws, C)
    xs = window_reverse(res, ws, H, W)
    if s > 0:
        xs = torch.roll(xs, shifts=(s, s), dims=(1, 2))
    return shortcut + xs.permute(0, 3, 1, 2)


def eager_gdn(m, x):
    """layers/GDN.py:74-90"""
    ped = m.pedestal
    beta = torch.max(m.beta, torch.ones_like(m.beta) * m.beta_bound) ** 2 - ped
    gamma = torch.max(m.gamma, torch.ones_like(m.gamma) * m.gamma_bound) ** 2 - ped
    C = x.shape[1]
    norm = torch.sqrt(F.conv2d(x ** 2, gamma.view(C, C, 1, 1), beta))
    return x * norm if m.inverse else x / norm


def eager_round(op):
    outs = [torch.round(op["mask"] * 255) / 255]
    z = op["z"] - op["med"]
    outs.append(torch.round(z) - z.detach() + z + op["med"])
    for ys, ms, ls in zip(op["y"].chunk(10, 1), op["mu"].chunk(10, 1), op["lrp"].chunk(10, 1)):
        d = ys - ms
        outs.append(torch.round(d) - d.detach() + d + ms + 0.5 * torch.tanh(ls))
    return outs


def main():
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = True            # the reference's (= torch's) defaults
    torch.backends.cuda.matmul.allow_tf32 = False
    flush = torch.zeros(128 * 1024 * 1024, device=dev)
    ops = bench.to_device(bench.build_workload(pkg, dev, bench.BATCH_PER_GPU, seed0=0), dev)
    tot_e = tot_o = 0.0
    with torch.no_grad():
        for op in ops:
            if op["kind"] == "attn":
                eager = lambda op=op: eager_attention(op["mod"], op["x"], op["alpha"])
                ours = lambda op=op: op["mod"](op["x"], op["alpha"])
                err = (eager() - ours()).abs().max().item()
            elif op["kind"] == "gdn":
                eager = lambda op=op: eager_gdn(op["mod"], op["x"])
                ours = lambda op=op: op["mod"](op["x"])
                err = (eager() - ours()).abs().max().item()
            else:
                eager = lambda op=op: eager_round(op)
                ours = lambda op=op: bench.run_op(pkg, op)
                err = max((a - b).abs().max().item() for a, b in zip(eager(), ours()))
            te, to = timeit(eager, 5, flush), timeit(ours, 10, flush)
            tot_e += te
            tot_o += to
            print(f"{op['name'][:46]:46s} eager torch {te:8.3f} ms   B200 kernels {to:7.3f} ms   {te / to:6.1f}x   "
                  f"max |diff| {err:.2e}", flush=True)
    print(f"{'hot path of one batch-16 step':46s} eager torch {tot_e:8.3f} ms   B200 kernels {tot_o:7.3f} ms   "
          f"{tot_e / tot_o:6.1f}x   = {16 / tot_e * 1e3:.0f} vs {16 / tot_o * 1e3:.0f} images/s", flush=True)


if __name__ == "__main__":
    main()
