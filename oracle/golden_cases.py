"""Seeded input recipes shared by make_golden.py (build container, live reference) and the tests
(anywhere).  TEST INFRASTRUCTURE ONLY.  Inputs are regenerated from torch's CPU generator (bit-stable
for a given torch build; an input checksum stored with every golden output detects drift), so the
committed fixtures only carry the reference's OUTPUTS.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def checksum(*tensors) -> int:
    c = 0
    for t in tensors:
        if t is not None:
            c = zlib.crc32(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes(), c)
    return c


def blob_alpha(B, H, W, ws, shift, drop_frac, seed, soft=True):
    """alpha (B,1,H,W) >= 0 whose SHIFTED-frame windows are all-zero with probability ~drop_frac."""
    g = _gen(seed)
    keep = (torch.rand(B, 1, H // ws, W // ws, generator=g) >= drop_frac).float()
    a = keep.repeat_interleave(ws, 2).repeat_interleave(ws, 3)
    if soft:
        a = a * torch.round(torch.rand(B, 1, H, W, generator=g) * 255) / 255
    return torch.roll(a, shifts=(shift, shift), dims=(2, 3))


ATTENTION_CASES = {
    # name: dict(C, heads, ws, shift, B, H, W, drop, masked, seed)
    "attn_c32_h4_ws4_s2": dict(C=32, heads=4, ws=4, shift=2, B=2, H=8, W=12, drop=0.3, masked=True, seed=11),
    "attn_c80_h8_ws4_s2": dict(C=80, heads=8, ws=4, shift=2, B=2, H=16, W=24, drop=0.4, masked=True, seed=12),
    "attn_c192_h8_ws8_s4": dict(C=192, heads=8, ws=8, shift=4, B=1, H=24, W=32, drop=0.4, masked=True, seed=13),
    "attn_c192_h6_ws8_s0": dict(C=192, heads=6, ws=8, shift=0, B=2, H=16, W=16, drop=0.5, masked=True, seed=14),
    "attn_c192_h8_ws8_s4_unmasked": dict(C=192, heads=8, ws=8, shift=4, B=1, H=16, W=24, drop=0.0, masked=False,
                                         seed=15),
    "attn_c80_h8_ws4_s0_unmasked": dict(C=80, heads=8, ws=4, shift=0, B=1, H=8, W=8, drop=0.0, masked=False,
                                        seed=16),
    "attn_c192_all_transparent": dict(C=192, heads=8, ws=8, shift=4, B=1, H=16, W=16, drop=1.0, masked=True,
                                      seed=17),
    "attn_c48_h3_ws2_s1_nobias": dict(C=48, heads=3, ws=2, shift=1, B=1, H=6, W=4, drop=0.2, masked=True, seed=18,
                                      qkv_bias=False),
}


def attention_inputs(cfg):
    """returns dict(x, alpha|None, qkv_w, qkv_b|None, proj_w, proj_b, table)"""
    g = _gen(cfg["seed"])
    C, h, ws = cfg["C"], cfg["heads"], cfg["ws"]
    x = torch.randn(cfg["B"], C, cfg["H"], cfg["W"], generator=g) * 1.5
    s = C ** -0.5
    p = dict(
        x=x,
        qkv_w=torch.randn(3 * C, C, generator=g) * s * 1.5,
        qkv_b=(torch.randn(3 * C, generator=g) * 0.2) if cfg.get("qkv_bias", True) else None,
        proj_w=torch.randn(C, C, generator=g) * s,
        proj_b=torch.randn(C, generator=g) * 0.1,
        table=torch.randn((2 * ws - 1) ** 2, h, generator=g) * 0.5,
    )
    if cfg["masked"]:
        a = blob_alpha(cfg["B"], cfg["H"], cfg["W"], ws, cfg["shift"], cfg["drop"], cfg["seed"] + 1000)
        if cfg["drop"] >= 1.0:
            a = torch.zeros_like(a)
        elif cfg["drop"] > 0:
            # one window that survives on a single tiny texel (the predicate is sum != 0, not a threshold)
            ys, xs = 0, 0
            a_s = torch.roll(a, (-cfg["shift"], -cfg["shift"]), (2, 3))
            a_s[0, 0, ys:ys + ws, xs:xs + ws] = 0
            a_s[0, 0, ys + 1, xs + 1] = 1e-30
            a = torch.roll(a_s, (cfg["shift"], cfg["shift"]), (2, 3))
        p["alpha"] = a
    else:
        p["alpha"] = None
    return p


WRAPPER_CASES = {
    # Win_noShift_Attention(dim, heads, ws, shift) on (B, dim, H, W) with a blob alpha
    "wrap_c32_h4_ws4_s2": dict(C=32, heads=4, ws=4, shift=2, B=2, H=8, W=12, drop=0.3, seed=41),
    "wrap_c80_h8_ws4_s2": dict(C=80, heads=8, ws=4, shift=2, B=1, H=8, W=16, drop=0.4, seed=42),
}


def wrapper_inputs(cfg):
    """returns dict(x, alpha, state): `state` = a reference-style state dict with seeded values"""
    g = _gen(cfg["seed"])
    C, h, ws = cfg["C"], cfg["heads"], cfg["ws"]
    x = torch.randn(cfg["B"], C, cfg["H"], cfg["W"], generator=g)
    state = {}
    for br in ("conv_a", "conv_b"):
        for i in range(3):
            state[f"{br}.{i}.conv.0.weight"] = torch.randn(C // 2, C, 1, 1, generator=g) * C ** -0.5
            state[f"{br}.{i}.conv.0.bias"] = torch.randn(C // 2, generator=g) * 0.1
            state[f"{br}.{i}.conv.2.weight"] = torch.randn(C // 2, C // 2, 3, 3, generator=g) * (9 * C / 2) ** -0.5
            state[f"{br}.{i}.conv.2.bias"] = torch.randn(C // 2, generator=g) * 0.1
            state[f"{br}.{i}.conv.4.weight"] = torch.randn(C, C // 2, 1, 1, generator=g) * (C / 2) ** -0.5
            state[f"{br}.{i}.conv.4.bias"] = torch.randn(C, generator=g) * 0.1
    state["conv_b.3.weight"] = torch.randn(C, C, 1, 1, generator=g) * C ** -0.5
    state["conv_b.3.bias"] = torch.randn(C, generator=g) * 0.1
    s = C ** -0.5
    state["attn.attn.qkv.weight"] = torch.randn(3 * C, C, generator=g) * s
    state["attn.attn.qkv.bias"] = torch.randn(3 * C, generator=g) * 0.1
    state["attn.attn.proj.weight"] = torch.randn(C, C, generator=g) * s
    state["attn.attn.proj.bias"] = torch.randn(C, generator=g) * 0.1
    state["attn.attn.relative_position_bias_table"] = torch.randn((2 * ws - 1) ** 2, h, generator=g) * 0.3
    alpha = blob_alpha(cfg["B"], cfg["H"], cfg["W"], ws, cfg["shift"], cfg["drop"], cfg["seed"] + 1000)
    return dict(x=x, alpha=alpha, state=state)


PYRAMID_CASES = {
    # SupplyMaskToTransform on (B, 1, H, W): even sizes, odd sizes (windows hang over the right / bottom padding), tiny
    "pyr_2x64x96": dict(B=2, H=64, W=96, seed=51),
    "pyr_1x37x53": dict(B=1, H=37, W=53, seed=52),
    "pyr_3x5x7": dict(B=3, H=5, W=7, seed=53),
    "pyr_1x130x70": dict(B=1, H=130, W=70, seed=54),
}


def constraint_inputs(cfg):
    """binary masks with isolated holes / specks (also on the border), and a soft k/255 variant + the raw decoder output
    (values outside [0, 1]) that trainRGB.py:285-286 clamps and quantises first"""
    g = _gen(cfg["seed"] + 7)
    B, H, W = cfg["B"], cfg["H"], cfg["W"]
    binary = (torch.rand(B, 1, H, W, generator=g) < 0.5).float()
    binary[:, :, : H // 2, : W // 2] = 1.0
    binary[:, :, H // 2:, W // 2:] = 0.0
    holes = torch.rand(B, 1, H, W, generator=g) < 0.05
    binary = torch.where(holes, 1.0 - binary, binary)
    raw = binary + 0.3 * torch.randn(B, 1, H, W, generator=g) * (torch.rand(B, 1, H, W, generator=g) < 0.3)
    return dict(binary=binary, raw=raw)


def pyramid_inputs(cfg):
    """alpha: k/255 values with all-zero and all-one regions (what the dataset feeds); raw: un-quantised values in
    [0, 1] (what the mask decoder emits before models/AutoEncoderRGB_Journal.py:212-214 rounds them)"""
    g = _gen(cfg["seed"])
    B, H, W = cfg["B"], cfg["H"], cfg["W"]
    raw = torch.rand(B, 1, H, W, generator=g)
    raw[:, :, : H // 3, : W // 2] = 0.0
    raw[:, :, H - H // 4:, W - W // 3:] = 1.0
    alpha = torch.round(raw * 255) / 255
    return dict(alpha=alpha, raw=raw)


GDN_CASES = {
    "gdn_c192": dict(C=192, B=2, H=8, W=12, inverse=False, seed=21),
    "igdn_c192": dict(C=192, B=2, H=8, W=12, inverse=True, seed=22),
    "gdn_c16_bounds": dict(C=16, B=1, H=5, W=7, inverse=False, seed=23, hit_bounds=True),
    "igdn_c40_5d": dict(C=40, B=1, H=6, W=4, D=3, inverse=True, seed=24),
}


def gdn_inputs(cfg):
    g = _gen(cfg["seed"])
    C = cfg["C"]
    shape = (cfg["B"], C, cfg["D"], cfg["H"], cfg["W"]) if "D" in cfg else (cfg["B"], C, cfg["H"], cfg["W"])
    x = torch.randn(*shape, generator=g) * 2.0
    pedestal = (2 ** -18) ** 2
    beta = torch.sqrt(0.5 + torch.rand(C, generator=g) * 2 + pedestal)
    gamma = torch.sqrt(0.1 * torch.eye(C) + torch.rand(C, C, generator=g) * 0.02 + pedestal)
    if cfg.get("hit_bounds"):
        beta[::3] = 1e-5            # below beta_bound  -> clamped
        gamma[::2, 1::2] = 0.0      # below gamma_bound -> clamped
        gamma[1, 2] = -0.3
    return dict(x=x, beta=beta, gamma=gamma)


def rounding_inputs():
    g = _gen(31)
    half = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5, 1e-8, -1e-8, 0.0, -0.0, 8388607.5, 1e10, -7.49999, 7.5])
    x = torch.cat([half, torch.randn(4081, generator=g) * 6])
    mu = torch.randn(x.numel(), generator=g) * 2
    lrp = torch.randn(x.numel(), generator=g) * 1.5
    m = torch.rand(x.numel(), generator=g)
    return dict(x=x, mu=mu, lrp=lrp, m=m)


# ------------------------------------------------------------------------------------------------ whole-model cases
MODEL_CASES = {
    # AutoEncoderRGB_Journal.AutoEncoder forward on (B, 3, H, W) + alpha; H, W multiples of 64 (window sizes x strides x
    # hyperprior strides) and > 160 (the 4 downsamplings of MS-SSIM)
    "rgb_1x192x256": dict(B=1, H=192, W=256, drop=0.35, seed=61),
}


def model_inputs(cfg):
    """image in [0, 1], alpha = k/255 blobs with ~drop of the 32x32 blocks fully transparent (what the dataset feeds,
    trainRGB.py:275-279: the image is pre-multiplied by its binarised alpha), reconmask = a slightly different alpha (what
    the mask codec hands over: clamp, quantise, `constraint`)."""
    g = _gen(cfg["seed"])
    B, H, W = cfg["B"], cfg["H"], cfg["W"]
    yy = torch.linspace(0, 1, H).view(1, 1, H, 1)
    xx = torch.linspace(0, 1, W).view(1, 1, 1, W)
    base = torch.cat([0.5 + 0.4 * torch.sin(6.0 * xx + 3.0 * yy), 0.5 + 0.4 * torch.cos(5.0 * yy - 2.0 * xx),
                      0.5 + 0.3 * torch.sin(9.0 * xx * yy + 1.0)], dim=1).expand(B, 3, H, W)
    image = (base + 0.08 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    keep = (torch.rand(B, 1, H // 32, W // 32, generator=g) >= cfg["drop"]).float()
    alpha = keep.repeat_interleave(32, 2).repeat_interleave(32, 3)
    soft = torch.round((0.3 + 0.7 * torch.rand(B, 1, H, W, generator=g)) * 255) / 255
    alpha = alpha * soft
    image = image * (alpha > 0).float()
    flip = (torch.rand(B, 1, H, W, generator=g) < 0.002).float()
    reconmask = (alpha * (1 - flip) + flip * 0.5 * (alpha == 0).float()).clamp(0, 1)
    return dict(image=image, alpha=alpha, reconmask=reconmask)


def model_state(table, seed):
    """Deterministic weights BY KEY NAME for a model whose state dict has the entries of `table` ({key: shape}; the
    committed tests/golden/model_rgb_keys.json lists the reference model's).  Independent of construction order, so
    the reference model (build container), the oracle restatement and the B200 codec all load the very same values.
    Scales keep the activations O(1) through the stack; integer buffers (relative_position_index) are skipped."""
    pedestal = (2.0 ** -18) ** 2
    out = {}
    for key in sorted(table):
        shape = tuple(table[key])
        if key.endswith("relative_position_index"):
            continue
        g = _gen((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7fffffff)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "beta":
            v = torch.sqrt(0.6 + 0.8 * torch.rand(shape, generator=g) + pedestal)
        elif leaf == "gamma":
            v = torch.sqrt(0.1 * torch.eye(shape[0]) + 0.01 * torch.rand(shape, generator=g) + pedestal)
        elif leaf == "relative_position_bias_table":
            v = torch.randn(shape, generator=g) * 0.2
        elif leaf == "quantiles":
            v = torch.tensor([-10.0, 0.0, 10.0]).repeat(shape[0], 1, 1)
            v[:, 0, 1] = torch.randn(shape[0], generator=g) * 0.3
        elif len(shape) >= 2:
            # Conv2d (out, in, kh, kw) / ConvTranspose2d (in, out, kh, kw) / Linear (out, in)
            fan = shape[1] * (shape[2] * shape[3] if len(shape) == 4 else 1)
            if len(shape) == 4 and key.startswith("Decoder.x") and shape[2] == 5:
                fan = shape[0] * 25 / 4                   # transposed, stride 2: a quarter of the taps per output
            v = torch.randn(shape, generator=g) * (1.0 / fan) ** 0.5
            if key == "Encoder.x4.weight":
                v = v * 6.0                               # latents that span several quantisation steps ...
            if key == "Decoder.x1.weight":
                v = v * 0.06                              # ... without the three IGDNs blowing the decoder up
        else:
            v = torch.randn(shape, generator=g) * 0.05
        out[key] = v
    return out
