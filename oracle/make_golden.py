"""Regenerate tests/golden/*.npz from the LIVE reference (build container only).

    python -m oracle.make_golden

Runs the unmodified reference modules (through oracle/live_reference.py) on the seeded recipes of
oracle/golden_cases.py, in fp32 on CPU, and stores their forward outputs and autograd gradients.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import golden_cases as G
from . import live_reference

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def make_attention(ref):
    out = {}
    for name, cfg in G.ATTENTION_CASES.items():
        p = G.attention_inputs(cfg)
        cls = ref.masked.WinBasedAttention if cfg["masked"] else ref.unmasked.WinBasedAttention
        m = cls(dim=cfg["C"], num_heads=cfg["heads"], window_size=cfg["ws"], shift_size=cfg["shift"],
                qkv_bias=cfg.get("qkv_bias", True))
        with torch.no_grad():
            m.attn.qkv.weight.copy_(p["qkv_w"])
            if p["qkv_b"] is not None:
                m.attn.qkv.bias.copy_(p["qkv_b"])
            m.attn.proj.weight.copy_(p["proj_w"])
            m.attn.proj.bias.copy_(p["proj_b"])
            m.attn.relative_position_bias_table.copy_(p["table"])
        x = p["x"].clone().requires_grad_(True)
        y = m(x, p["alpha"]) if cfg["masked"] else m(x)
        gsel = torch.Generator().manual_seed(cfg["seed"] + 5)
        gy = torch.randn(y.shape, generator=gsel)
        y.backward(gy)
        out[name + "/y"] = _np(y)
        out[name + "/dx"] = _np(x.grad)
        if cfg["C"] <= 80:                      # keep the fixtures small: weight gradients for the small cases only
            out[name + "/dqkv_w"] = _np(m.attn.qkv.weight.grad)
            out[name + "/dproj_w"] = _np(m.attn.proj.weight.grad)
        out[name + "/dtable"] = _np(m.attn.relative_position_bias_table.grad)
        out[name + "/crc"] = np.array(G.checksum(*[p[k] for k in ("x", "alpha", "qkv_w", "qkv_b", "proj_w", "proj_b",
                                                                 "table")]), dtype=np.int64)
    return out


def make_gdn(ref):
    out = {}
    for name, cfg in G.GDN_CASES.items():
        p = G.gdn_inputs(cfg)
        m = ref.gdn.GDN(cfg["C"], inverse=cfg["inverse"])
        with torch.no_grad():
            m.beta.copy_(p["beta"])
            m.gamma.copy_(p["gamma"])
        x = p["x"].clone().requires_grad_(True)
        y = m(x)
        gsel = torch.Generator().manual_seed(cfg["seed"] + 5)
        gy = torch.randn(y.shape, generator=gsel)
        y.backward(gy)
        out[name + "/y"] = _np(y)
        out[name + "/dx"] = _np(x.grad)
        out[name + "/dbeta"] = _np(m.beta.grad)
        out[name + "/dgamma"] = _np(m.gamma.grad)
        out[name + "/crc"] = np.array(G.checksum(p["x"], p["beta"], p["gamma"]), dtype=np.int64)
    return out


def make_wrapper(ref):
    out = {}
    for name, cfg in G.WRAPPER_CASES.items():
        p = G.wrapper_inputs(cfg)
        m = ref.wrapper.Win_noShift_Attention(cfg["C"], num_heads=cfg["heads"], window_size=cfg["ws"],
                                              shift_size=cfg["shift"])
        missing, unexpected = m.load_state_dict(p["state"], strict=False)
        assert not unexpected and all(k.endswith("relative_position_index") for k in missing), (missing, unexpected)
        x = p["x"].clone().requires_grad_(True)
        y = m(x, p["alpha"])
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(cfg["seed"] + 5))
        y.backward(gy)
        out[name + "/y"] = _np(y)
        out[name + "/dx"] = _np(x.grad)
        out[name + "/dconv_b3_w"] = _np(m.conv_b[3].weight.grad)
        out[name + "/dconv_a0_w"] = _np(m.conv_a[0].conv[0].weight.grad)
        out[name + "/keys"] = np.array(sorted(m.state_dict().keys()))
        out[name + "/crc"] = np.array(G.checksum(p["x"], p["alpha"], *[p["state"][k] for k in sorted(p["state"])]),
                                      dtype=np.int64)
    return out


def make_pyramid(ref):
    out = {}
    sm = ref.supply.SupplyMaskToTransform()
    for name, cfg in G.PYRAMID_CASES.items():
        p = G.pyramid_inputs(cfg)
        for k, lvl in enumerate(sm(p["alpha"]), 1):
            out[f"{name}/mask{k}"] = _np(lvl)
        # decoder side: models/AutoEncoderRGB_Journal.py:212-215
        recon = p["raw"] * 255
        recon = torch.round(recon)
        recon = recon / 255
        out[name + "/recon"] = _np(recon)
        for k, lvl in enumerate(sm(recon), 1):
            out[f"{name}/md{k}"] = _np(lvl)
        out[name + "/crc"] = np.array(G.checksum(p["alpha"], p["raw"]), dtype=np.int64)
        # constraint (trainRGB.py:98-111) on a binary mask, and the evaluation chain of trainRGB.py:285-287 on raw values
        script = ref.script("trainRGB")
        c = G.constraint_inputs(cfg)
        out[name + "/constraint_binary"] = _np(script.constraint(c["binary"].clone()))
        m = torch.clamp(c["raw"], 0, 1)
        m = torch.round(m * 255) / 255
        out[name + "/constraint_chain"] = _np(script.constraint(m))
        out[name + "/crc_c"] = np.array(G.checksum(c["binary"], c["raw"]), dtype=np.int64)
    return out


def make_rounding(ref):
    p = G.rounding_inputs()
    rgb = ref.model("rgb")
    out = {
        "ste_round": _np(rgb.ste_round(p["x"])),
        "quantize_offset": _np(rgb.ste_round(p["x"] - p["mu"]) + p["mu"]),
        "lrp_add": _np(p["x"] + 0.5 * torch.tanh(p["lrp"])),
        "levels255": _np(torch.round(p["m"] * 255) / 255),
        "crc": np.array(G.checksum(p["x"], p["mu"], p["lrp"], p["m"]), dtype=np.int64),
    }
    return out


def make_model(ref):
    """whole-model forward of the UNMODIFIED AutoEncoderRGB_Journal.AutoEncoder (weights by key name, G.model_state) +
    the reference's own masked MS-SSIM / PSNR on its output; also (re)writes the key table the weight recipe needs."""
    import json
    rgb = ref.model("rgb")
    import importlib
    metric = importlib.import_module("metrics.masked_ms_ssim_torch")      # the masked MS-SSIM of the reference
    out = {}
    for name, cfg in G.MODEL_CASES.items():
        torch.manual_seed(234)
        net = rgb.AutoEncoder().eval()
        table = {k: list(v.shape) for k, v in net.state_dict().items()}
        with open(os.path.join(OUT_DIR, "model_rgb_keys.json"), "w") as f:
            json.dump(table, f, indent=0, sort_keys=True)
        missing = net.load_state_dict(G.model_state(table, cfg["seed"]), strict=False)
        assert not missing.unexpected_keys and all(k.endswith("relative_position_index") for k in missing.missing_keys)
        p = G.model_inputs(cfg)
        cap = {}
        net.Encoder.register_forward_hook(lambda m, i, o: cap.__setitem__("y", o))
        net.h_a.register_forward_hook(lambda m, i, o: cap.__setitem__("z", o))
        net.Decoder.register_forward_pre_hook(lambda m, i: cap.__setitem__("y_hat", i[0]))
        with torch.no_grad():
            me = net.EncMakeMask(p["alpha"])
            x_hat, mse, bpp, bpp_y, bpp_z = net(p["image"], p["alpha"], p["reconmask"], me[0], me[1], me[2], me[3])
            clipped = torch.clamp(x_hat, 0, 1)
            msssim = metric.ms_ssim(p["image"], clipped, p["alpha"], data_range=1.0, size_average=True)
            mse_clipped = rgb.reconstruct_error(p["image"], clipped, p["alpha"], p["reconmask"])
        for k in ("y", "z", "y_hat"):
            out[f"{name}/{k}"] = _np(cap[k])
        out[f"{name}/x_hat"] = _np(x_hat)
        out[f"{name}/mse"] = _np(mse)
        out[f"{name}/mse_clipped"] = _np(mse_clipped)
        out[f"{name}/ms_ssim"] = _np(msssim)
        out[f"{name}/crc"] = np.array(G.checksum(p["image"], p["alpha"], p["reconmask"]), dtype=np.int64)
        print(f"{name}: mse {float(mse):.6f} psnr {10 * np.log10(1 / float(mse_clipped)):.4f} ms-ssim {float(msssim):.6f} "
              f"|y| {float(cap['y'].abs().mean()):.3f} |x_hat| {float(x_hat.abs().mean()):.3f}")
    return out


def main():
    ref = live_reference.load()
    torch.set_num_threads(1)           # fixed reduction order
    os.makedirs(OUT_DIR, exist_ok=True)
    import sys
    only = set(sys.argv[1:])
    makers = (("attention.npz", make_attention), ("gdn.npz", make_gdn), ("rounding.npz", make_rounding),
              ("wrapper.npz", make_wrapper), ("pyramid.npz", make_pyramid), ("model_rgb.npz", make_model))
    for fname, data in ((f, mk(ref)) for f, mk in makers if not only or f in only):
        path = os.path.join(OUT_DIR, fname)
        np.savez_compressed(path, **data)
        print(f"wrote {path}: {len(data)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
