"""CPU restatement (numpy, float64) of the rate / distortion tail of the codec's forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  What it restates:
  * reconstruct_error                        models/AutoEncoderRGB_Journal.py:36-64
  * the bpp terms of AutoEncoder.forward     models/AutoEncoderRGB_Journal.py:283-291
  * the two likelihood models behind them, which the reference takes from its third-party dependency CompressAI
    (`compressai` in the reference's requirements; NOT in /root/reference and not installed here -- oracle/shims/compressai
    is a stand-in for imports only).  PARITY UNPINNED at this boundary: the formulas below follow CompressAI's published
    algorithm (Balle et al. 2018, appendix 6.1, as implemented by compressai.entropy_models.EntropyBottleneck
    ._logits_cumulative / ._likelihood with filters (3, 3, 3, 3), likelihood_bound 1e-9, and GaussianConditional
    ._likelihood with scale_bound 0.11), anchored on the reference's own call sites (:226, :274-276, :283-291).
tests/test_oracle_rate.py holds the package's torch expressions (codec.py) to this restatement on the CPU; the GPU test
(tests/test_gpu_model.py) holds rate_forward (csrc/rate.cu) to those expressions.
"""
from __future__ import annotations

import math

import numpy as np


def _softplus(x):
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def _erfc(x):
    return np.vectorize(math.erfc, otypes=[np.float64])(x)


def bits(lik):
    """sum of clamp(-log2(lik + 1e-10), 0, 50)   (:283-287)"""
    return float(np.clip(-np.log(lik + 1e-10) / math.log(2.0), 0.0, 50.0).sum())


def masked_mse(x, x_hat, mask):
    """(:36-64) squared error over the pixels whose alpha is > 0, per image over C * count, mean over the batch"""
    x, x_hat, mask = (np.asarray(t, dtype=np.float64) for t in (x, x_hat, mask))
    m = (mask > 0.0).astype(np.float64)                          # (B, 1, H, W)
    se = (((x - x_hat) * m) ** 2).sum(axis=(1, 2, 3))
    cnt = np.maximum(m.sum(axis=(1, 2, 3)) * x.shape[1], 1.0)
    return float((se / cnt).mean())


def gaussian_likelihood(y, scales, means):
    """GaussianConditional._likelihood: mass of N(mu, max(s, 0.11)) on [y - 0.5, y + 0.5], floored at 1e-9"""
    y, scales, means = (np.asarray(t, dtype=np.float64) for t in (y, scales, means))
    v = np.abs(y - means)
    s = np.maximum(scales, 0.11)
    c = 2.0 ** -0.5
    upper = 0.5 * _erfc(-c * (0.5 - v) / s)
    lower = 0.5 * _erfc(-c * (-0.5 - v) / s)
    return np.maximum(upper - lower, 1e-9)


def logits_cumulative(v, params):
    """EntropyBottleneck._logits_cumulative; v (C, 1, N); params: dict of _matrix{i} (C, f_{i+1}, f_i), _bias{i}, _factor{i}"""
    n = sum(1 for k in params if k.startswith("_matrix"))
    for i in range(n):
        m = _softplus(np.asarray(params[f"_matrix{i}"], dtype=np.float64))
        v = np.einsum("crk,ckn->crn", m, v) + np.asarray(params[f"_bias{i}"], dtype=np.float64)
        if i < n - 1:
            v = v + np.tanh(np.asarray(params[f"_factor{i}"], dtype=np.float64)) * np.tanh(v)
    return v


def factorised_likelihood(z_hat, params):
    """EntropyBottleneck._likelihood on z_hat (B, C, h, w): |sigmoid(s up) - sigmoid(s lo)|, s = -sign(lo + up), floored at 1e-9"""
    z = np.asarray(z_hat, dtype=np.float64)
    B, C = z.shape[:2]
    v = z.transpose(1, 0, 2, 3).reshape(C, 1, -1)
    lo, up = logits_cumulative(v - 0.5, params), logits_cumulative(v + 0.5, params)
    sg = -np.sign(lo + up)
    lik = np.maximum(np.abs(_sigmoid(sg * up) - _sigmoid(sg * lo)), 1e-9)
    return lik.reshape(C, B, *z.shape[2:]).transpose(1, 0, 2, 3)


def rate_terms(x, x_hat, mask, y, scales, means, z_hat, eb_params):
    """[mse, y bpp, z bpp, total bpp] of AutoEncoder.forward (:283-297); bpp = bits / (B * H * W)"""
    px = x.shape[0] * x.shape[2] * x.shape[3]
    yb = bits(gaussian_likelihood(y, scales, means)) / px
    zb = bits(factorised_likelihood(z_hat, eb_params)) / px
    return masked_mse(x, x_hat, mask), yb, zb, yb + zb
