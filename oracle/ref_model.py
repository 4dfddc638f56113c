"""CPU restatement of the RGBA codec's encode + decode forward and of its evaluation metrics.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
It exists because the reference's Python cannot travel to the GPU box; it is pinned on outputs of the UNMODIFIED
reference model run in the build container (oracle/make_golden.py -> tests/golden/model_rgb.npz, checked by
tests/test_oracle_golden.py and, live, by tests/test_oracle_vs_reference.py).

Everything is a pure function of a reference-style state dict `w` (the key names of
models/AutoEncoderRGB_Journal.py's AutoEncoder).  Citations are into /root/reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ref_ops as R


def _sub(w, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in w.items() if k.startswith(prefix)}


def _conv(x, w, name, stride=1, padding=0):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=stride, padding=padding)


def _deconv5(x, w, name):
    """nn.ConvTranspose2d(k=5, stride=2, padding=2, output_padding=1)   (layers/TransformRGB.py:83, 86, 88)"""
    return F.conv_transpose2d(x, w[name + ".weight"], w[name + ".bias"], stride=2, padding=2, output_padding=1)


# --------------------------------------------------------------------------------------------------- transforms
def analysis(w, x, me2, me3):
    """Analysis_transform.forward   (layers/TransformRGB.py:52-75); `w` = the `Encoder.` sub-dict."""
    y = R.gdn(_conv(x, w, "x1", 2, 2), w["gdn1.beta"], w["gdn1.gamma"])
    y = R.gdn(_conv(y, w, "x2", 2, 2), w["gdn2.beta"], w["gdn2.gamma"])
    y = R.win_noshift_attention(y, me2, _sub(w, "attention1."), 8, 8, 4)
    y = R.gdn(_conv(y, w, "x3", 2, 2), w["gdn3.beta"], w["gdn3.gamma"])
    y = _conv(y, w, "x4")
    return R.win_noshift_attention(y, me3, _sub(w, "attention2."), 8, 4, 2)


def dse(w, x):
    """DSE.forward   (layers/TransformRGB.py:30-49)"""
    first = _conv(x, w, "input_conv")
    t = first
    for name in ("enh1", "enh2", "enh3"):
        u = _conv(F.relu(_conv(t, w, name + ".conv1", padding=1)), w, name + ".conv2", padding=1)
        t = u + t
    return _conv(t + first, w, "output_conv") + x


def synthesis(w, y_hat, md2, md3):
    """Synthesis_transform.forward   (layers/TransformRGB.py:77-100); `w` = the `Decoder.` sub-dict."""
    y = R.win_noshift_attention(y_hat, md3, _sub(w, "attention1."), 8, 4, 2)
    y = R.gdn(_conv(y, w, "x1"), w["igdn1.beta"], w["igdn1.gamma"], inverse=True)
    y = R.gdn(_deconv5(y, w, "x2"), w["igdn2.beta"], w["igdn2.gamma"], inverse=True)
    y = R.win_noshift_attention(y, md2, _sub(w, "attention2."), 8, 8, 4)
    y = R.gdn(_deconv5(y, w, "x3"), w["igdn3.beta"], w["igdn3.gamma"], inverse=True)
    return dse(_sub(w, "dse."), _deconv5(y, w, "x4"))


def hyper_analysis(w, y):
    """h_a   (models/AutoEncoderRGB_Journal.py:139-149): conv3x3 s2, GELU, conv3x3, GELU, s2, GELU, conv3x3, GELU, s2"""
    t = y
    for i, stride in ((0, 2), (2, 1), (4, 2), (6, 1), (8, 2)):
        t = _conv(t, w, str(i), stride, 1)
        if i < 8:
            t = F.gelu(t)
    return t


def hyper_synthesis(w, z_hat):
    """h_mean_s / h_scale_s   (:151-173): subpel, GELU, conv3x3, GELU, subpel, GELU, conv3x3, GELU, subpel"""
    t = F.gelu(F.pixel_shuffle(_conv(z_hat, w, "0.0", padding=1), 2))
    t = F.gelu(_conv(t, w, "2", padding=1))
    t = F.gelu(F.pixel_shuffle(_conv(t, w, "4.0", padding=1), 2))
    t = F.gelu(_conv(t, w, "6", padding=1))
    return F.pixel_shuffle(_conv(t, w, "8.0", padding=1), 2)


def _cc(w, x):
    """one channel-conditional stack (:175-203): conv3x3, GELU, conv3x3, GELU, conv3x3"""
    t = F.gelu(_conv(x, w, "0", padding=1))
    t = F.gelu(_conv(t, w, "2", padding=1))
    return _conv(t, w, "4", padding=1)


def slice_loop(w, y, latent_means, latent_scales, num_slices=10, max_support=5):
    """the channel-conditional slice loop (:240-266).  Returns y_hat, the means and the scales."""
    hh, ww = y.shape[2:]
    y_hat_slices, mus, scales = [], [], []
    for i, y_slice in enumerate(y.chunk(num_slices, 1)):
        support = y_hat_slices[:max_support]
        mean_support = torch.cat([latent_means] + support, dim=1)
        mu = _cc(_sub(w, f"cc_mean_transforms.{i}."), mean_support)[:, :, :hh, :ww]
        scale = _cc(_sub(w, f"cc_scale_transforms.{i}."), torch.cat([latent_scales] + support, dim=1))[:, :, :hh, :ww]
        y_hat = R.quantize_offset(y_slice, mu)
        lrp = _cc(_sub(w, f"lrp_transforms.{i}."), torch.cat([mean_support, y_hat], dim=1))
        y_hat = R.lrp_add(y_hat, lrp)
        y_hat_slices.append(y_hat)
        mus.append(mu)
        scales.append(scale)
    return torch.cat(y_hat_slices, 1), torch.cat(mus, 1), torch.cat(scales, 1)


def masked_mse(x, x_hat, mask):
    """reconstruct_error   (models/AutoEncoderRGB_Journal.py:36-64): squared error on the pixels whose input alpha is
    > 0, per image divided by the number of such (pixel, channel) entries (at least 1), then the batch mean."""
    m = (mask.expand(-1, 3, -1, -1) > 0).to(x.dtype)
    se = ((x * m - x_hat * m) ** 2).sum(dim=(1, 2, 3))
    return (se / m.sum(dim=(1, 2, 3)).clamp(min=1)).mean()


def rgb_forward(w, image, mask, reconmask):
    """AutoEncoder.forward   (models/AutoEncoderRGB_Journal.py:203-297) incl. the caller's EncMakeMask (trainRGB.py:283).
    image (B,3,H,W), mask / reconmask (B,1,H,W).  Returns a dict of every tensor the parity tests look at.
    Likelihoods / bpp are not restated (CompressAI boundary)."""
    me = R.alpha_pyramid(mask)
    reconmask = R.quantize_levels(reconmask, 255)
    md = R.alpha_pyramid(reconmask)
    y = analysis(_sub(w, "Encoder."), image, me[1], me[2])
    z = hyper_analysis(_sub(w, "h_a."), y)
    med = w["entropy_bottleneck.quantiles"][:, :, 1:2].reshape(1, -1, 1, 1)
    z_hat = R.quantize_offset(z, med)
    latent_scales = hyper_synthesis(_sub(w, "h_scale_s."), z_hat)
    latent_means = hyper_synthesis(_sub(w, "h_mean_s."), z_hat)
    y_hat, mus, scales = slice_loop(w, y, latent_means, latent_scales)
    x_hat = synthesis(_sub(w, "Decoder."), y_hat, md[1], md[2])
    return dict(y=y, z=z, z_hat=z_hat, y_hat=y_hat, means=mus, scales=scales, x_hat=x_hat,
                mse=masked_mse(image, x_hat, mask))


def psnr(mse):
    """trainRGB.py:303"""
    return 10.0 * math.log10(1.0 / float(mse))


# --------------------------------------------------------------------------------------------------- masked MS-SSIM
def _gauss_1d(size=11, sigma=1.5):
    c = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _blur(x, g):
    """separable valid-mode Gaussian, rows then columns, a dimension shorter than the window is left alone
    (metrics/masked_ms_ssim_torch.py:27-55)"""
    C = x.shape[1]
    k = g.to(x.dtype).view(1, 1, -1).repeat(C, 1, 1)
    out = x
    if x.shape[2] >= g.numel():
        out = F.conv2d(out, k.unsqueeze(-1), groups=C)
    if x.shape[3] >= g.numel():
        out = F.conv2d(out, k.unsqueeze(-2), groups=C)
    return out


def _nearest_resize(m, oh, ow):
    """torchvision resize(..., NEAREST) = F.interpolate(mode='nearest'): source index floor(dst * in / out)"""
    ih, iw = m.shape[2:]
    ys = torch.clamp((torch.arange(oh, dtype=torch.float32) * (ih / oh)).floor().long(), max=ih - 1)
    xs = torch.clamp((torch.arange(ow, dtype=torch.float32) * (iw / ow)).floor().long(), max=iw - 1)
    return m[:, :, ys][:, :, :, xs]


def _masked_ssim_level(X, Y, mask, g, data_range, K=(0.01, 0.03)):
    """_ssim   (metrics/masked_ms_ssim_torch.py:58-121): SSIM / CS maps averaged over the positions where the
    nearest-resized mask is non-zero"""
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _blur(X, g), _blur(Y, g)
    s1 = _blur(X * X, g) - mu1 * mu1
    s2 = _blur(Y * Y, g) - mu2 * mu2
    s12 = _blur(X * Y, g) - mu1 * mu2
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs_map
    one = g.numel() - 1
    m = (_nearest_resize(mask, mask.shape[2] - one, mask.shape[3] - one) > 0).to(X.dtype)
    nz = m.flatten(2) != 0
    den = nz.sum(-1) + 1e-10
    return (ssim_map.flatten(2) * nz).sum(-1) / den, (cs_map.flatten(2) * nz).sum(-1) / den


def masked_ms_ssim(X, Y, mask, data_range=1.0, weights=(0.0448, 0.2856, 0.3001, 0.2363, 0.1333)):
    """ms_ssim(X, Y, mask, data_range, size_average=True)   (metrics/masked_ms_ssim_torch.py:181-265): five levels;
    at each one the mask is binarised and multiplied into both images, then everything is 2x2 average-pooled."""
    if min(X.shape[-2:]) <= 160:
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    g = _gauss_1d()
    wts = torch.tensor(weights, dtype=X.dtype)
    vals = []
    for lvl in range(len(weights)):
        mask = (mask > 0).to(X.dtype)
        X, Y = X * mask, Y * mask
        s, cs = _masked_ssim_level(X, Y, mask, g, data_range)
        if lvl < len(weights) - 1:
            vals.append(torch.relu(cs))
            pad = [d % 2 for d in X.shape[2:]]
            X, Y, mask = (F.avg_pool2d(t, 2, padding=pad) for t in (X, Y, mask))
    vals.append(torch.relu(s))
    stack = torch.stack(vals, 0)
    return torch.prod(stack ** wts.view(-1, 1, 1), dim=0).mean()
