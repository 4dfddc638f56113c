"""Masked MS-SSIM / PSNR on the B200 (csrc/msssim.cu) against the CPU oracle (oracle/ref_model.py, pinned on the values of
the reference's metrics/masked_ms_ssim_torch.ms_ssim committed in tests/golden/model_rgb.npz) -- "masked MS-SSIM / PSNR equal
to 3 decimals" is BASELINE.json's model-level criterion; here the device metric itself is held to 1e-5."""
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_model as M

pytestmark = pytest.mark.gpu


def _pair(B, H, W, seed, noise=0.05):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(B, 3, H, W, generator=g)
    Y = (X + noise * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    mask = torch.zeros(B, 1, H, W)
    for b in range(B):
        cy, cx, r = (torch.rand(3, generator=g) * torch.tensor([H, W, min(H, W) / 3.0]) + torch.tensor([0.0, 0.0, min(H, W) / 4.0])).tolist()
        d = r - torch.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
        mask[b, 0] = torch.round(torch.clamp(d / 3.0, 0, 1) * 255) / 255
    return X, Y, mask


@pytest.mark.parametrize("shape", [(2, 192, 256), (1, 161, 203), (2, 176, 331)], ids=lambda s: "x".join(map(str, s)))
def test_masked_ms_ssim_matches_oracle(pkg, cuda_dev, shape):
    B, H, W = shape
    X, Y, mask = _pair(B, H, W, seed=H * 7 + W)
    want = float(M.masked_ms_ssim(X, Y, mask))
    got = float(pkg.masked_ms_ssim(X.to(cuda_dev), Y.to(cuda_dev), mask.to(cuda_dev)))
    assert abs(got - want) < 1e-5, (got, want)
    assert abs(float(pkg.masked_psnr(X.to(cuda_dev), Y.to(cuda_dev), mask.to(cuda_dev))) - M.psnr(M.masked_mse(X, Y, mask))) < 1e-3


def test_masked_ms_ssim_properties(pkg, cuda_dev):
    X, Y, mask = _pair(2, 192, 256, seed=11)
    Xd, Yd, md = X.to(cuda_dev), Y.to(cuda_dev), mask.to(cuda_dev)
    assert abs(float(pkg.masked_ms_ssim(Xd, Xd, md)) - 1.0) < 1e-6                 # identical images
    a, b = float(pkg.masked_ms_ssim(Xd, Yd, md)), float(pkg.masked_ms_ssim(Yd, Xd, md))
    assert abs(a - b) < 1e-6                                                         # symmetric
    # pixels outside the mask do not matter
    Y2 = torch.where(md > 0, Yd, torch.rand_like(Yd))
    assert abs(float(pkg.masked_ms_ssim(Xd, Y2, md)) - a) < 1e-6
    with pytest.raises(AssertionError):
        pkg.masked_ms_ssim(Xd[:, :, :160], Yd[:, :, :160], md[:, :, :160])


def test_codec_metrics_on_device_match_golden(pkg, cuda_dev, golden, model_keys):
    """the model-level criterion evaluated entirely on the device: forward on the B200, masked PSNR and masked MS-SSIM of
    the clipped reconstruction against the values of the unmodified reference"""
    name = next(iter(G.MODEL_CASES))
    cfg, g = G.MODEL_CASES[name], golden["model_rgb"]
    p = G.model_inputs(cfg)
    net = pkg.RGBACodec().eval()
    net.load_state_dict(G.model_state(model_keys, cfg["seed"]), strict=False)
    net = net.to(cuda_dev)
    image, alpha, recon = (p[k].to(cuda_dev) for k in ("image", "alpha", "reconmask"))
    with torch.no_grad():
        x_hat = net.detail(image, alpha, recon)["x_hat"].clamp(0, 1)
        ms = float(pkg.masked_ms_ssim(image, x_hat, alpha))
        ps = float(pkg.masked_psnr(image, x_hat, alpha))
    assert round(ms, 3) == round(float(g[name + "/ms_ssim"]), 3), (ms, float(g[name + "/ms_ssim"]))
    assert round(ps, 3) == round(M.psnr(float(g[name + "/mse_clipped"])), 3)
