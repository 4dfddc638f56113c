"""The rate / distortion tail of the forward: the package's torch expressions (codec.py: reconstruct_error,
GaussianConditional.likelihood, EntropyBottleneck.likelihood, _bits -- what `rate_forward` is held to on the GPU) against
the float64 restatement of oracle/ref_rate.py (models/AutoEncoderRGB_Journal.py:36-64, :283-291; CompressAI's published
likelihood models: parity unpinned at that third-party boundary, see the oracle's header).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_rate as RR


def _bottleneck(pkg, C, seed):
    torch.manual_seed(seed)
    eb = pkg.codec.EntropyBottleneck(C)
    with torch.no_grad():
        for n, p in eb.named_parameters():
            if n != "quantiles":
                p.add_(torch.randn_like(p) * 0.3)
    return eb


@pytest.mark.parametrize("C,B,h,w", [(8, 2, 3, 5), (192, 1, 4, 6)])
def test_factorised_prior_matches_float64_restatement(pkg, C, B, h, w):
    eb = _bottleneck(pkg, C, 3 + C)
    z_hat = torch.round(torch.randn(B, C, h, w) * 4.0)
    with torch.no_grad():
        lik = eb.likelihood(z_hat)
    want = RR.factorised_likelihood(z_hat.numpy(), {n: p.detach().numpy() for n, p in eb.named_parameters()})
    assert lik.shape == z_hat.shape
    np.testing.assert_allclose(lik.numpy(), want, rtol=2e-4, atol=2e-8)
    assert float(lik.min()) >= 1e-9


def test_gaussian_conditional_matches_float64_restatement(pkg):
    g = torch.Generator().manual_seed(9)
    means = torch.randn(2, 16, 6, 7, generator=g) * 3.0
    scales = torch.rand(2, 16, 6, 7, generator=g) * 4.0               # some below the 0.11 bound
    scales[0, 0] = 0.01
    y = means + torch.randn(2, 16, 6, 7, generator=g) * scales.clamp_min(0.11) * 1.5
    y[1, 3] += 40.0                                                   # far tail: the 1e-9 floor
    lik = pkg.codec.GaussianConditional.likelihood(y, scales, means)
    want = RR.gaussian_likelihood(y.numpy(), scales.numpy(), means.numpy())
    np.testing.assert_allclose(lik.numpy(), want, rtol=2e-4, atol=3e-8)
    assert float(lik.min()) == pytest.approx(1e-9)


def test_rate_terms_match_float64_restatement(pkg):
    g = torch.Generator().manual_seed(21)
    B, H, W = 2, 32, 48
    x = torch.rand(B, 3, H, W, generator=g)
    x_hat = (x + 0.05 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1)
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.4).float() * torch.rand(B, 1, H, W, generator=g)
    mask[1] = 0.0                                                     # fully transparent image: count clamps to 1
    means = torch.randn(B, 20, 4, 6, generator=g) * 2.0
    scales = torch.rand(B, 20, 4, 6, generator=g) * 2.0
    y = means + torch.randn(B, 20, 4, 6, generator=g)
    eb = _bottleneck(pkg, 12, 5)
    z_hat = torch.round(torch.randn(B, 12, 2, 3, generator=g) * 3.0)
    with torch.no_grad():
        mse = pkg.codec.reconstruct_error(x, x_hat, mask)
        px = B * H * W
        yb = pkg.codec._bits(pkg.codec.GaussianConditional.likelihood(y, scales, means)) / px
        zb = pkg.codec._bits(eb.likelihood(z_hat)) / px
    want = RR.rate_terms(x.numpy(), x_hat.numpy(), mask.numpy(), y.numpy(), scales.numpy(), means.numpy(), z_hat.numpy(),
                         {n: p.detach().numpy() for n, p in eb.named_parameters()})
    for got, ref, what in zip((mse, yb, zb, yb + zb), want, ("mse", "y bpp", "z bpp", "bpp")):
        assert float(got) == pytest.approx(ref, rel=2e-5), what
