"""Structural stand-ins (bpp NOT comparable with real CompressAI)."""
import torch
from torch import nn


class EntropyBottleneck(nn.Module):
    def __init__(self, channels, *a, **k):
        super().__init__()
        self.channels = channels
        self.quantiles = nn.Parameter(torch.tensor([-10.0, 0.0, 10.0]).repeat(channels, 1, 1))

    def _get_medians(self):
        return self.quantiles[:, :, 1:2].detach().reshape(1, -1, 1, 1)

    def forward(self, x, training=None):
        lik = torch.sigmoid(x + 0.5) - torch.sigmoid(x - 0.5)
        return torch.round(x), lik.clamp_min(1e-9)

    def loss(self):
        return self.quantiles.sum() * 0.0

    def update(self, force=False):
        return False


class GaussianConditional(nn.Module):
    def __init__(self, scale_table, *a, **k):
        super().__init__()

    def forward(self, x, scales, means=None, training=None):
        mu = 0.0 if means is None else means
        s = scales.abs().clamp_min(0.11)
        v = (x - mu).abs()
        n = torch.distributions.Normal(0.0, 1.0)
        lik = n.cdf((0.5 - v) / s) - n.cdf((-0.5 - v) / s)
        return torch.round(x - mu) + mu, lik.clamp_min(1e-9)

    def update_scale_table(self, scale_table, force=False):
        return False
