"""`timm.models.layers` names used by the reference (layers/masked_win_attention.py:3)."""
import collections.abc
from itertools import repeat

import torch
from torch import nn
from torch.nn.init import trunc_normal_  # noqa: F401  (re-export)


def to_2tuple(v):
    if isinstance(v, collections.abc.Iterable) and not isinstance(v, str):
        return tuple(v)
    return tuple(repeat(v, 2))


class DropPath(nn.Module):
    """Stochastic depth. The reference only ever builds it with p == 0 (identity)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        return x * torch.empty(shape, dtype=x.dtype, device=x.device).bernoulli_(keep) / keep
