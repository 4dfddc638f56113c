"""B200 drop-in for the reference's `layers/win_attention.py` (the alpha-less twin).

`WinBasedAttention.forward(x)` (reference :153-207) is the masked block with every window kept:
the same fused sm_100a kernel is launched with alpha == NULL.
"""
import torch
import torch.nn as nn

from .. import _abi
from ._attention_core import WindowAttentionBase, WindowAttentionFunction, _window_partition, _window_reverse


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def window_partition(x, window_size=8):
    """(B, H, W, C) -> (num_windows*B, window_size, window_size, C)   [reference :6-18]"""
    return _window_partition(x, window_size)


def window_reverse(windows, window_size, H, W):
    """(num_windows*B, window_size, window_size, C) -> (B, H, W, C)   [reference :21-34]"""
    return _window_reverse(windows, window_size, H, W)


class WindowAttention(WindowAttentionBase):
    """Window based multi-head self attention (W-MSA) module with relative position bias (reference :37-115)."""


class WinBasedAttention(nn.Module):
    """(S)W-MSA block on NCHW input (reference :118-207); same arguments as the reference."""

    def __init__(self, dim=192, num_heads=8, window_size=8, shift_size=0,
                 qkv_bias=True, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.shift_size = shift_size

        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        if drop_path != 0.:
            raise NotImplementedError("drop_path > 0 is not implemented by the fused sm_100a kernel")

        self.attn = WindowAttention(
            dim, window_size=to_2tuple(self.window_size), num_heads=num_heads,
            qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)

        self.drop_path = nn.Identity()
        self.algo = _abi.ALGO_AUTO

    def forward(self, x):
        a = self.attn
        a._refresh_if_training()
        return WindowAttentionFunction.apply(x, None, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                                             a.relative_position_bias_table, a, self.window_size, self.shift_size,
                                             self.algo)
