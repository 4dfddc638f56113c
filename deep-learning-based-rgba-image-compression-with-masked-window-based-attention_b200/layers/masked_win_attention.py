"""B200 drop-in for the reference's `layers/masked_win_attention.py`.

Same module-level names (`window_partition`, `window_reverse`, `remove_zero_windows`,
`WindowAttention`, `WinBasedAttention`), constructor / forward signatures and state-dict keys
(`attn.qkv.weight|bias`, `attn.proj.weight|bias`, `attn.relative_position_bias_table`,
buffer `attn.relative_position_index`).

`WinBasedAttention.forward(x, img_alpha)` (reference :169-251) runs ONE fused sm_100a kernel:
cyclic shift + window partition (index arithmetic, nothing is rolled or copied in HBM), the
keep predicate "sum of the window's alpha != 0" evaluated on the device (no host sync, CUDA-graph
capturable), QKV projection, QK^T + relative-position bias + SW-MSA region mask, softmax, PV,
output projection, window reverse / un-shift and the residual add.  Dropped windows return x.
"""
import torch
import torch.nn as nn

from .. import _abi
from ._attention_core import WindowAttentionBase, WindowAttentionFunction, _window_partition, _window_reverse


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def window_partition(x, window_size=8):
    """(B, H, W, C) -> (num_windows*B, window_size, window_size, C)   [reference :6-18; torch view ops]"""
    return _window_partition(x, window_size)


def window_reverse(windows, window_size, H, W):
    """(num_windows*B, window_size, window_size, C) -> (B, H, W, C)   [reference :20-33]"""
    return _window_reverse(windows, window_size, H, W)


def remove_zero_windows(x, alpha):
    """Keep the windows whose alpha sums to != 0; returns (kept windows, bool mask)   [reference :35-47].

    Utility kept for API compatibility (it host-syncs through boolean indexing, like the reference);
    the fused forward never calls it -- it evaluates the predicate inside the kernel.
    """
    mask = alpha.sum(dim=(1, 2, 3)) != 0
    return x[mask], mask


class WindowAttention(WindowAttentionBase):
    """Window based multi-head self attention (W-MSA) module with relative position bias (reference :49-131)."""


class WinBasedAttention(nn.Module):
    """Alpha-aware (S)W-MSA block (reference :134-251).

    Args: dim, num_heads, window_size, shift_size, qkv_bias, qk_scale, drop, attn_drop, drop_path --
    identical to the reference.  drop / attn_drop / drop_path must be 0 (the models never set them).
    """

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

    def forward(self, x, img_alpha):
        a = self.attn
        a._refresh_if_training()
        return WindowAttentionFunction.apply(x, img_alpha, a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                                             a.relative_position_bias_table, a, self.window_size, self.shift_size,
                                             self.algo)
