"""B200 drop-in for the reference's `layers/Masked_Attention.py` (first widening step).

    Win_noShift_Attention(dim, num_heads=8, window_size=8, shift_size=0).forward(x, mask)
        a = conv_a(x) ; b = conv_b(attn(x, mask)) ; out = a * sigmoid(b) + x          (reference :143-189)

Same constructor, children and state-dict keys as the reference (`conv_a.{0,1,2}.conv.{0,2,4}.*`, `attn.attn.*`,
`conv_b.{0,1,2}.conv.{0,2,4}.*`, `conv_b.3.*`).  `attn` is the fused sm_100a masked window attention.  In inference the
residual units run on the tcgen05 implicit-GEMM kernel (csrc/conv_tc.cu) as one chain: between two convolutions the
activation exists only as the fp16 hi / lo planes the consumer's TMA reads, written by the producer's epilogue; GELU, the
identity add and -- in conv_b's last 1x1 -- the gate + residual `a * sigmoid(b) + x` are epilogues too.  With autograd
recording the convolutions are torch's and the gate is ONE hand-written kernel (`gate_residual`, csrc/gate.cu), forward
and backward, instead of the reference's three elementwise kernels.
Only the names the Journal models use are provided (`conv1x1`, `Win_noShift_Attention`); the reference file's legacy
blocks (`ResBlockMask`, `AttentionMask`: they reference an undefined `CustomConv2DPyMV3`) are not reproduced.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _abi
from .masked_win_attention import *  # noqa: F401,F403  (the reference file star-imports it too, :6)
from .masked_win_attention import WinBasedAttention
from .conv import ACT_GATE, ACT_GELU, Conv2d, ConvStack, SplitAct


def conv1x1(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    """1x1 convolution (reference :11-13)."""
    return Conv2d(in_ch, out_ch, kernel_size=1, stride=stride)


def conv3x3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    """3x3 convolution with padding (compressai.layers.conv3x3, imported by the reference :8-10)."""
    return Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


class _GateResidual(Function):
    @staticmethod
    def forward(ctx, a, b, x):
        lib = _abi.load()
        for name, t in (("a", a), ("b", b), ("x", x)):
            _abi.require_cuda_f32(t, f"gate_residual {name}")
        if not (a.shape == b.shape == x.shape):
            raise RuntimeError(f"gate_residual: shape mismatch {tuple(a.shape)} {tuple(b.shape)} {tuple(x.shape)}")
        # one dense memory format for all three (x decides); the kernel is layout-agnostic
        fmt = torch.channels_last if (x.dim() == 4 and not x.is_contiguous()
                                      and x.is_contiguous(memory_format=torch.channels_last)) else torch.contiguous_format
        a, b, x = (t.contiguous(memory_format=fmt) for t in (a, b, x))
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _abi.check(lib.gate_residual_forward(a.data_ptr(), b.data_ptr(), x.data_ptr(), out.data_ptr(), x.numel(),
                                                 _abi.stream_handle()), "gate_residual_forward")
        ctx.fmt = fmt
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _abi.load()
        a, b = ctx.saved_tensors
        g = g.contiguous(memory_format=ctx.fmt)
        ga, gb = torch.empty_like(a), torch.empty_like(b)
        with torch.cuda.device(a.device):
            _abi.check(lib.gate_residual_backward(a.data_ptr(), b.data_ptr(), g.data_ptr(), ga.data_ptr(), gb.data_ptr(),
                                                  a.numel(), _abi.stream_handle()), "gate_residual_backward")
        return ga, gb, g


def gate_residual(a, b, x):
    """a * sigmoid(b) + x in one pass (reference layers/Masked_Attention.py:186-188)."""
    return _GateResidual.apply(a, b, x)


class Win_noShift_Attention(nn.Module):
    """Window-based self-attention module (reference :143-189)."""

    def __init__(self, dim, num_heads=8, window_size=8, shift_size=0):
        super().__init__()
        N = dim

        class ResidualUnit(nn.Module):
            """Simple residual unit."""

            def __init__(self):
                super().__init__()
                self.conv = ConvStack(
                    conv1x1(N, N // 2),
                    nn.GELU(),
                    conv3x3(N // 2, N // 2),
                    nn.GELU(),
                    conv1x1(N // 2, N),
                )
                self.relu = nn.GELU()

            def forward(self, x, **kw):
                # conv -> + identity -> GELU: the add and the activation ride in the last convolution's epilogue.
                # x may be a SplitAct (planes for the first 1x1, dense tensor for the identity); kw: emit_ps / want_dense
                return self.conv(x, residual=x.dense if isinstance(x, SplitAct) else x, final_act=ACT_GELU, **kw)

        self.conv_a = nn.Sequential(ResidualUnit(), ResidualUnit(), ResidualUnit())

        self.attn = WinBasedAttention(dim=dim, num_heads=num_heads, window_size=window_size, shift_size=shift_size)

        self.conv_b = nn.Sequential(
            ResidualUnit(),
            ResidualUnit(),
            ResidualUnit(),
            conv1x1(N, N),
        )

        self._emit_ps = 0

    def request_planes(self, ps):
        """one-shot request (the reference's forward signature stays as it is): if the NEXT forward call runs as the
        convolution chain, its result comes back only as the fp16 hi / lo planes (SplitAct) that a convolution of this
        package reads (ps = 2 for a stride-2 consumer) -- written by the gate epilogue, no fp32 tensor, no split launch"""
        self._emit_ps = int(ps)
        return self

    def forward(self, x, mask):
        emit_ps, self._emit_ps = getattr(self, "_emit_ps", 0), 0
        if self.conv_a[0].conv[0].input_ps(x) is None:            # training / uncovered: module by module
            a = self.conv_a(x)
            b = self.attn(x, mask)
            b = self.conv_b(b)
            return gate_residual(a, b, x)
        a = x
        for j in range(3):                                        # dense (identity of the next unit) + planes (its 1x1)
            a = self.conv_a[j](a, emit_ps=1 if j < 2 else 0)
        b = self.attn(x, mask)
        for j in range(3):                                        # the last unit feeds only the closing 1x1: planes alone
            b = self.conv_b[j](b, emit_ps=1, want_dense=j < 2)
        if emit_ps:
            return self.conv_b[3](b, act=ACT_GATE, aux=a, residual=x, emit_ps=emit_ps, want_dense=False)
        return self.conv_b[3](b, act=ACT_GATE, aux=a, residual=x)
