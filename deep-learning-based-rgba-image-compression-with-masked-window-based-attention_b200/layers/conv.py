"""Convolutions of the transforms on the B200 (the callers of the hot path).

    Conv2d / ConvTranspose2d     drop-in subclasses of torch.nn.Conv2d / ConvTranspose2d (same constructor, same
                                 parameters and state-dict keys); forward(x, act=..., residual=...) can fuse the
                                 activation and a residual add that follow the convolution in the reference's blocks
    ConvStack                    nn.Sequential that runs [conv, GELU / ReLU, conv, ..., PixelShuffle] chains with the
                                 activations folded into the convolutions' epilogues

Inference (no autograd history) on CUDA fp32 NCHW tensors runs `conv_forward` of csrc/conv_tc.cu: implicit GEMM on
tcgen05 with fp16 hi + lo operands in three passes and fp32 accumulation -- fp32-faithful (~1e-6 relative), where the
reference's torch.nn.Conv2d in fp32 is cuDNN's SIMT path on a GPU.  Covered: k in {1, 3, 5}, stride 1 or 2, padding k / 2,
groups 1, dilation 1; transposed k = 5, stride 2, padding 2, output padding 1 -- every convolution of the Journal
models.  Anything else, and every call that records autograd history (training: the backward is torch's), goes to
torch.nn.functional (cuDNN) unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _abi

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _act_of(m):
    if isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none":
        return ACT_GELU
    if isinstance(m, nn.ReLU):
        return ACT_RELU
    return None


def _apply_act(y, act):
    return F.gelu(y) if act == ACT_GELU else F.relu(y) if act == ACT_RELU else y


class _WeightImage:
    """kernel-ready weight image of one convolution, rebuilt when the weight changes (same key as ParamBlock; these
    modules only take the fast path without autograd history, i.e. with frozen weights)"""

    def __init__(self):
        self.key, self.img = None, None

    def invalidate(self):
        self.key = None

    def __deepcopy__(self, memo):
        return _WeightImage()

    def get(self, w, kind, k, stride):
        key = (w.data_ptr(), w._version, str(w.device))
        if self.img is None or key != self.key:
            lib = _abi.load()
            cin, cout = (w.shape[1], w.shape[0]) if kind == 0 else (w.shape[0], w.shape[1])
            nbytes = int(lib.conv_image_bytes(kind, cin, cout, k, stride))
            if nbytes <= 0:
                raise _abi.MwaB200Error("conv: unsupported geometry")
            img = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
            wc = w.detach().contiguous()
            _abi.check(lib.conv_prepare(wc.data_ptr(), kind, cin, cout, k, stride, img.data_ptr(), nbytes,
                                        _abi.stream_handle()), "conv_prepare")
            self.img, self.key = img, key
        return self.img


def _fast_ok(x, *tensors):
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(t is not None and t.requires_grad for t in tensors)):
        return False
    C, H, W = x.shape[1:]
    return x.stride(3) == 1 and x.stride(2) == W and x.stride(1) == H * W and x.stride(0) >= C * H * W


def _run(x, weight, bias, image, kind, k, stride, act, residual, cout, ho, wo):
    lib = _abi.load()
    B, cin, H, W = x.shape
    out = torch.empty(B, cout, ho, wo, device=x.device, dtype=x.dtype)
    nsplit = int(lib.conv_split_bytes(B, cin, H, W))
    split = torch.empty(2, nsplit, dtype=torch.uint8, device=x.device)
    if residual is not None:
        if residual.shape != out.shape:
            raise RuntimeError(f"conv: residual shape {tuple(residual.shape)} != output shape {tuple(out.shape)}")
        residual = residual.contiguous()
    b = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(x.device):
        _abi.check(lib.conv_forward(x.data_ptr(), x.stride(0), _abi.ptr(b), _abi.ptr(residual), out.data_ptr(),
                                    cout * ho * wo, image.data_ptr(), split[0].data_ptr(), split[1].data_ptr(), kind, B, cin,
                                    cout, H, W, k, stride, act, _abi.stream_handle()), "conv_forward")
    return out


class Conv2d(nn.Conv2d):
    min_channels = 8          # narrower ends go to the library (see _covered); tests set 1 to exercise the kernel there

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._img = _WeightImage()

    def _covered(self, x):
        k, s = self.kernel_size[0], self.stride[0]
        return (self.kernel_size[0] == self.kernel_size[1] and k in (1, 3, 5) and self.stride[0] == self.stride[1]
                and s in (1, 2) and tuple(self.padding) == (k // 2, k // 2) and tuple(self.dilation) == (1, 1)
                and self.groups == 1 and self.padding_mode == "zeros" and x.shape[1] == self.in_channels
                and (s == 1 or (x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0))
                # 3-channel ends of the transforms (x1, the DSE 1x1s): pure memory streams with K or N = 3, where the
                # 64-channel K blocks and 16-column N blocks of the GEMM kernel are mostly padding -- measured slower than
                # the library's fp32 kernels there, so they stay with it
                and self.in_channels >= self.min_channels and self.out_channels >= self.min_channels)

    def forward(self, x, act=ACT_NONE, residual=None):
        if _fast_ok(x, self.weight, self.bias, residual) and self._covered(x):
            k, s = self.kernel_size[0], self.stride[0]
            img = self._img.get(self.weight, 0, k, s)
            return _run(x, self.weight, self.bias, img, 0, k, s, act, residual, self.out_channels, x.shape[2] // s,
                        x.shape[3] // s)
        y = super().forward(x)
        if residual is not None:
            y = y + residual
        return _apply_act(y, act)

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if "_img" in self.__dict__:
            self._img.invalidate()
        return out

    def _load_from_state_dict(self, *a, **kw):
        super()._load_from_state_dict(*a, **kw)
        if "_img" in self.__dict__:
            self._img.invalidate()


class ConvTranspose2d(nn.ConvTranspose2d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._img = _WeightImage()

    def _covered(self, x):
        return (tuple(self.kernel_size) == (5, 5) and tuple(self.stride) == (2, 2) and tuple(self.padding) == (2, 2)
                and tuple(self.output_padding) == (1, 1) and tuple(self.dilation) == (1, 1) and self.groups == 1
                and x.shape[1] == self.in_channels)

    def forward(self, x, output_size=None, act=ACT_NONE, residual=None):
        if output_size is None and _fast_ok(x, self.weight, self.bias, residual) and self._covered(x):
            img = self._img.get(self.weight, 1, 5, 2)
            return _run(x, self.weight, self.bias, img, 1, 5, 2, act, residual, self.out_channels, 2 * x.shape[2],
                        2 * x.shape[3])
        y = super().forward(x, output_size)
        if residual is not None:
            y = y + residual
        return _apply_act(y, act)

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        if "_img" in self.__dict__:
            self._img.invalidate()
        return out

    def _load_from_state_dict(self, *a, **kw):
        super()._load_from_state_dict(*a, **kw)
        if "_img" in self.__dict__:
            self._img.invalidate()


class ConvStack(nn.Sequential):
    """nn.Sequential whose forward folds the GELU / ReLU that follows a convolution -- directly, or after the PixelShuffle
    of a sub-pixel convolution (an elementwise function commutes with the shuffle) -- into that convolution's epilogue.
    Children and state-dict keys are those of the plain Sequential."""

    def forward(self, x, residual=None, final_act=ACT_NONE):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            last = i == len(mods) - 1
            nxt = _act_of(mods[i + 1]) if i + 1 < len(mods) else None
            if isinstance(m, (Conv2d, ConvTranspose2d)):
                if nxt is not None:
                    x = m(x, act=nxt)
                    i += 2
                elif last:
                    x = m(x, act=final_act, residual=residual)
                    i += 1
                else:
                    x = m(x)
                    i += 1
            elif isinstance(m, nn.Sequential) and len(m) == 2 and isinstance(m[0], Conv2d) and isinstance(m[1], nn.PixelShuffle):
                if nxt is not None:
                    x = m[1](m[0](x, act=nxt))
                    i += 2
                else:
                    x = m[1](m[0](x))
                    i += 1
            else:
                x = m(x)
                i += 1
        return x
