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
models.  With autograd history (training) the forward and the input gradient run on the same kernel
(the input gradient of a convolution is a convolution with the flipped / transposed weights: stride 1 -> stride 1, the 5x5
stride-2 convolution -> the transposed kernel, the transposed convolution -> the stride-2 kernel); the weight / bias
gradients (a contraction over all pixels) are the library's.  Anything else goes to torch.nn.functional (cuDNN) unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _abi

ACT_NONE, ACT_GELU, ACT_RELU, ACT_QUANT, ACT_LRP, ACT_GATE, ACT_ADD2 = 0, 1, 2, 3, 4, 5, 6


def _act_of(m):
    if isinstance(m, nn.GELU) and getattr(m, "approximate", "none") == "none":
        return ACT_GELU
    if isinstance(m, nn.ReLU):
        return ACT_RELU
    return None


def _apply_act(y, act):
    return F.gelu(y) if act == ACT_GELU else F.relu(y) if act == ACT_RELU else y


class SplitAct:
    """An activation as the fp16 hi / lo planes a convolution's TMA reads: int16 tensors [B][ps*ps][H/ps][W/ps][cstride]
    (channels last, ps = 2: the four pixel-parity planes a stride-2 convolution wants).  Written by the epilogue of the
    convolution that produced the activation (conv_forward_ex, out_hi / out_lo), so that between two convolutions the
    tensor crosses HBM once, as 4 bytes per element, instead of fp32 out + fp32 in + split out.  `C` channels starting at
    channel 0 are valid; `dense` optionally carries the fp32 NCHW tensor as well (residual connections need it)."""

    __slots__ = ("hi", "lo", "B", "C", "H", "W", "ps", "dense")

    def __init__(self, hi, lo, C, H, W, ps, dense=None):
        self.hi, self.lo, self.B, self.C, self.H, self.W, self.ps, self.dense = hi, lo, hi.shape[0], C, H, W, ps, dense

    @staticmethod
    def empty(B, C, H, W, ps, device, channels=None):
        cs = (max(C, channels or 0) + 7) // 8 * 8
        shape = (B, ps * ps, H // ps, W // ps, cs)
        return SplitAct(torch.empty(shape, dtype=torch.int16, device=device), torch.empty(shape, dtype=torch.int16, device=device),
                        C, H, W, ps)

    @property
    def cstride(self):
        return self.hi.shape[-1]

    @property
    def shape(self):
        return (self.B, self.C, self.H, self.W)

    @property
    def device(self):
        return self.hi.device

    def prefix(self, C):
        """the first C channels (same buffers): a torch.cat support of the slice loop"""
        return SplitAct(self.hi, self.lo, C, self.H, self.W, self.ps, None)


def split_into(x, dst: SplitAct, coff=0):
    """write the fp16 hi / lo planes of the fp32 NCHW tensor x into channels [coff, coff + C) of dst (conv_act_split)"""
    lib = _abi.load()
    B, C, H, W = x.shape
    if not (x.is_cuda and x.dtype == torch.float32 and _dense_nchw(x)) or (B, H, W) != (dst.B, dst.H, dst.W):
        raise RuntimeError(f"split_into: bad source {tuple(x.shape)} / {x.stride()} for planes of {dst.shape}")
    with torch.cuda.device(x.device):
        _abi.check(lib.conv_act_split(x.data_ptr(), x.stride(0), B, C, H, W, dst.ps, dst.hi.data_ptr(), dst.lo.data_ptr(),
                                      dst.cstride, coff, _abi.stream_handle()), "conv_act_split")
    _abi.count_launches(1)
    return dst


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
        """kind 0 / 1: convolution / transposed convolution weight; 2: the input gradient of a stride-1 convolution whose
        weight w is.  The buffer is kept across weight updates (training re-prepares every step)."""
        key = (w.data_ptr(), w._version, str(w.device))
        if self.img is None or key != self.key:
            lib = _abi.load()
            cin, cout = (w.shape[1], w.shape[0]) if kind == 0 else (w.shape[0], w.shape[1])
            nbytes = int(lib.conv_image_bytes(0 if kind == 2 else kind, cin, cout, k, stride))
            if nbytes <= 0:
                raise _abi.MwaB200Error("conv: unsupported geometry")
            img = self.img
            if img is None or img.numel() != nbytes or img.device != w.device:
                img = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
            wc = w.detach().contiguous()
            _abi.check(lib.conv_prepare(wc.data_ptr(), kind, cin, cout, k, stride, img.data_ptr(), nbytes,
                                        _abi.stream_handle()), "conv_prepare")
            self.img, self.key = img, key
        return self.img


GEMM_TOKENS_MIN_FLOP = 2e9
USE_KERNEL = True          # False: every convolution through torch.nn.functional (bench.py's library-convolution leg)


def _fast_ok(x, *tensors):
    if not USE_KERNEL:
        return False
    if isinstance(x, SplitAct):
        return True
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(t is not None and t.requires_grad for t in tensors)):
        return False
    return _dense_nchw(x)


def _run(x, bias, image, kind, k, stride, act, residual, cout, ho, wo, emit_ps=0, want_dense=True, aux=None, out=None,
         out2=None, emit_into=None, in_scale=None):
    """one conv_forward_ex call.  x: fp32 NCHW tensor or SplitAct.  Returns the dense result, or -- with emit_ps / emit_into
    -- a SplitAct (carrying the dense result too when want_dense)."""
    lib = _abi.load()
    B, cin, H, W = x.shape
    dev = x.device
    if isinstance(x, SplitAct):
        need_ps = 2 if (kind == 0 and stride == 2) else 1
        if x.ps != need_ps:
            raise RuntimeError(f"conv: split input has ps {x.ps}, this convolution needs {need_ps}")
        xp, xbs, in_hi, in_lo, in_cs = None, 0, x.hi, x.lo, x.cstride
    else:
        nsplit = int(lib.conv_split_bytes(B, cin, H, W))
        scratch = torch.empty(2, nsplit, dtype=torch.uint8, device=dev)
        xp, xbs, in_hi, in_lo, in_cs = x.data_ptr(), x.stride(0), scratch[0], scratch[1], 0
    if out is None and want_dense:
        out = torch.empty(B, cout, ho, wo, device=dev, dtype=torch.float32)
    if out is not None and (tuple(out.shape) != (B, cout, ho, wo) or not _dense_nchw(out)):
        raise RuntimeError(f"conv: bad output tensor {tuple(out.shape)} / {out.stride()}")
    if residual is not None:
        if tuple(residual.shape) != (B, cout, ho, wo):
            raise RuntimeError(f"conv: residual shape {tuple(residual.shape)} != output shape {(B, cout, ho, wo)}")
        residual = residual.contiguous()
    for name, t in (("aux", aux), ("out2", out2)):
        if t is not None and (tuple(t.shape) != (B, cout, ho, wo) or not _dense_nchw(t)):
            raise RuntimeError(f"conv: bad {name} tensor {tuple(t.shape)} / {t.stride()}")
    sp, coff = None, 0
    if emit_into is not None:
        sp, coff = emit_into
    elif emit_ps:
        sp = SplitAct.empty(B, cout, ho, wo, emit_ps, dev)
    b = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(dev):
        _abi.check(lib.conv_forward_ex(
            xp, xbs, in_hi.data_ptr(), in_lo.data_ptr(), in_cs, _abi.ptr(b), _abi.ptr(residual), _abi.ptr(out),
            0 if out is None else out.stride(0), _abi.ptr(aux), 0 if aux is None else aux.stride(0), _abi.ptr(out2),
            0 if out2 is None else out2.stride(0), None if sp is None else sp.hi.data_ptr(),
            None if sp is None else sp.lo.data_ptr(), 0 if sp is None else sp.ps, 0 if sp is None else sp.cstride, coff,
            image.data_ptr(), kind, B, cin, cout, H, W, k, stride, act, _abi.ptr(in_scale), _abi.stream_handle()),
            "conv_forward_ex")
    _abi.count_launches(1 if xp is None else 2)
    if sp is None:
        return out
    if emit_into is None:
        sp.dense = out
        return sp
    return out


def _dense_nchw(t):
    C, H, W = t.shape[1:]
    return t.stride(3) == 1 and t.stride(2) == W and t.stride(1) == H * W and t.stride(0) >= C * H * W


class _ShapeOnly:
    """stands in for a tensor that does not exist yet when a consumer is asked whether it will take the fast path"""

    is_cuda, dtype, requires_grad = True, torch.float32, False

    def __init__(self, shape, device):
        self.shape, self.device = tuple(shape), device

    def dim(self):
        return len(self.shape)

    def numel(self):
        n = 1
        for v in self.shape:
            n *= v
        return n

    def stride(self, i):
        st = [1] * len(self.shape)
        for j in range(len(self.shape) - 2, -1, -1):
            st[j] = st[j + 1] * self.shape[j + 1]
        return st[i]


def _fallback(x):
    if isinstance(x, SplitAct):
        if x.dense is None:
            raise RuntimeError("conv: a split activation without its dense tensor reached a convolution outside the fast path")
        return x.dense
    return x


def gradient_scale(g):
    """device scalar 2^k that brings max |g| to ~2^10: gradients sit far below fp16's normal range (6e-5), where the
    hi + lo split of the kernel's operands has no precision left; a power of two changes nothing else.  No host sync."""
    amax = g.detach().abs().amax().clamp_min(1e-30)
    return torch.exp2(torch.floor(torch.log2(1024.0 / amax))).reshape(1).float()


def gemm_tokens(x2d, w, bias=None, weight_is_in_out=False, scale_input=False):
    """out (T, Cout) = x2d (T, Cin) @ W^T (+ bias) on the convolution kernel (gemm_tokens_forward): the token GEMMs of the
    attention backward.  W is (Cout, Cin), or -- weight_is_in_out -- given as (Cin, Cout) and used as it lies (out = x2d @ w).
    Shapes the kernel does not take (T % 32, Cout % 8) and small problems (< 2 GFLOP: the split + prepare launches cost more
    than the library's fp32 GEMM there, measured on the 4x4-window layer and on 256x256 crops) go to the library GEMM."""
    T, cin = x2d.shape
    cout = w.shape[1] if weight_is_in_out else w.shape[0]
    if not (USE_KERNEL and x2d.is_cuda and x2d.dtype == torch.float32 and T > 0 and T % 32 == 0 and cout % 8 == 0
            and x2d.is_contiguous() and 2.0 * T * cin * cout >= GEMM_TOKENS_MIN_FLOP):
        out = x2d @ (w if weight_is_in_out else w.t())
        return out if bias is None else out + bias
    lib = _abi.load()
    nbytes = int(lib.conv_image_bytes(0, cin, cout, 1, 1))
    img = torch.empty(nbytes, dtype=torch.uint8, device=x2d.device)
    out = torch.empty(T, cout, device=x2d.device, dtype=torch.float32)
    cpad = (cin + 7) // 8 * 8
    split = torch.empty(2, T * cpad, dtype=torch.int16, device=x2d.device)
    sc = gradient_scale(x2d) if scale_input else None
    b = None if bias is None else bias.detach().contiguous()
    with torch.cuda.device(x2d.device):
        _abi.check(lib.conv_prepare(w.detach().contiguous().data_ptr(), 2 if weight_is_in_out else 0, cin, cout, 1, 1,
                                    img.data_ptr(), nbytes, _abi.stream_handle()), "conv_prepare")
        _abi.check(lib.gemm_tokens_forward(x2d.data_ptr(), T, cin, _abi.ptr(b), out.data_ptr(), cout, img.data_ptr(),
                                           split[0].data_ptr(), split[1].data_ptr(), _abi.ptr(sc), _abi.stream_handle()),
                   "gemm_tokens_forward")
    _abi.count_launches(3)
    return out


class _ConvTrainFn(torch.autograd.Function):
    """forward and input gradient on conv_forward_ex; weight / bias gradients through aten.convolution_backward"""

    @staticmethod
    def forward(ctx, x, weight, bias, mod):
        ctx.mod = mod
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        with torch.no_grad():
            return mod._fast(x, ACT_NONE, None)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        mod = ctx.mod
        g = g.contiguous()
        gx = gw = gb = None
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        with torch.no_grad():
            if need_x and mod._dgrad_ok(x):
                gx = mod._dgrad(g, x)
                need_x = False
            if need_x or need_w or need_b:
                transposed = mod._kind == 1
                r = torch.ops.aten.convolution_backward(
                    g, x, weight, [mod.out_channels] if ctx.has_bias else None, list(mod.stride), list(mod.padding),
                    list(mod.dilation), transposed, list(mod.output_padding) if transposed else [0, 0], mod.groups,
                    [bool(need_x), bool(need_w), bool(need_b)])
                gx = r[0] if need_x else gx
                gw = r[1] if need_w else None
                gb = r[2] if need_b else None
        return gx, gw, gb, None


def _train_ok(x, mod):
    """a call that records autograd history and that the kernel covers: dense fp32 NCHW on CUDA"""
    return (USE_KERNEL and torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.numel() > 0
            and _dense_nchw(x) and x.stride(0) == x.shape[1] * x.shape[2] * x.shape[3] and mod._covered(x))


class _FastConv:
    """what Conv2d and ConvTranspose2d share: the cached weight image, its invalidation, and the fast-path call"""

    def _init_fast(self):
        self._img = _WeightImage()
        self._dimg = _WeightImage()          # operand image of the input-gradient convolution

    def input_ps(self, x):
        """parity planes this layer wants its input split into (1 or 2) if it will take the fast path on x, else None"""
        if not (_fast_ok(x, self.weight, self.bias) and self._covered(x)):
            return None
        return 2 if (self._kind == 0 and self.stride[0] == 2) else 1

    def _fast(self, x, act, residual, **kw):
        k, s = self.kernel_size[0], self.stride[0]
        img = self._img.get(self.weight, self._kind, k, s)
        if self._kind == 0:
            ho, wo = x.shape[2] // s, x.shape[3] // s
        else:
            ho, wo = 2 * x.shape[2], 2 * x.shape[3]
        return _run(x, self.bias, img, self._kind, k, s, act, residual, self.out_channels, ho, wo, **kw)

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        for name in ("_img", "_dimg"):
            if name in self.__dict__:
                self.__dict__[name].invalidate()
        return out

    def _load_from_state_dict(self, *a, **kw):
        super()._load_from_state_dict(*a, **kw)
        for name in ("_img", "_dimg"):
            if name in self.__dict__:
                self.__dict__[name].invalidate()

    def _train_forward(self, x, act, residual, kw):
        """training: kernel forward + kernel input gradient; the epilogue steps as ordinary differentiable torch ops"""
        y = _ConvTrainFn.apply(x, self.weight, self.bias, self)
        return _plain_epilogue(y, act, residual, kw)


_EPI_KW = ("emit_ps", "want_dense", "aux", "out", "out2", "emit_into")


def _plain_epilogue(y, act, residual, kw):
    """the library path's version of the fused epilogues (training, uncovered geometries)"""
    aux = kw.get("aux")
    if act == ACT_QUANT:
        from .. import quant
        if kw.get("out2") is not None:
            kw["out2"].copy_(y)
        y = quant.quantize_offset(aux, y)
    elif act == ACT_LRP:
        y = aux + 0.5 * torch.tanh(y)
    elif act == ACT_GATE:
        y = aux * torch.sigmoid(y) + residual
    elif act == ACT_ADD2:
        y = (y + residual) + aux
    else:
        if residual is not None:
            y = y + residual
        y = _apply_act(y, act)
    if kw.get("out") is not None:
        kw["out"].copy_(y)
        y = kw["out"]
    if kw.get("emit_ps") or kw.get("emit_into") is not None:
        raise RuntimeError("conv: split emission was requested from a convolution outside the fast path")
    return y


class Conv2d(_FastConv, nn.Conv2d):
    min_channels = 1          # set 8 to send the 3-channel ends of the transforms (x1, the DSE 1x1s) to the library
    _kind = 0

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._init_fast()

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

    def _dgrad_ok(self, x):
        k, s = self.kernel_size[0], self.stride[0]
        return self.out_channels >= 8 and (s == 1 or k == 5)

    def _dgrad(self, g, x):
        """d loss / d x: stride 1 -> the same convolution with the taps flipped and the channel roles swapped; 5x5 stride 2 ->
        the transposed-convolution plan with the weight as it is ((Cout, Cin, k, k) reads as (in, out, k, k))"""
        k, s = self.kernel_size[0], self.stride[0]
        B, cin, H, W = x.shape
        if s == 1:
            img = self._dimg.get(self.weight, 2, k, 1)
            return _run(g, None, img, 0, k, 1, ACT_NONE, None, cin, H, W, in_scale=gradient_scale(g))
        img = self._dimg.get(self.weight, 1, k, 2)
        return _run(g, None, img, 1, k, 2, ACT_NONE, None, cin, H, W, in_scale=gradient_scale(g))

    def forward(self, x, act=ACT_NONE, residual=None, **kw):
        if _fast_ok(x, self.weight, self.bias, residual, kw.get("aux")) and self._covered(x):
            return self._fast(x, act, residual, **kw)
        if torch.is_grad_enabled() and _train_ok(x, self):
            return self._train_forward(x, act, residual, kw)
        return _plain_epilogue(super().forward(_fallback(x)), act, residual, kw)


class ConvTranspose2d(_FastConv, nn.ConvTranspose2d):
    min_channels = 8          # input channels; any number of output channels (<= 8 of them: the kernel's merged-class plan)
    _kind = 1

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._init_fast()

    def _covered(self, x):
        return (tuple(self.kernel_size) == (5, 5) and tuple(self.stride) == (2, 2) and tuple(self.padding) == (2, 2)
                and tuple(self.output_padding) == (1, 1) and tuple(self.dilation) == (1, 1) and self.groups == 1
                and x.shape[1] == self.in_channels and self.in_channels >= self.min_channels)

    def _dgrad_ok(self, x):
        return self.out_channels >= 8

    def _dgrad(self, g, x):
        """d loss / d x of the transposed convolution = the 5x5 stride-2 convolution of the output gradient with the same
        weight ((Cin, Cout, k, k) reads as (out, in, k, k))"""
        B, cin, H, W = x.shape
        img = self._dimg.get(self.weight, 0, 5, 2)
        return _run(g, None, img, 0, 5, 2, ACT_NONE, None, cin, H, W, in_scale=gradient_scale(g))

    def forward(self, x, output_size=None, act=ACT_NONE, residual=None, **kw):
        if output_size is None and _fast_ok(x, self.weight, self.bias, residual, kw.get("aux")) and self._covered(x):
            return self._fast(x, act, residual, **kw)
        if output_size is None and torch.is_grad_enabled() and _train_ok(x, self):
            return self._train_forward(x, act, residual, kw)
        return _plain_epilogue(super().forward(_fallback(x), output_size), act, residual, kw)


class ConvStack(nn.Sequential):
    """nn.Sequential whose forward folds the GELU / ReLU that follows a convolution -- directly, or after the PixelShuffle
    of a sub-pixel convolution (an elementwise function commutes with the shuffle) -- into that convolution's epilogue, and
    hands the activation between two adjacent convolutions over as fp16 hi / lo planes written by the producer's epilogue
    (SplitAct) instead of an fp32 tensor.  Children and state-dict keys are those of the plain Sequential.

    forward(x, residual=None, final_act=ACT_NONE, **epilogue options of the LAST convolution: emit_ps, want_dense, aux, out,
    out2, emit_into); x may be a SplitAct."""

    def forward(self, x, residual=None, final_act=ACT_NONE, **kw):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = _act_of(mods[i + 1]) if i + 1 < len(mods) else None
            step = 2 if nxt is not None else 1
            last = i + step >= len(mods)
            if isinstance(m, (Conv2d, ConvTranspose2d)):
                if last:
                    x = m(x, act=nxt if nxt is not None else final_act, residual=residual, **kw)
                else:
                    # the consumer is the next module: if it is a convolution on the fast path, feed it split planes only
                    consumer, ps = mods[i + step], None
                    if isinstance(consumer, (Conv2d, ConvTranspose2d)) and m.input_ps(x) is not None:
                        s = m.stride[0]
                        oshape = (x.shape[0], m.out_channels) + ((x.shape[2] // s, x.shape[3] // s) if m._kind == 0
                                                                 else (2 * x.shape[2], 2 * x.shape[3]))
                        ps = consumer.input_ps(_ShapeOnly(oshape, x.device))
                    if ps is not None:
                        x = m(x, act=nxt or ACT_NONE, emit_ps=ps, want_dense=False)
                    else:
                        x = m(x, act=nxt or ACT_NONE)
                i += step
            elif isinstance(m, nn.Sequential) and len(m) == 2 and isinstance(m[0], Conv2d) and isinstance(m[1], nn.PixelShuffle):
                x = m[1](m[0](x, act=nxt or ACT_NONE))
                i += step
            else:
                x = m(_fallback(x))
                i += 1
        return x
