"""Latent rounding on the B200 (reference: `ste_round` and its call sites in the model files).

    ste_round(x)                 models/AutoEncoderRGB_Journal.py:31-32   round-half-even, identity gradient
    quantize_offset(x, mu)       :227-229 (z - medians) and :257 (y_slice - mu):  ste_round(x - mu) + mu  in ONE pass
    lrp_add(y_hat, lrp)          :262-264   y_hat + 0.5 * tanh(lrp)
    quantize_levels(m, 255)      :212-214   round(m * 255) / 255

All run the vectorised kernels of csrc/round.cu through the C ABI; channel chunks of a
(B, C, H, W) tensor (`y.chunk(10, 1)`) are consumed in place through a row stride.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import _abi


def _rows(t: torch.Tensor):
    """View `t` as rows of contiguous floats: returns (rows, row_len, row_stride) or None if not expressible."""
    if t.is_contiguous():
        return 1, t.numel(), t.numel()
    if t.dim() >= 2:
        inner = t[0]
        if inner.is_contiguous() and t.stride(0) >= inner.numel():
            return t.shape[0], inner.numel(), t.stride(0)
    return None


def _as_rows(t: torch.Tensor):
    r = _rows(t)
    if r is None:
        t = t.contiguous()
        r = _rows(t)
    return t, r


class _RoundSTE(Function):
    @staticmethod
    def forward(ctx, x):
        lib = _abi.load()
        _abi.require_cuda_f32(x, "ste_round input")
        x, (rows, n, sx) = _as_rows(x)
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _abi.check(lib.round_ste_forward(x.data_ptr(), out.data_ptr(), rows, n, sx, n, _abi.stream_handle()),
                       "round_ste_forward")
        return out

    @staticmethod
    def backward(ctx, g):
        return g


class _QuantizeOffset(Function):
    @staticmethod
    def forward(ctx, x, mu):
        lib = _abi.load()
        _abi.require_cuda_f32(x, "quantize_offset x")
        _abi.require_cuda_f32(mu, "quantize_offset mu")
        x, (rows, n, sx) = _as_rows(x)
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        ctx.mu_shape = mu.shape
        with torch.cuda.device(x.device):
            if mu.shape == x.shape:
                mu, (mrows, mn, sm) = _as_rows(mu)
                if (mrows, mn) != (rows, n):
                    if mrows == 1 and rows > 1:
                        sm = n                       # contiguous mu viewed with x's row split
                    elif rows == 1 and mrows > 1 and mrows * mn == n:
                        rows, n, sx = mrows, mn, mn  # contiguous x viewed with mu's row split
                    else:
                        mu = mu.contiguous()
                        sm = n if rows > 1 else mu.numel()
                st = lib.quantize_offset_forward(x.data_ptr(), mu.data_ptr(), out.data_ptr(), rows, n, sx, sm, n, 0, 1,
                                                 _abi.stream_handle())
            elif x.dim() == 4 and mu.numel() == x.shape[1] and tuple(mu.shape[-3:]) == (x.shape[1], 1, 1):
                C, hw = x.shape[1], x.shape[2] * x.shape[3]
                if rows == 1:
                    rows, n, sx = x.shape[0], C * hw, C * hw
                mu = mu.contiguous()
                st = lib.quantize_offset_forward(x.data_ptr(), mu.data_ptr(), out.data_ptr(), rows, n, sx, 0, n, C, hw,
                                                 _abi.stream_handle())
            else:
                raise _abi.MwaB200Error(f"quantize_offset: mu shape {tuple(mu.shape)} must equal x shape "
                                        f"{tuple(x.shape)} or be per-channel (1, C, 1, 1)")
            _abi.check(st, "quantize_offset_forward")
        return out

    @staticmethod
    def backward(ctx, g):
        # d/dx [ste_round(x - mu) + mu] = 1 ; d/dmu = -1 + 1 = 0
        return g, None if ctx.mu_shape is None else torch.zeros(ctx.mu_shape, dtype=g.dtype, device=g.device)


class _LrpAdd(Function):
    @staticmethod
    def forward(ctx, y_hat, lrp):
        lib = _abi.load()
        _abi.require_cuda_f32(y_hat, "lrp_add y_hat")
        _abi.require_cuda_f32(lrp, "lrp_add lrp")
        if y_hat.shape != lrp.shape:
            raise _abi.MwaB200Error("lrp_add: shapes differ")
        y_hat, (rows, n, sy) = _as_rows(y_hat)
        lrp, (lrows, ln, sl) = _as_rows(lrp)
        if (lrows, ln) != (rows, n):                 # bring both operands to a common row split
            if rows == 1 and lrows > 1 and ln * lrows == n:
                rows, n, sy = lrows, ln, ln
            elif lrows == 1 and rows > 1:
                sl = n
            else:
                lrp = lrp.contiguous()
                sl = n if rows > 1 else lrp.numel()
        out = torch.empty(y_hat.shape, dtype=y_hat.dtype, device=y_hat.device)
        with torch.cuda.device(y_hat.device):
            _abi.check(lib.lrp_add_forward(y_hat.data_ptr(), lrp.data_ptr(), out.data_ptr(), rows, n, sy, sl, n,
                                           _abi.stream_handle()), "lrp_add_forward")
        ctx.save_for_backward(lrp)
        return out

    @staticmethod
    def backward(ctx, g):
        (lrp,) = ctx.saved_tensors
        t = torch.tanh(lrp)
        return g, g * (0.5 * (1.0 - t * t))


def ste_round(x: torch.Tensor) -> torch.Tensor:
    """round-half-to-even with a straight-through (identity) gradient."""
    return _RoundSTE.apply(x)


def _rows_like(t: torch.Tensor, rows: int, n: int):
    """row stride of `t` when read as `rows` rows of `n` contiguous floats (what the kernels address), or None"""
    r = _rows(t)
    if r is None:
        return None
    trows, tn, ts = r
    if (trows, tn) == (rows, n):
        return ts
    if trows == 1 and tn == rows * n:
        return n
    return None


def _into(fn_name, a, b, out):
    """inference fast path: the elementwise kernels write straight into `out`, a dense tensor or a channel slice of one
    (e.g. the slice loop's support buffer), so that no `torch.cat` / copy kernel is needed afterwards"""
    lib = _abi.load()
    for name, t in (("input", a), ("second operand", b), ("out", out)):
        _abi.require_cuda_f32(t, f"{fn_name} {name}")
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        raise _abi.MwaB200Error(f"{fn_name}(out=...) is an inference path: it does not record autograd history")
    if not (a.shape == b.shape == out.shape):
        raise _abi.MwaB200Error(f"{fn_name}(out=...): shapes differ")
    r = _rows(a) or _rows(out)
    if r is None:
        a = a.contiguous()
        r = _rows(a)
    rows, n, _ = r
    sa, sb, so = _rows_like(a, rows, n), _rows_like(b, rows, n), _rows_like(out, rows, n)
    if sa is None:
        a = a.contiguous(); sa = _rows_like(a, rows, n)
    if sb is None:
        b = b.contiguous(); sb = _rows_like(b, rows, n)
    if so is None or sa is None or sb is None:
        raise _abi.MwaB200Error(f"{fn_name}(out=...): `out` must be dense or a channel slice of a dense tensor")
    with torch.cuda.device(a.device):
        if fn_name == "quantize_offset":
            st = lib.quantize_offset_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), rows, n, sa, sb, so, 0, 1,
                                             _abi.stream_handle())
        else:
            st = lib.lrp_add_forward(a.data_ptr(), b.data_ptr(), out.data_ptr(), rows, n, sa, sb, so, _abi.stream_handle())
        _abi.check(st, fn_name + "_forward")
    return out


def quantize_offset(x: torch.Tensor, mu: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """ste_round(x - mu) + mu, fused; mu has x's shape or is per-channel (1, C, 1, 1).  `out` (inference only, mu of x's
    shape): write into this tensor -- it may be a channel slice of a larger one, or x itself."""
    if out is not None:
        return _into("quantize_offset", x, mu, out)
    return _QuantizeOffset.apply(x, mu)


def lrp_add(y_hat: torch.Tensor, lrp: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """y_hat + 0.5 * tanh(lrp).  `out` (inference only): destination, may be a channel slice or y_hat itself."""
    if out is not None:
        return _into("lrp_add", y_hat, lrp, out)
    return _LrpAdd.apply(y_hat, lrp)


def quantize_levels(m: torch.Tensor, levels: int = 255) -> torch.Tensor:
    """round(m * levels) / levels  (no gradient, like the reference's plain torch.round)."""
    lib = _abi.load()
    _abi.require_cuda_f32(m, "quantize_levels input")
    m = m.contiguous()
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        _abi.check(lib.quantize_levels_forward(m.data_ptr(), out.data_ptr(), m.numel(), float(levels),
                                               _abi.stream_handle()), "quantize_levels_forward")
    return out
