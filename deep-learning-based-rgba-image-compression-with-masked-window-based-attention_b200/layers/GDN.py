"""B200 drop-in for the reference's `layers/GDN.py` (same names, constructor, parameters, forward).

    GDN(ch, inverse=False, beta_min=1e-6, gamma_init=0.1, reparam_offset=2**-18).forward(inputs)
        y_i = x_i / sqrt(beta_i + sum_j gamma[i, j] * x_j^2)        (inverse: multiplied)

Reference: layers/GDN.py:26-94 (module), :9-23 (LowerBound).  Forward and backward run the
hand-written sm_100a kernels behind include/mwa_b200.h (`gdn_prepare`, `gdn_forward`,
`gdn_backward`); CPU tensors are rejected -- there is no fallback.
State-dict keys: `beta` (C,), `gamma` (C, C) -- identical to the reference.
"""
import torch
import torch.utils.data
from torch import nn, optim
from torch.nn import functional as F
from torch.autograd import Function

from .. import _abi
from ._params import ParamBlock, ParamBlockOwner


class LowerBound(Function):
    """max(inputs, bound) with the reference's pass-through gradient (layers/GDN.py:9-23).

    Elementwise on tiny (C,) / (C, C) parameters; kept in PyTorch so `LowerBound.apply` stays usable
    by callers.  The GDN kernels apply the same rule internally (gdn_prepare / gdn_backward).
    """

    @staticmethod
    def forward(ctx, inputs, bound):
        b = torch.ones_like(inputs) * bound
        ctx.save_for_backward(inputs, b)
        return torch.max(inputs, b)

    @staticmethod
    def backward(ctx, grad_output):
        inputs, b = ctx.saved_tensors
        keep = (inputs >= b) | (grad_output < 0)
        return keep.type(grad_output.dtype) * grad_output, None


def _composed_backward(lib, x, grad_y, beta, gamma, blk, mod, channels_last, gx, gb, gg):
    """GDN backward as hand-written elementwise stages around three plain GEMMs over the pixel dimension
    (n = gamma x^2, t = gamma^T dn, dgamma = dn (x^2)^T; include/mwa_b200.h, "GEMM-composed GDN backward").
    The GEMMs run on the fp32 effective gamma / gamma^T stored in the parameter block."""
    B, C = x.shape[0], x.shape[1]
    hw = x.shape[2] * x.shape[3]
    st = _abi.stream_handle()

    def mat(which):
        off = int(lib.gdn_param_offset(C, which))
        return blk[off:off + 4 * C * C].view(torch.float32).view(C, C)

    g_eff, gT_eff = mat(1), mat(2)                       # gamma[i][j], gamma^T[j][i]

    def tok(t):      # (B, C, H, W) tensor -> GEMM view: NCHW (B, C, HW); NHWC (B*HW, C)
        return t.permute(0, 2, 3, 1).reshape(B * hw, C) if channels_last else t.view(B, C, hw)

    x2 = torch.empty_like(x)
    _abi.check(lib.gdn_bwd_square(x.data_ptr(), x2.data_ptr(), x.numel(), st), "gdn_bwd_square")
    nb = torch.empty_like(x)
    if channels_last:
        torch.matmul(tok(x2), gT_eff, out=tok(nb))       # n_i = sum_j x2_j gamma[i][j]
    else:
        torch.matmul(g_eff, tok(x2), out=tok(nb))
    dn = torch.empty_like(x)
    _abi.check(lib.gdn_bwd_dn(grad_y.data_ptr(), x.data_ptr(), nb.data_ptr(), blk.data_ptr(), dn.data_ptr(), B, C, hw,
                              int(mod.inverse), int(channels_last), st), "gdn_bwd_dn")
    t = torch.empty_like(x)
    if channels_last:
        torch.matmul(tok(dn), g_eff, out=tok(t))         # t_j = sum_i dn_i gamma[i][j]
    else:
        torch.matmul(gT_eff, tok(dn), out=tok(t))
    _abi.check(lib.gdn_bwd_dx(nb.data_ptr(), x.data_ptr(), t.data_ptr(), gx.data_ptr(), x.numel(), st), "gdn_bwd_dx")
    with _abi.tf32_reduction(hw if not channels_last else B * hw):   # K = pixels: TF32 products, fp32 accumulation
        if channels_last:
            dg = tok(dn).t() @ tok(x2)                   # (C, BHW) x (BHW, C)
        else:
            dg = torch.zeros(C, C, device=x.device)
            dnv, x2v = tok(dn), tok(x2)
            for b in range(B):                           # accumulating GEMMs: no reduction kernel
                dg.addmm_(dnv[b], x2v[b].t())
    scratch = torch.empty(C, device=x.device)
    _abi.check(lib.gdn_bwd_finalize(dn.data_ptr(), dg.data_ptr(), beta.data_ptr(), gamma.data_ptr(), mod.beta_bound,
                                    mod.gamma_bound, gb.data_ptr(), gg.data_ptr(), scratch.data_ptr(), B, C, hw,
                                    int(channels_last), st), "gdn_bwd_finalize")


class _GDNFunction(Function):
    @staticmethod
    def forward(ctx, x, beta, gamma, mod):
        lib = _abi.load()
        _abi.require_cuda_f32(x, "GDN input")
        _abi.require_cuda_f32(beta, "GDN.beta")
        _abi.require_cuda_f32(gamma, "GDN.gamma")
        B, C = x.shape[0], x.shape[1]
        if gamma.shape != (C, C) or beta.shape != (C,):
            raise RuntimeError(f"GDN built for {beta.shape[0]} channels got input with {C}")
        channels_last = (not x.is_contiguous()) and x.is_contiguous(memory_format=torch.channels_last)
        if not channels_last:
            x = x.contiguous()
        hw = x.shape[2] * x.shape[3]
        if x.numel() == 0:                   # empty batch: nothing to launch
            ctx.mod, ctx.channels_last = mod, channels_last
            ctx.save_for_backward(x, beta, gamma, beta)
            return torch.empty_like(x)
        with torch.cuda.device(x.device):
            blk = mod._param_block(beta, gamma)
            y = torch.empty_like(x)          # preserves the memory format
            _abi.check(lib.gdn_forward(x.data_ptr(), y.data_ptr(), blk.data_ptr(), B, C, hw, int(mod.inverse),
                                       int(channels_last), mod.algo, _abi.stream_handle()), "gdn_forward")
        ctx.mod = mod
        ctx.channels_last = channels_last
        ctx.save_for_backward(x, beta, gamma, blk)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        lib = _abi.load()
        x, beta, gamma, blk = ctx.saved_tensors
        mod = ctx.mod
        if x.numel() == 0:
            return grad_y, torch.zeros_like(beta), torch.zeros_like(gamma), None
        B, C = x.shape[0], x.shape[1]
        hw = x.shape[2] * x.shape[3]
        if ctx.channels_last:
            grad_y = grad_y.contiguous(memory_format=torch.channels_last)
        else:
            grad_y = grad_y.contiguous()
        with torch.cuda.device(x.device):
            gx = torch.empty_like(x)
            gb = torch.empty_like(beta)
            gg = torch.empty_like(gamma)
            if mod.algo != _abi.ALGO_SIMT and x.numel() % 4 == 0 and x.numel() > 0:
                _composed_backward(lib, x, grad_y, beta, gamma, blk, mod, ctx.channels_last, gx, gb, gg)
                return gx, gb, gg, None
            ws_bytes = lib.gdn_backward_workspace_bytes(B, C, hw)
            ws = torch.empty(max(int(ws_bytes), 16), dtype=torch.uint8, device=x.device)
            _abi.check(lib.gdn_backward(x.data_ptr(), grad_y.data_ptr(), beta.data_ptr(), gamma.data_ptr(),
                                        blk.data_ptr(), mod.beta_bound, mod.gamma_bound, gx.data_ptr(),
                                        gb.data_ptr(), gg.data_ptr(), ws.data_ptr(), ws.numel(), B, C, hw,
                                        int(mod.inverse), int(ctx.channels_last), _abi.stream_handle()),
                       "gdn_backward")
        return gx, gb, gg, None


class GDN(ParamBlockOwner, nn.Module):
    """Generalized divisive normalization layer (B200 kernels).
    y[i] = x[i] / sqrt(beta[i] + sum_j(gamma[i, j] * x[j]^2))
    """

    def __init__(self,
                 ch,
                 inverse=False,
                 beta_min=1e-6,
                 gamma_init=0.1,
                 reparam_offset=2**-18,
                 ):
        super(GDN, self).__init__()
        self.inverse = inverse
        self.beta_min = beta_min
        self.gamma_init = gamma_init
        self.reparam_offset = reparam_offset
        self.algo = _abi.ALGO_AUTO
        self._blk = ParamBlock()
        self._emit_ps = 0
        self.build(ch)

    def build(self, ch):
        self.pedestal = self.reparam_offset ** 2
        self.beta_bound = (self.beta_min + self.reparam_offset ** 2) ** 0.5
        self.gamma_bound = self.reparam_offset
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + self.pedestal))
        self.gamma = nn.Parameter(torch.sqrt(self.gamma_init * torch.eye(ch) + self.pedestal))

    def _param_block(self, beta, gamma):
        lib = _abi.load()
        C = beta.shape[0]
        nbytes = int(lib.gdn_param_bytes(C))

        def fill(blk):
            _abi.check(lib.gdn_prepare(beta.data_ptr(), gamma.data_ptr(), C, self.beta_bound, self.gamma_bound,
                                       self.pedestal, blk.data_ptr(), blk.numel(), _abi.stream_handle()),
                       "gdn_prepare")

        return self._blk.get((beta, gamma), nbytes, fill)

    def _planes_ok(self, x):
        """the call can hand its result to a convolution of this package as fp16 hi / lo planes (gdn_forward_planes): inference
        (no autograd history) on a dense fp32 NCHW CUDA tensor with the 192 channels of the tcgen05 kernel"""
        if torch.is_grad_enabled() and (x.requires_grad or self.beta.requires_grad or self.gamma.requires_grad):
            return False
        return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 192 and x.numel() > 0
                and x.is_contiguous() and self.algo in (_abi.ALGO_AUTO, _abi.ALGO_TCGEN05))

    def forward_planes(self, inputs, ps):
        """GDN for a convolution consumer (layers/TransformRGB.py:55-61, :81-88: gdn -> x2 / x4): returns the SplitAct that
        convolution reads (ps = 2 for a stride-2 consumer) -- the fp32 result is never written and no split launch runs."""
        from .conv import SplitAct
        lib = _abi.load()
        B, C, H, W = inputs.shape
        self._refresh_if_training()
        with torch.cuda.device(inputs.device):
            blk = self._param_block(self.beta, self.gamma)
            sp = SplitAct.empty(B, C, H, W, ps, inputs.device)
            _abi.check(lib.gdn_forward_planes(inputs.data_ptr(), sp.hi.data_ptr(), sp.lo.data_ptr(), ps, sp.cstride,
                                              blk.data_ptr(), B, C, H, W, int(self.inverse), _abi.stream_handle()),
                       "gdn_forward_planes")
        return sp

    def request_planes(self, ps):
        """one-shot request (the reference's forward signature stays as it is): the NEXT forward call returns the SplitAct
        a convolution of this package reads (ps = 2 for a stride-2 consumer) if it can produce it, else the dense result"""
        self._emit_ps = int(ps)
        return self

    def forward(self, inputs):
        emit_ps, self._emit_ps = getattr(self, "_emit_ps", 0), 0
        if emit_ps and inputs.dim() == 4 and self._planes_ok(inputs):
            return self.forward_planes(inputs, emit_ps)
        unfold = False
        if inputs.dim() == 5:
            unfold = True
            bs, ch, d, w, h = inputs.size()
            inputs = inputs.view(bs, ch, d * w, h)
        self._refresh_if_training()
        outputs = _GDNFunction.apply(inputs, self.beta, self.gamma, self)
        if unfold:
            outputs = outputs.view(bs, ch, d, w, h)
        return outputs
