"""B200 drop-in for the reference's `layers/SupplyMask.py`.

    SupplyMaskToTransform(kernel=3).forward(inputs) -> (mask1, ..., mask6)          (reference :7-18)

The alpha pyramid that feeds every masked window attention: six cascaded AvgPool2d(3, stride 2, padding 1).  The
reference runs six pooling kernels (plus three elementwise ones when the decoder first quantises the mask,
models/AutoEncoderRGB_Journal.py:212-215); here `alpha_pyramid` produces three levels per launch (csrc/alpha.cu) and
folds the quantisation into the first one.  Values are bit-exact with torch's AvgPool2d (same summation order).
No gradient flows through it (alpha is data in both training scripts).
Like the reference file, this module star-exports `torch`, `nn`, `math`, `F` and everything of `GDN` -- the model files
pick those names up through `from layers.SupplyMask import *` (models/AutoEncoderRGB_Journal.py:3).
"""
import math  # noqa: F401
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from .. import _abi
from .GDN import *  # noqa: F401,F403  (reference :4)


def alpha_pyramid(alpha: torch.Tensor, nlevels: int = 6, quant_levels: int = 0):
    """alpha (B, C, H, W) fp32 CUDA -> (recon, [level1 .. level_n]).

    level_k = AvgPool2d(3, 2, 1) applied k times; recon = round(alpha * quant_levels) / quant_levels (the plane the
    pyramid is built from) when quant_levels > 0, else `alpha` itself."""
    lib = _abi.load()
    _abi.require_cuda_f32(alpha, "alpha_pyramid input")
    if alpha.dim() != 4:
        raise RuntimeError(f"alpha_pyramid expects (B, C, H, W), got {tuple(alpha.shape)}")
    if not 1 <= nlevels <= 6:
        raise ValueError("alpha_pyramid: 1 <= nlevels <= 6")
    B, C, H, W = alpha.shape
    a = alpha.detach().contiguous()
    planes = B * C
    sizes, h, w = [], H, W
    for _ in range(nlevels):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        sizes.append((h, w))
    total = sum(planes * h * w for h, w in sizes)
    buf = torch.empty(total, dtype=torch.float32, device=alpha.device)
    recon = torch.empty_like(a) if quant_levels > 0 else None
    if planes > 0:
        with torch.cuda.device(alpha.device):
            _abi.check(lib.alpha_pyramid_forward(a.data_ptr(), _abi.ptr(recon), buf.data_ptr(), planes, H, W, nlevels,
                                                 int(quant_levels), _abi.stream_handle()), "alpha_pyramid_forward")
    levels, off = [], 0
    for h, w in sizes:
        levels.append(buf[off:off + planes * h * w].view(B, C, h, w))
        off += planes * h * w
    return (recon if recon is not None else alpha), levels


def constraint(tensor: torch.Tensor, quant_levels: int = 0) -> torch.Tensor:
    """Isolated-pixel clean-up of the decoded mask (reference trainRGB.py:98-111): a zero whose 8 neighbours sum to 8
    becomes 1, a positive value whose neighbours sum to 0 becomes 0.  Like the reference it updates `tensor` in place and
    returns it; one launch and no host synchronisation instead of conv2d + two boolean-mask assignments.
    quant_levels > 0 folds in the `round(clamp(t, 0, 1) * q) / q` the scripts run just before (trainRGB.py:285-286)."""
    lib = _abi.load()
    _abi.require_cuda_f32(tensor, "constraint input")
    if tensor.dim() != 4:
        raise RuntimeError(f"constraint expects (B, C, H, W), got {tuple(tensor.shape)}")
    B, C, H, W = tensor.shape
    src = tensor.detach().contiguous()
    out = torch.empty_like(src)
    if src.numel():
        with torch.cuda.device(tensor.device):
            _abi.check(lib.mask_constraint_forward(src.data_ptr(), out.data_ptr(), B * C, H, W, int(quant_levels),
                                                   _abi.stream_handle()), "mask_constraint_forward")
    with torch.no_grad():
        tensor.copy_(out)
    return tensor


class SupplyMaskToTransform(nn.Module):
    """Same constructor and return value as the reference (:7-18); only the 3 x 3 kernel it is built with is supported."""

    def __init__(self, kernel=3):
        super().__init__()
        if kernel != 3:
            raise NotImplementedError("SupplyMaskToTransform: the sm_100a kernel implements the reference's kernel=3")
        self.pool = nn.AvgPool2d(kernel, stride=2, padding=1)      # kept for attribute parity; holds no state

    def forward(self, inputs):
        _, levels = alpha_pyramid(inputs, 6)
        return tuple(levels)
