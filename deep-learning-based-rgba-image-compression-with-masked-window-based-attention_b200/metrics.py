"""Evaluation metrics of the RGBA codec on the B200.

    masked_ms_ssim(X, Y, mask, data_range=1.0)   metrics/masked_ms_ssim_torch.py:181-265  ms_ssim(X, Y, mask, data_range)
    masked_psnr(X, Y, mask)                      models/AutoEncoderRGB_Journal.py:36-64 + trainRGB.py:303

Two kernel launches per MS-SSIM level (csrc/msssim.cu) instead of the reference's ~40 torch kernels; no host
synchronisation: the result is a 0-dim tensor on the device.
"""
from __future__ import annotations

import torch

from . import _abi

MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def masked_ms_ssim(X: torch.Tensor, Y: torch.Tensor, mask: torch.Tensor, data_range: float = 1.0,
                   weights=MS_SSIM_WEIGHTS) -> torch.Tensor:
    lib = _abi.load()
    for name, t in (("X", X), ("Y", Y), ("mask", mask)):
        _abi.require_cuda_f32(t, f"masked_ms_ssim {name}")
    if X.shape != Y.shape or X.dim() != 4:
        raise RuntimeError(f"masked_ms_ssim: X {tuple(X.shape)} and Y {tuple(Y.shape)} must be equal (B, C, H, W)")
    B, C, H, W = X.shape
    if mask.numel() != B * H * W:
        raise RuntimeError(f"masked_ms_ssim: mask {tuple(mask.shape)} must be (B, 1, H, W)")
    if min(H, W) <= (11 - 1) * 2 ** 4:
        # the reference's assertion (metrics/masked_ms_ssim_torch.py:222-225)
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    X, Y, m = X.contiguous(), Y.contiguous(), mask.reshape(B, H, W).contiguous()
    vals = []
    st = _abi.stream_handle()
    with torch.cuda.device(X.device):
        for lvl in range(len(weights)):
            sums = torch.empty(B, C, 2, device=X.device, dtype=torch.float32)
            cnt = torch.empty(B, device=X.device, dtype=torch.float32)
            _abi.check(lib.ms_ssim_level_forward(X.data_ptr(), Y.data_ptr(), m.data_ptr(), B, C, H, W, float(data_range),
                                                 sums.data_ptr(), cnt.data_ptr(), st), "ms_ssim_level_forward")
            per = sums / (cnt.view(B, 1, 1) + 1e-10)
            if lvl < len(weights) - 1:
                vals.append(torch.relu(per[..., 1]))
                Hp, Wp = (H + 2 * (H % 2) - 2) // 2 + 1, (W + 2 * (W % 2) - 2) // 2 + 1
                Xo = torch.empty(B, C, Hp, Wp, device=X.device, dtype=torch.float32)
                Yo, Mo = torch.empty_like(Xo), torch.empty(B, Hp, Wp, device=X.device, dtype=torch.float32)
                _abi.check(lib.ms_ssim_pool_forward(X.data_ptr(), Y.data_ptr(), m.data_ptr(), B, C, H, W, Xo.data_ptr(),
                                                    Yo.data_ptr(), Mo.data_ptr(), st), "ms_ssim_pool_forward")
                X, Y, m, H, W = Xo, Yo, Mo, Hp, Wp
            else:
                vals.append(torch.relu(per[..., 0]))
    w = torch.tensor(weights, device=X.device, dtype=torch.float32).view(-1, 1, 1)
    return torch.prod(torch.stack(vals, 0) ** w, dim=0).mean()


def masked_mse(X: torch.Tensor, Y: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """squared error over the pixels whose alpha is > 0, per image, averaged (models/AutoEncoderRGB_Journal.py:36-64)"""
    m = (mask.expand(-1, X.shape[1], -1, -1) > 0).to(X.dtype)
    se = ((X * m - Y * m) ** 2).sum(dim=(1, 2, 3))
    return torch.mean(se / torch.clamp(m.sum(dim=(1, 2, 3)), min=1))


def masked_psnr(X: torch.Tensor, Y: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """trainRGB.py:303"""
    return 10.0 * torch.log10(1.0 / masked_mse(X, Y, mask))
