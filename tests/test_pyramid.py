"""Alpha pyramid (the alpha that feeds every masked window attention):
oracle vs the committed outputs of the live reference's SupplyMaskToTransform (CPU), drop-in surface (CPU), and the
two-launch CUDA pyramid against both, BIT-EXACT (GPU)."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_ops as R


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", list(G.PYRAMID_CASES))
def test_oracle_pyramid_matches_reference(golden, name):
    p = G.pyramid_inputs(G.PYRAMID_CASES[name])
    g = golden["pyramid"]
    assert int(g[name + "/crc"]) == G.checksum(p["alpha"], p["raw"])
    for k, lvl in enumerate(R.alpha_pyramid(p["alpha"]), 1):
        assert torch.equal(lvl, _t(g[f"{name}/mask{k}"]))
    recon = R.quantize_levels(p["raw"], 255)
    assert torch.equal(recon, _t(g[name + "/recon"]))
    for k, lvl in enumerate(R.alpha_pyramid(recon), 1):
        assert torch.equal(lvl, _t(g[f"{name}/md{k}"]))


def test_dropin_surface(pkg):
    m = pkg.SupplyMaskToTransform()
    assert isinstance(m.pool, torch.nn.AvgPool2d) and not list(m.state_dict())
    with pytest.raises(NotImplementedError):
        pkg.SupplyMaskToTransform(kernel=5)
    with pytest.raises(pkg.MwaB200Error):
        m(torch.zeros(1, 1, 8, 8))                       # CPU tensor: no fallback
    # the reference's model files take torch / nn / F / GDN from `from layers.SupplyMask import *`
    for name in ("torch", "nn", "F", "math", "GDN", "LowerBound", "SupplyMaskToTransform"):
        assert hasattr(pkg.SupplyMask, name), name
    lib = pkg._abi.load()
    assert lib.alpha_pyramid_level_offset(16, 512, 768, 6) == 16 * (256 * 384 + 128 * 192 + 64 * 96 + 32 * 48 + 16 * 24 + 8 * 12)
    assert lib.alpha_pyramid_level_offset(1, 37, 53, 1) == 19 * 27
    assert lib.alpha_pyramid_forward(None, None, None, 1, 8, 8, 7, 0, None) < 0      # nlevels out of range
    assert lib.alpha_pyramid_forward(None, None, None, 0, 8, 8, 6, 0, None) == 0     # empty batch is valid


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.PYRAMID_CASES))
def test_cuda_pyramid_bit_exact_vs_golden(pkg, cuda_dev, golden, name):
    p = G.pyramid_inputs(G.PYRAMID_CASES[name])
    g = golden["pyramid"]
    masks = pkg.SupplyMaskToTransform().to(cuda_dev)(p["alpha"].to(cuda_dev))
    assert len(masks) == 6
    for k, lvl in enumerate(masks, 1):
        assert torch.equal(lvl.cpu(), _t(g[f"{name}/mask{k}"])), f"mask{k}"
    recon, md = pkg.alpha_pyramid(p["raw"].to(cuda_dev), 6, quant_levels=255)
    assert torch.equal(recon.cpu(), _t(g[name + "/recon"]))
    for k, lvl in enumerate(md, 1):
        assert torch.equal(lvl.cpu(), _t(g[f"{name}/md{k}"])), f"md{k}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nlevels", [((16, 1, 512, 768), 6), ((2, 3, 100, 130), 4), ((1, 1, 1, 1), 6),
                                           ((4, 1, 257, 511), 5), ((1, 1, 64, 64), 1), ((0, 1, 32, 32), 6)])
def test_cuda_pyramid_bit_exact_vs_oracle(pkg, cuda_dev, shape, nlevels):
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.round(torch.rand(*shape, generator=g) * 255) / 255
    if shape[0]:
        a[0, :, : shape[2] // 2] = 0
    _, lv = pkg.alpha_pyramid(a.to(cuda_dev), nlevels)
    want = R.alpha_pyramid(a, nlevels)
    assert len(lv) == nlevels
    for k, (mine, ref) in enumerate(zip(lv, want), 1):
        assert mine.shape == ref.shape and torch.equal(mine.cpu(), ref), f"level {k}"


@pytest.mark.gpu
def test_pyramid_feeds_attention_like_the_reference_wiring(pkg, cuda_dev):
    """layers/TransformRGB.py:68,72: me2 -> attention1 (C=192, ws 8), me3 -> attention2: the keep decision taken from the
    CUDA pyramid equals the one taken from the oracle's."""
    g = torch.Generator().manual_seed(5)
    alpha = torch.zeros(2, 1, 128, 192)
    alpha[:, :, 30:90, 50:150] = torch.round(torch.rand(2, 1, 60, 100, generator=g) * 255) / 255
    masks = pkg.SupplyMaskToTransform()(alpha.to(cuda_dev))
    want = R.alpha_pyramid(alpha)
    for lvl, ws, s in ((1, 8, 4), (2, 4, 2)):
        assert torch.equal(R.window_keep(masks[lvl].cpu(), ws, s), R.window_keep(want[lvl], ws, s))


# ------------------------------------------------------------------------------------------------ constraint
@pytest.mark.parametrize("name", list(G.PYRAMID_CASES))
def test_oracle_constraint_matches_reference(golden, name):
    c = G.constraint_inputs(G.PYRAMID_CASES[name])
    g = golden["pyramid"]
    assert int(g[name + "/crc_c"]) == G.checksum(c["binary"], c["raw"])
    assert torch.equal(R.constraint(c["binary"]), _t(g[name + "/constraint_binary"]))
    assert not torch.equal(R.constraint(c["binary"]), c["binary"]) or c["binary"].numel() < 64   # the cases do fire
    m = R.quantize_levels(torch.clamp(c["raw"], 0, 1), 255)
    assert torch.equal(R.constraint(m), _t(g[name + "/constraint_chain"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.PYRAMID_CASES))
def test_cuda_constraint_bit_exact_vs_golden(pkg, cuda_dev, golden, name):
    c = G.constraint_inputs(G.PYRAMID_CASES[name])
    g = golden["pyramid"]
    t = c["binary"].to(cuda_dev)
    r = pkg.constraint(t)
    assert r is t                                          # in place, like the reference
    assert torch.equal(t.cpu(), _t(g[name + "/constraint_binary"]))
    chain = pkg.constraint(c["raw"].to(cuda_dev), quant_levels=255)      # clamp + quantise + clean-up in one launch
    assert torch.equal(chain.cpu(), _t(g[name + "/constraint_chain"]))


@pytest.mark.gpu
def test_cuda_constraint_full_size_properties(pkg, cuda_dev):
    """BASELINE size (16, 1, 512, 768): idempotent on its own output for binary masks without adjacent defects, and equal
    to the oracle"""
    g = torch.Generator().manual_seed(3)
    m = (torch.rand(16, 1, 512, 768, generator=g) < 0.5).float()
    want = R.constraint(m)
    got = pkg.constraint(m.to(cuda_dev).clone())
    assert torch.equal(got.cpu(), want)
    solid = torch.ones(2, 1, 64, 64)
    solid[:, :, 10, 10] = 0
    solid[:, :, 0, 5] = 0                                   # border pixel: only 5 neighbours -> stays
    out = pkg.constraint(solid.to(cuda_dev).clone()).cpu()
    assert out[0, 0, 10, 10] == 1 and out[0, 0, 0, 5] == 0
    assert torch.equal(pkg.constraint(out.to(cuda_dev).clone()).cpu(), out)
