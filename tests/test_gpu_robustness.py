"""Robustness of the product path on the GPU: empty batches, non-default streams, parameter updates between calls,
run-to-run determinism, odd window counts (a tile that is only partly filled), NHWC through the same modules,
NaN / Inf containment (a poisoned window must not leak into its neighbours)."""
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_ops as R

pytestmark = pytest.mark.gpu


def _attn(pkg, dev, C=192, heads=8, ws=8, shift=4, seed=0):
    torch.manual_seed(seed)
    m = pkg.MaskedWinBasedAttention(C, heads, ws, shift)
    with torch.no_grad():
        m.attn.relative_position_bias_table.normal_(0, 0.2)
        m.attn.qkv.bias.normal_(0, 0.1)
    return m.to(dev)


def test_empty_batch(pkg, cuda_dev):
    m = _attn(pkg, cuda_dev)
    x = torch.empty(0, 192, 16, 16, device=cuda_dev)
    a = torch.empty(0, 1, 16, 16, device=cuda_dev)
    with torch.no_grad():
        assert m(x, a).shape == x.shape
        assert pkg.GDN(192).to(cuda_dev)(x).shape == x.shape
        assert pkg.ste_round(torch.empty(0, device=cuda_dev)).numel() == 0


def test_attention_on_a_side_stream_and_deterministic(pkg, cuda_dev):
    m = _attn(pkg, cuda_dev)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 192, 24, 40, generator=g).to(cuda_dev)            # 45 windows: the last tile is half empty
    a = G.blob_alpha(3, 24, 40, 8, 4, 0.3, seed=8).to(cuda_dev)
    with torch.no_grad():
        y0 = m(x, a)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            y1 = m(x, a)
        s.synchronize()
        y2 = m(x, a)
    assert torch.equal(y0, y1) and torch.equal(y0, y2)                   # same stream or not, bit-identical every run
    keep = R.window_keep(a.cpu(), 8, 4)
    assert 0 < int(keep.sum()) < keep.numel()


def test_attention_follows_parameter_updates(pkg, cuda_dev):
    """the cached parameter block (weight images, folded biases) must be rebuilt after an in-place update"""
    m = _attn(pkg, cuda_dev, C=80, heads=8, ws=4, shift=2)
    x = torch.randn(2, 80, 8, 12, device=cuda_dev)
    a = torch.ones(2, 1, 8, 12, device=cuda_dev)
    with torch.no_grad():
        y0 = m(x, a).clone()
        m.attn.qkv.bias.add_(0.5)                     # touches q, k and v biases (k: no effect; v: via the folded bias)
        m.attn.proj.weight.mul_(1.5)
        y1 = m(x, a)
    w = [t.detach().cpu() for t in (m.attn.qkv.weight, m.attn.qkv.bias, m.attn.proj.weight, m.attn.proj.bias,
                                    m.attn.relative_position_bias_table)]
    ref = R.masked_window_attention(x.cpu(), a.cpu(), *w, 8, 4, 2)
    assert (y1 - y0).abs().max() > 1e-2
    assert (y1.cpu() - ref).abs().max() < 5e-3


def test_nan_stays_inside_its_window(pkg, cuda_dev):
    m = _attn(pkg, cuda_dev)
    x = torch.randn(1, 192, 32, 32, device=cuda_dev)
    a = torch.ones(1, 1, 32, 32, device=cuda_dev)
    with torch.no_grad():
        y_clean = m(x, a)
        xp = x.clone()
        xp[0, 5, 4 + 9, 4 + 10] = float("nan")        # shifted-frame window (1, 1): rows / cols 12..19 of the image
        y = m(xp, a)
    bad = torch.isnan(y[0]).any(dim=0)                # (H, W) map of poisoned pixels
    ys, xs = torch.nonzero(bad, as_tuple=True)
    assert bad.any() and ys.min() >= 12 and ys.max() <= 19 and xs.min() >= 12 and xs.max() <= 19
    ok = ~bad
    assert torch.equal(y[0][:, ok], y_clean[0][:, ok])


@pytest.mark.parametrize("C,heads,ws,shift", [(192, 8, 8, 4), (80, 8, 4, 2)])
def test_channels_last_gives_the_same_result(pkg, cuda_dev, C, heads, ws, shift):
    m = _attn(pkg, cuda_dev, C, heads, ws, shift)
    m.algo = pkg.ALGO_SIMT                            # fp32 on both layouts -> tight comparison
    x = torch.randn(2, C, 4 * ws, 6 * ws, device=cuda_dev)
    a = G.blob_alpha(2, 4 * ws, 6 * ws, ws, shift, 0.3, seed=4).to(cuda_dev)
    with torch.no_grad():
        y = m(x, a)
        ycl = m(x.contiguous(memory_format=torch.channels_last), a)
    assert ycl.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(ycl.contiguous(), y, rtol=1e-5, atol=1e-6)
    gm = pkg.GDN(C).to(cuda_dev)
    with torch.no_grad():
        torch.testing.assert_close(gm(x.contiguous(memory_format=torch.channels_last)).contiguous(), gm(x), rtol=1e-5,
                                   atol=1e-6)


# ------------------------------------------------------------------------------------------------ parameter-block cache
def test_param_block_follows_weight_updates(pkg, cuda_dev):
    """the cached kernel-ready parameter block must never serve stale weights: in-place updates through the parameter
    (optimizer step, load_state_dict) are seen through `_version`; writes through `.data` are not visible to any key and
    need `invalidate_param_blocks`; a captured CUDA graph re-derives the block at replay time"""
    import copy
    torch.manual_seed(0)
    g = pkg.GDN(192).to(cuda_dev)
    x = torch.randn(2, 192, 16, 24, device=cuda_dev)
    with torch.no_grad():
        y0 = g(x)
        g.beta.mul_(1.5)                                     # bumps _version
        y1 = g(x)
        assert not torch.equal(y0, y1)
        ref = copy.deepcopy(g)
        g.beta.data.mul_(2.0)                                # invisible to the cache key
        ref.beta.data.mul_(2.0)
        assert pkg.invalidate_param_blocks(g) == 1
        torch.testing.assert_close(g(x), ref(x), rtol=0, atol=0)
        sd = {k: v.clone() for k, v in g.state_dict().items()}
        sd["beta"] = sd["beta"] * 0.5
        g.load_state_dict(sd)
        ref.load_state_dict(sd)
        torch.testing.assert_close(g(x), ref(x), rtol=0, atol=0)
        # CUDA graph: capture, then change the weights, replay -> the replay uses the new weights
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = g(x)
        graph.replay()
        torch.cuda.synchronize()
        before = out.clone()
        g.beta.mul_(1.25)
        graph.replay()
        torch.cuda.synchronize()
        assert not torch.equal(before, out)
        torch.testing.assert_close(out, g(x), rtol=0, atol=0)


def test_param_block_is_rebuilt_every_call_in_training(pkg, cuda_dev):
    m = pkg.MaskedWinBasedAttention(80, 8, 4, 2).to(cuda_dev)
    x = torch.randn(1, 80, 8, 8, device=cuda_dev)
    a = torch.ones(1, 1, 8, 8, device=cuda_dev)
    y0 = m(x, a)                                             # grad enabled: prepare runs on every call
    m.attn.proj.weight.data.mul_(0.5)                        # a write no key can see
    y1 = m(x, a)
    assert not torch.equal(y0, y1)
    with torch.no_grad():
        m.attn.proj.weight.mul_(2.0)
        torch.testing.assert_close(m(x, a), y0.detach(), rtol=1e-5, atol=1e-6)


def test_param_block_cross_stream_hit_waits_for_the_fill(pkg, cuda_dev):
    g = pkg.GDN(192).to(cuda_dev)
    x = torch.randn(4, 192, 64, 96, device=cuda_dev)
    with torch.no_grad():
        ref = g(x).clone()
        g.invalidate_param_block()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            y1 = g(x)                                        # fills the block on s1
        with torch.cuda.stream(s2):
            y2 = g(x)                                        # cache hit on another stream: must wait for the fill
        torch.cuda.synchronize()
    assert torch.equal(y1, ref) and torch.equal(y2, ref)
