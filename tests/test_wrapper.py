"""Win_noShift_Attention wrapper: oracle vs the committed outputs of the live reference
module (CPU), drop-in surface (CPU), and the B200 drop-in with the fused gate kernel against both (GPU)."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_ops as R


def _t(a):
    return torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize("name", list(G.WRAPPER_CASES))
def test_oracle_wrapper_matches_reference(golden, name):
    cfg = G.WRAPPER_CASES[name]
    p = G.wrapper_inputs(cfg)
    g = golden["wrapper"]
    assert int(g[name + "/crc"]) == G.checksum(p["x"], p["alpha"], *[p["state"][k] for k in sorted(p["state"])])
    x = p["x"].clone().requires_grad_(True)
    w = {k: v.clone().requires_grad_(True) for k, v in p["state"].items()}
    y = R.win_noshift_attention(x, p["alpha"], w, cfg["heads"], cfg["ws"], cfg["shift"])
    torch.testing.assert_close(y, _t(g[name + "/y"]), rtol=1e-5, atol=2e-6)
    y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(cfg["seed"] + 5)))
    torch.testing.assert_close(x.grad, _t(g[name + "/dx"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w["conv_b.3.weight"].grad, _t(g[name + "/dconv_b3_w"]), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w["conv_a.0.conv.0.weight"].grad, _t(g[name + "/dconv_a0_w"]), rtol=1e-4, atol=1e-5)


def test_dropin_wrapper_has_the_reference_keys(pkg, golden):
    for name, cfg in G.WRAPPER_CASES.items():
        m = pkg.Win_noShift_Attention(cfg["C"], num_heads=cfg["heads"], window_size=cfg["ws"], shift_size=cfg["shift"])
        assert sorted(m.state_dict().keys()) == [str(k) for k in golden["wrapper"][name + "/keys"]]
        p = G.wrapper_inputs(cfg)
        missing, unexpected = m.load_state_dict(p["state"], strict=False)
        assert not unexpected and all(k.endswith("relative_position_index") for k in missing)


def test_gate_rejects_cpu_tensors(pkg):
    with pytest.raises(pkg.MwaB200Error):
        pkg.gate_residual(torch.zeros(4), torch.zeros(4), torch.zeros(4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.WRAPPER_CASES))
def test_dropin_wrapper_vs_golden(pkg, cuda_dev, golden, name):
    cfg = G.WRAPPER_CASES[name]
    p = G.wrapper_inputs(cfg)
    m = pkg.Win_noShift_Attention(cfg["C"], num_heads=cfg["heads"], window_size=cfg["ws"], shift_size=cfg["shift"])
    m.load_state_dict(p["state"], strict=False)
    m = m.to(cuda_dev)
    m.attn.algo = pkg.ALGO_SIMT                        # fp32 attention: isolates the wrapper + gate
    x = p["x"].to(cuda_dev).requires_grad_(True)
    y = m(x, p["alpha"].to(cuda_dev))
    g = golden["wrapper"]
    torch.testing.assert_close(y.detach().cpu(), _t(g[name + "/y"]), rtol=1e-3, atol=1e-4)
    y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(cfg["seed"] + 5)).to(cuda_dev))
    torch.testing.assert_close(x.grad.cpu(), _t(g[name + "/dx"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(m.conv_b[3].weight.grad.cpu(), _t(g[name + "/dconv_b3_w"]), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(m.conv_a[0].conv[0].weight.grad.cpu(), _t(g[name + "/dconv_a0_w"]), rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,channels_last", [((2, 192, 16, 24), False), ((2, 80, 8, 12), True), ((3, 5, 7), False)])
def test_gate_residual_forward_backward(pkg, cuda_dev, shape, channels_last):
    g = torch.Generator().manual_seed(1)
    a, b, x, gy = (torch.randn(shape, generator=g) * s for s in (1.0, 3.0, 1.0, 1.0))
    dev = [t.to(cuda_dev) for t in (a, b, x)]
    if channels_last:
        dev = [t.contiguous(memory_format=torch.channels_last) for t in dev]
    dev = [t.requires_grad_(True) for t in dev]
    out = pkg.gate_residual(*dev)
    out.backward(gy.to(cuda_dev))
    ref_in = [t.clone().double().requires_grad_(True) for t in (a, b, x)]
    ref = R.gate_residual(*ref_in)
    ref.backward(gy.double())
    torch.testing.assert_close(out.detach().double().cpu(), ref.detach(), rtol=1e-5, atol=1e-6)
    for d, r in zip(dev, ref_in):
        torch.testing.assert_close(d.grad.double().cpu(), r.grad, rtol=1e-5, atol=1e-6)


def test_accelerate_convs_switches_classes_and_keeps_values_and_keys(pkg):
    """post-construction switch of a model's convolutions to the drop-in classes: same state dict, same CPU results
    (without a CUDA tensor the drop-ins run torch.nn.functional, exactly like the classes they replace)"""
    import torch
    import torch.nn as nn
    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 8, 5, stride=2, padding=2), nn.GELU(),
                        nn.Sequential(nn.Conv2d(8, 32, 3, padding=1), nn.PixelShuffle(2)), nn.GELU(),
                        nn.ConvTranspose2d(8, 4, 5, stride=2, padding=2, output_padding=1), nn.ReLU(), nn.Conv2d(4, 2, 1))
    x = torch.randn(2, 3, 16, 24)
    ref = net(x)
    keys = list(net.state_dict().keys())
    assert pkg.accelerate_convs(net) == 4
    assert isinstance(net, pkg.conv.ConvStack) and isinstance(net[0], pkg.conv.Conv2d)
    assert isinstance(net[4], pkg.conv.ConvTranspose2d) and type(net[2]) is nn.Sequential
    assert list(net.state_dict().keys()) == keys
    torch.testing.assert_close(net(x), ref, rtol=1e-6, atol=1e-6)
