"""The transforms' convolutions on the B200 (csrc/conv_tc.cu: implicit GEMM on tcgen05, fp16 hi + lo operands, fp32
accumulation) against the reference's own operator, torch.nn.functional.conv2d / conv_transpose2d in fp32 on the CPU
(layers/TransformRGB.py:16-100, layers/Masked_Attention.py:150-181, models/AutoEncoderRGB_Journal.py:139-203 call
exactly these).  Tolerance: BASELINE.json's 1e-3 relative / 1e-4 absolute, on every element."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# kind, Cin, Cout, k, stride, B, H, W, act, residual, bias
CASES = [
    ("conv", 192, 96, 1, 1, 2, 16, 32, 1, False, True),       # residual unit: conv1x1 + GELU
    ("conv", 96, 96, 3, 1, 2, 16, 32, 1, False, True),        # residual unit: conv3x3 + GELU
    ("conv", 96, 192, 1, 1, 2, 16, 32, 1, True, True),        # residual unit: conv1x1 + identity + GELU
    ("conv", 40, 40, 3, 1, 1, 8, 12, 1, False, True),         # C = 80 wrapper, grid smaller than one tile
    ("conv", 120, 224, 3, 1, 2, 24, 40, 1, False, True),      # slice loop, K not a multiple of 64, ragged tiles
    ("conv", 224, 128, 3, 1, 1, 16, 16, 1, False, True),
    ("conv", 128, 8, 3, 1, 2, 16, 16, 0, False, True),        # slice loop head: 8 output channels
    ("conv", 288, 320, 3, 1, 1, 16, 16, 0, False, True),      # hyper synthesis: two N blocks (sub-pixel conv 288 -> 4 x 80)
    ("conv", 80, 320, 3, 2, 1, 32, 32, 1, False, True),       # h_a: stride 2
    ("conv", 3, 192, 5, 2, 2, 32, 48, 0, False, True),        # analysis x1: 3 input channels, 5x5 stride 2
    ("conv", 192, 192, 5, 2, 1, 32, 32, 0, False, True),      # analysis x2 / x3
    ("conv", 32, 32, 3, 1, 1, 40, 48, 2, False, True),        # DSE: ReLU
    ("conv", 32, 3, 1, 1, 1, 40, 48, 0, True, True),          # DSE output conv + identity
    ("conv", 64, 64, 3, 1, 1, 8, 16, 0, False, False),        # no bias
    ("conv", 96, 192, 1, 1, 1, 6, 128, 1, True, True),        # 1x1 tiles follow the width: 1 x 128 pixels
    ("conv", 192, 96, 1, 1, 1, 5, 192, 1, False, True),       # 2 x 64 pixels, ragged last tile row
    ("conv", 32, 3, 1, 1, 1, 3, 256, 0, True, True),          # 1 x 128 pixels, 3 output channels
    ("deconv", 192, 192, 5, 2, 1, 16, 24, 0, False, True),    # synthesis x2 / x3
    ("deconv", 192, 3, 5, 2, 2, 12, 20, 0, False, True),      # synthesis x4: 3 output channels, ragged tiles
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}_{c[1]}to{c[2]}_k{c[3]}s{c[4]}_{c[6]}x{c[7]}")
def test_convolution_matches_torch_cpu_fp32(pkg, cuda_dev, case):
    kind, cin, cout, k, s, B, H, W, act, use_res, use_bias = case
    conv_mod = pkg.conv
    g = torch.Generator().manual_seed(cin * 7 + cout + k)
    x = torch.randn(B, cin, H, W, generator=g) * 1.5
    if kind == "conv":
        m = conv_mod.Conv2d(cin, cout, k, stride=s, padding=k // 2, bias=use_bias)
        m.min_channels = 1                 # exercise the kernel on the 3-channel ends too (the module routes them to the library)
    else:
        m = conv_mod.ConvTranspose2d(cin, cout, k, stride=s, padding=k // 2, output_padding=1, bias=use_bias)
    with torch.no_grad():
        m.weight.normal_(0, (1.0 / (cin * k * k)) ** 0.5, generator=g)
        if use_bias:
            m.bias.normal_(0, 0.2, generator=g)
        ref = F.conv2d(x, m.weight, m.bias, stride=s, padding=k // 2) if kind == "conv" else \
            F.conv_transpose2d(x, m.weight, m.bias, stride=s, padding=k // 2, output_padding=1)
        res = torch.randn(ref.shape, generator=g) if use_res else None
        if use_res:
            ref = ref + res
        ref = F.gelu(ref) if act == 1 else F.relu(ref) if act == 2 else ref
        m = m.to(cuda_dev)
        y = m(x.to(cuda_dev), act=act, residual=None if res is None else res.to(cuda_dev))
    assert y.shape == ref.shape
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-4)
    assert float((y.cpu() - ref).abs().max()) < 5e-6 * max(1.0, float(ref.abs().max()))     # fp16 hi + lo is ~2^-22 per product


def test_convolution_on_a_channel_slice_and_weight_update(pkg, cuda_dev):
    """the slice loop feeds channel-prefix views of its support buffer (batch stride > C H W); a weight update must reach
    the cached operand image"""
    m = pkg.conv.Conv2d(88, 224, 3, padding=1).to(cuda_dev)
    big = torch.randn(2, 128, 16, 24, device=cuda_dev)
    with torch.no_grad():
        y = m(big[:, :88])
        ref = F.conv2d(big[:, :88].cpu(), m.weight.cpu(), m.bias.cpu(), padding=1)
        torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-4)
        m.weight.mul_(0.5)
        y2 = m(big[:, :88])
        ref2 = F.conv2d(big[:, :88].cpu(), m.weight.cpu(), m.bias.cpu(), padding=1)
        torch.testing.assert_close(y2.cpu(), ref2, rtol=1e-3, atol=1e-4)


GRAD_CASES = [
    ("conv", 48, 96, 1, 1, 16, 24),        # residual-unit 1x1
    ("conv", 40, 40, 3, 1, 16, 24),        # 3x3 stride 1: flipped taps
    ("conv", 64, 72, 5, 2, 16, 32),        # 5x5 stride 2: input gradient on the transposed plan
    ("conv", 80, 96, 3, 2, 16, 16),        # 3x3 stride 2 (h_a): input gradient through the library
    ("deconv", 72, 64, 5, 2, 8, 12),       # transposed: input gradient on the stride-2 plan
]


@pytest.mark.parametrize("case", GRAD_CASES, ids=lambda c: f"{c[0]}_{c[1]}to{c[2]}_k{c[3]}s{c[4]}")
def test_training_forward_and_gradients_match_torch_cpu(pkg, cuda_dev, case):
    """with autograd recording the forward and the input gradient run on the kernel (the weight / bias gradients are the
    library's): everything against torch CPU fp32 autograd"""
    kind, cin, cout, k, s, H, W = case
    conv = pkg.conv
    g = torch.Generator().manual_seed(cin + cout + k)
    if kind == "conv":
        m = conv.Conv2d(cin, cout, k, stride=s, padding=k // 2)
    else:
        m = conv.ConvTranspose2d(cin, cout, k, stride=s, padding=k // 2, output_padding=1)
    x = torch.randn(2, cin, H, W, generator=g)
    ref_m = type(m).__mro__[2](*( (cin, cout, k) ), stride=s, padding=k // 2, **({"output_padding": 1} if kind == "deconv" else {}))
    ref_m.load_state_dict(m.state_dict())
    xr = x.clone().requires_grad_(True)
    yr = torch.nn.functional.gelu(ref_m(xr))
    go = torch.randn(yr.shape, generator=g)
    yr.backward(go)
    m = m.to(cuda_dev)
    xd = x.to(cuda_dev).requires_grad_(True)
    calls = []
    orig = conv._run

    def spy(xin, *a, **kw):
        calls.append(a[2])                   # kind of every kernel launch
        return orig(xin, *a, **kw)

    conv._run = spy
    try:
        yd = m(xd, act=conv.ACT_GELU)
        yd.backward(go.to(cuda_dev))
    finally:
        conv._run = orig
    assert len(calls) == (1 if (kind == "conv" and k == 3 and s == 2) else 2), calls
    torch.testing.assert_close(yd.detach().cpu(), yr.detach(), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(xd.grad.cpu(), xr.grad, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(m.weight.grad.cpu(), ref_m.weight.grad, rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(m.bias.grad.cpu(), ref_m.bias.grad, rtol=1e-3, atol=2e-4)


def test_conv_stack_fuses_activations(pkg, cuda_dev):
    conv = pkg.conv
    torch.manual_seed(0)
    stack = conv.ConvStack(conv.Conv2d(24, 32, 3, padding=1), torch.nn.GELU(),
                           torch.nn.Sequential(conv.Conv2d(32, 64, 3, padding=1), torch.nn.PixelShuffle(2)), torch.nn.GELU(),
                           conv.Conv2d(16, 8, 3, padding=1))
    plain = torch.nn.Sequential(*[m for m in stack])
    x = torch.randn(2, 24, 16, 16)
    with torch.no_grad():
        ref = plain(x)
        y = stack.to(cuda_dev)(x.to(cuda_dev))
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-4)


def _cpu_stack(stack):
    return torch.nn.Sequential(*[m for m in stack])


def test_chained_stack_hands_planes_between_convolutions(pkg, cuda_dev):
    """stride 2 -> 1 -> 2 chain (h_a of the model): every intermediate exists only as fp16 hi / lo planes written by the
    producer's epilogue (ps = 1 and ps = 2 layouts); result against torch CPU fp32"""
    conv = pkg.conv
    torch.manual_seed(1)
    g = torch.nn.GELU
    stack = conv.ConvStack(conv.Conv2d(80, 96, 3, stride=2, padding=1), g(), conv.Conv2d(96, 72, 3, padding=1), g(),
                           conv.Conv2d(72, 64, 3, stride=2, padding=1), g(), conv.Conv2d(64, 40, 1))
    x = torch.randn(2, 80, 32, 48)
    with torch.no_grad():
        ref = _cpu_stack(stack)(x)
        stack = stack.to(cuda_dev)
        calls = []
        orig = conv._run

        def spy(xin, *a, **kw):
            calls.append((type(xin).__name__, kw.get("emit_ps", 0), kw.get("want_dense", True)))
            return orig(xin, *a, **kw)

        conv._run = spy
        try:
            y = stack(x.to(cuda_dev))
        finally:
            conv._run = orig
    assert calls == [("Tensor", 1, False), ("SplitAct", 2, False), ("SplitAct", 1, False), ("SplitAct", 0, True)]
    torch.testing.assert_close(y.cpu(), ref, rtol=1e-3, atol=1e-4)


def test_planes_from_the_epilogue_equal_planes_from_the_split_kernel(pkg, cuda_dev):
    """a convolution fed by its producer's planes gives bit-identical results to one fed the producer's fp32 output"""
    conv = pkg.conv
    torch.manual_seed(2)
    a, b = conv.Conv2d(48, 56, 3, padding=1).to(cuda_dev), conv.Conv2d(56, 24, 3, padding=1).to(cuda_dev)
    x = torch.randn(2, 48, 20, 28, device=cuda_dev)
    with torch.no_grad():
        mid = a(x, act=1)
        y_dense = b(mid)
        sp = a(x, act=1, emit_ps=1, want_dense=True)
        assert torch.equal(sp.dense, mid)
        y_planes = b(sp)
    assert torch.equal(y_dense, y_planes)


def test_quantise_and_lrp_epilogues_are_bit_identical_to_the_rounding_kernels(pkg, cuda_dev):
    """models/AutoEncoderRGB_Journal.py:257, :262-264 as epilogues of the slice loop's last convolutions"""
    conv, quant = pkg.conv, pkg.quant
    torch.manual_seed(3)
    m = conv.Conv2d(128, 8, 3, padding=1).to(cuda_dev)
    x = torch.randn(2, 128, 16, 24, device=cuda_dev)
    ybig = torch.randn(2, 80, 16, 24, device=cuda_dev) * 4
    y_slice = ybig[:, 16:24]
    with torch.no_grad():
        mu = m(x)
        want = quant.quantize_offset(y_slice, mu)
        got, mu2 = torch.empty_like(mu), torch.empty_like(mu)
        sup = conv.SplitAct.empty(2, 96, 16, 24, 1, cuda_dev)
        sup.hi.zero_(); sup.lo.zero_()
        m(x, act=conv.ACT_QUANT, aux=y_slice, out=got, out2=mu2, emit_into=(sup, 88))
        assert torch.equal(mu2, mu)
        assert torch.equal(got, want)
        # the planes written behind the supports are the hi / lo halves of the same values
        hi = sup.hi[:, 0, :, :, 88:96].view(torch.float16).float().permute(0, 3, 1, 2)
        lo = sup.lo[:, 0, :, :, 88:96].view(torch.float16).float().permute(0, 3, 1, 2)
        torch.testing.assert_close(hi + lo, want, rtol=2e-6, atol=1e-7)     # fp16 hi + lo: ~2^-22 relative
        assert int(sup.hi[..., :88].abs().max()) == 0
        want_lrp = quant.lrp_add(want, mu)
        got_lrp = torch.empty_like(mu)
        m(x, act=conv.ACT_LRP, aux=want, out=got_lrp)
        assert torch.equal(got_lrp, want_lrp)
        inplace = want.clone()
        m(x, act=conv.ACT_LRP, aux=inplace, out=inplace)
        assert torch.equal(inplace, want_lrp)


def test_gate_epilogue_matches_the_gate_kernel(pkg, cuda_dev):
    """layers/Masked_Attention.py:186-188 as the epilogue of conv_b's last 1x1"""
    conv = pkg.conv
    torch.manual_seed(4)
    m = conv.Conv2d(80, 80, 1).to(cuda_dev)
    x, a, ident = (torch.randn(2, 80, 16, 24, device=cuda_dev) for _ in range(3))
    with torch.no_grad():
        want = pkg.Masked_Attention.gate_residual(a, m(x), ident)
        got = m(x, act=conv.ACT_GATE, aux=a, residual=ident)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)


def test_wrapper_chain_equals_module_by_module(pkg, cuda_dev):
    """Win_noShift_Attention in inference (planes between the residual units' convolutions, gate as an epilogue) against
    the same module run convolution by convolution"""
    torch.manual_seed(5)
    w = pkg.Masked_Attention.Win_noShift_Attention(80, 8, 4, 2).to(cuda_dev)
    x = torch.randn(2, 80, 16, 24, device=cuda_dev)
    alpha = (torch.rand(2, 1, 16, 24, device=cuda_dev) > 0.3).float()
    with torch.no_grad():
        y = w(x, alpha)
        a = w.conv_a(x)
        b = w.conv_b(w.attn(x, alpha))
        ref = pkg.Masked_Attention.gate_residual(a, b, x)
    torch.testing.assert_close(y, ref, rtol=1e-6, atol=1e-6)


def test_fused_slice_loop_equals_the_loop_with_separate_rounding_launches(pkg, cuda_dev):
    """models/AutoEncoderRGB_Journal.py:240-266: planes as supports + quantise / lrp epilogues against the same
    convolutions with torch-style supports and the rounding kernels: identical bits"""
    torch.manual_seed(6)
    model = pkg.codec.AutoEncoder().to(cuda_dev).eval()
    y = torch.randn(2, 80, 16, 24, device=cuda_dev) * 3
    lm, ls = torch.randn_like(y), torch.rand_like(y) + 0.2
    with torch.no_grad():
        got = model._slice_loop_fused(y, lm, ls, True)
        want = model._slice_loop_in_place(y, lm, ls, True)
    for g, w, name in zip(got, want, ("y_hat", "means", "scales")):
        assert torch.equal(g, w), name


def test_input_gradient_of_tiny_gradients_keeps_fp32_accuracy(pkg, cuda_dev):
    """real loss gradients are ~1e-7: far below fp16's normal range.  The power-of-two input scale (conv_forward_ex
    in_scale) keeps the hi + lo split exact enough: relative error of the input gradient as for O(1) gradients"""
    conv = pkg.conv
    torch.manual_seed(9)
    m = conv.Conv2d(64, 64, 3, padding=1).to(cuda_dev)
    x = torch.randn(2, 64, 16, 24, device=cuda_dev, requires_grad=True)
    go = torch.randn(2, 64, 16, 24, device=cuda_dev) * 3e-8
    m(x).backward(go)
    ref = torch.nn.functional.conv_transpose2d(go.double().cpu(), m.weight.double().cpu(), padding=1).float()
    err = (x.grad.detach().cpu() - ref).abs().max() / ref.abs().max()
    assert float(err) < 1e-5, float(err)


def test_weight_image_follows_data_writes_after_invalidate(pkg, cuda_dev):
    """a write through .data changes neither the pointer nor the version counter: invalidate_param_blocks() is the
    documented way to make the cached operand images follow it"""
    m = pkg.conv.Conv2d(16, 16, 3, padding=1).to(cuda_dev)
    x = torch.randn(1, 16, 8, 16, device=cuda_dev)
    with torch.no_grad():
        y0 = m(x)
        m.weight.data.mul_(2.0)
        assert pkg.invalidate_param_blocks(m) >= 1
        y1 = m(x)
        ref = F.conv2d(x.cpu(), m.weight.cpu(), m.bias.cpu(), padding=1)
    torch.testing.assert_close(y1.cpu(), ref, rtol=1e-3, atol=1e-4)
    assert not torch.allclose(y0, y1)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("ps,B,H,W", [(1, 2, 6, 10), (2, 2, 8, 12), (1, 1, 16, 24), (2, 3, 32, 48)])
def test_gdn_writes_the_planes_its_consumer_reads(pkg, cuda_dev, inverse, ps, B, H, W):
    """gdn_forward_planes (GDN -> convolution sites of layers/TransformRGB.py:55-61, :81-88): the planes are bit-identical to
    conv_act_split of gdn_forward's dense result, for ragged pixel counts (60 pixels < one 128-pixel tile) and both parities"""
    conv_mod = pkg.conv
    g = torch.Generator().manual_seed(17 + ps + H)
    m = pkg.GDN(192, inverse=inverse)
    with torch.no_grad():
        m.beta.add_(torch.rand(192, generator=g) * 0.5)
        m.gamma.add_(torch.rand(192, 192, generator=g) * 0.02)
        m = m.to(cuda_dev)
        x = (torch.randn(B, 192, H, W, generator=g) * 2.0).to(cuda_dev)
        dense = m(x)
        want = conv_mod.split_into(dense, conv_mod.SplitAct.empty(B, 192, H, W, ps, cuda_dev))
        got = m.request_planes(ps)(x)
    assert isinstance(got, conv_mod.SplitAct) and got.ps == ps and got.dense is None
    assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
    with torch.enable_grad():                                   # with autograd history the call stays dense
        xg = x.clone().requires_grad_(True)
        assert torch.is_tensor(m.request_planes(ps)(xg))


def test_transform_with_gdn_planes_matches_dense_gdn(pkg, cuda_dev, monkeypatch):
    """the analysis transform's GDN -> convolution hand-over through planes gives the result of the dense path bit for bit"""
    torch.manual_seed(5)
    enc = pkg.codec.Analysis_transform(192, 80).eval().to(cuda_dev)
    x = torch.rand(1, 3, 64, 96, device=cuda_dev)
    a = torch.ones(1, 1, 64, 96, device=cuda_dev)
    with torch.no_grad():
        _, me = pkg.alpha_pyramid(a, 3)
        y = enc(x, a, None, me[1], me[2], None)
        monkeypatch.setattr(pkg.GDN, "_planes_ok", lambda self, t: False)     # force the dense GDN + conv_act_split path
        y_dense = enc(x, a, None, me[1], me[2], None)
    assert torch.equal(y, y_dense)


@pytest.mark.parametrize("C,ws,ps", [(192, 8, 2), (192, 8, 1), (80, 4, 1)])
def test_attention_wrapper_writes_the_planes_its_consumer_reads(pkg, cuda_dev, C, ws, ps):
    """Win_noShift_Attention.request_planes (wrapper -> x3 / x1 sites of layers/TransformRGB.py:66-70, :91-97): the gate
    epilogue's planes are bit-identical to conv_act_split of the dense result"""
    conv_mod = pkg.conv
    torch.manual_seed(3 + C + ps)
    m = pkg.Win_noShift_Attention(C, num_heads=8, window_size=ws, shift_size=ws // 2).eval().to(cuda_dev)
    B, H, W = 2, 4 * ws, 6 * ws
    x = torch.randn(B, C, H, W, device=cuda_dev)
    alpha = (torch.rand(B, 1, H, W, device=cuda_dev) > 0.3).float()
    with torch.no_grad():
        dense = m(x, alpha)
        want = conv_mod.split_into(dense, conv_mod.SplitAct.empty(B, C, H, W, ps, cuda_dev))
        got = m.request_planes(ps)(x, alpha)
        again = m(x, alpha)                                       # the request is one-shot
    assert isinstance(got, conv_mod.SplitAct) and got.ps == ps and got.dense is None
    assert torch.equal(got.hi, want.hi) and torch.equal(got.lo, want.lo)
    assert torch.is_tensor(again) and torch.equal(again, dense)
