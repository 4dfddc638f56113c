"""The whole RGBA codec forward on the B200 (product: <package>/codec.py on the sm_100a modules, convolutions through
cuDNN in fp32) against the outputs of the UNMODIFIED reference model committed in tests/golden/model_rgb.npz and against
the CPU oracle (oracle/ref_model.py) -- BASELINE.json's model-level criteria: x_hat within 1e-3 rel / 1e-4 abs, rounded
latents bit-exact except where the pre-rounding value sits on a .5 boundary, masked PSNR and masked MS-SSIM equal to 3
decimals."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as G
from oracle import ref_model as M

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-4


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _codec(pkg, model_keys, seed, dev):
    net = pkg.RGBACodec().eval()
    res = net.load_state_dict(G.model_state(model_keys, seed), strict=False)
    assert not res.unexpected_keys
    assert all(k.endswith("relative_position_index") or k.startswith("entropy_bottleneck._") for k in res.missing_keys)
    return net.to(dev)


@pytest.mark.parametrize("name", list(G.MODEL_CASES))
def test_full_forward_matches_reference_model(pkg, cuda_dev, golden, model_keys, name):
    cfg = G.MODEL_CASES[name]
    g = golden["model_rgb"]
    p = G.model_inputs(cfg)
    assert int(g[name + "/crc"]) == G.checksum(p["image"], p["alpha"], p["reconmask"]), "seeded inputs drifted"
    net = _codec(pkg, model_keys, cfg["seed"], cuda_dev)
    image, alpha, recon = (p[k].to(cuda_dev) for k in ("image", "alpha", "reconmask"))
    with torch.no_grad():
        r = {k: v.cpu() for k, v in net.detail(image, alpha, recon).items()}
        me = net.EncMakeMask(alpha)
        x_hat2, mse, bpp, bpp_y, bpp_z = net(image, alpha, recon, me[0], me[1], me[2], me[3])
    # the reference-signature forward is the same path (cuDNN may pick a different algorithm on the second call)
    torch.testing.assert_close(x_hat2.cpu(), r["x_hat"], rtol=1e-4, atol=1e-5)
    assert torch.isfinite(bpp).all() and float(bpp) > 0
    # pre-quantiser latent and hyper-latent
    torch.testing.assert_close(r["y"], _t(g[name + "/y"]), rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(r["z"], _t(g[name + "/z"]), rtol=RTOL, atol=ATOL)
    # rounded latents: same symbols except where y - mu is within float error of k + 0.5.  A flipped symbol changes the
    # support of the later slices, so only the slices up to the first flip are comparable symbol by symbol.
    yq, yq_ref = r["y_hat"], _t(g[name + "/y_hat"])
    diff = (yq - yq_ref).abs()
    flipped = diff > 0.5
    frac = float(flipped.float().mean())
    assert frac < 2e-3, frac
    sl = yq.shape[1] // 10
    first_flip = next((i for i in range(10) if bool(flipped[:, i * sl:(i + 1) * sl].any())), 10)
    if first_flip > 0:
        torch.testing.assert_close(yq[:, :first_flip * sl], yq_ref[:, :first_flip * sl], rtol=RTOL, atol=ATOL)
    if first_flip < 10:
        # every flip in the first differing slice is a genuine .5 tie: |(y - mu) - round| ~ 0.5 in the ORACLE's values
        w = G.model_state(model_keys, cfg["seed"])
        with torch.no_grad():
            o = M.rgb_forward(w, p["image"], p["alpha"], p["reconmask"])
        s = slice(first_flip * sl, (first_flip + 1) * sl)
        resid = (o["y"][:, s] - o["means"][:, s])
        tie = (resid - torch.floor(resid) - 0.5).abs()
        assert float(tie[flipped[:, s]].max()) < 2e-3
    # reconstruction
    x_ref = _t(g[name + "/x_hat"])
    if frac == 0.0:
        torch.testing.assert_close(r["x_hat"], x_ref, rtol=RTOL, atol=ATOL)
    else:
        assert float(((r["x_hat"] - x_ref).abs() > ATOL + RTOL * x_ref.abs()).float().mean()) < 0.05
    # metrics to 3 decimals: the reference's masked MSE -> PSNR (trainRGB.py:303) and masked MS-SSIM on the clipped output
    clipped = r["x_hat"].clamp(0, 1)
    psnr = M.psnr(M.masked_mse(p["image"], clipped, p["alpha"]))
    assert round(psnr, 3) == round(M.psnr(float(g[name + "/mse_clipped"])), 3), (psnr, float(g[name + "/mse_clipped"]))
    ms = float(M.masked_ms_ssim(p["image"], clipped, p["alpha"]))
    assert round(ms, 3) == round(float(g[name + "/ms_ssim"]), 3), (ms, float(g[name + "/ms_ssim"]))
    assert abs(float(mse) - float(g[name + "/mse"])) <= 2e-3 * float(g[name + "/mse"])


def test_full_forward_batch_and_baseline_size_against_oracle_samples(pkg, cuda_dev, model_keys):
    """BASELINE config 2 geometry (768 x 512), batch 2: latent of the first image against the CPU oracle (bounded: the
    oracle runs the analysis transform of one image), batch items independent of each other"""
    cfg = dict(B=2, H=512, W=768, drop=0.35, seed=77)
    p = G.model_inputs(cfg)
    net = _codec(pkg, model_keys, 61, cuda_dev)
    with torch.no_grad():
        r = net.detail(p["image"].to(cuda_dev), p["alpha"].to(cuda_dev), p["reconmask"].to(cuda_dev))
        r1 = net.detail(p["image"][1:].to(cuda_dev), p["alpha"][1:].to(cuda_dev), p["reconmask"][1:].to(cuda_dev))
    assert r["x_hat"].shape == (2, 3, 512, 768) and torch.isfinite(r["x_hat"]).all()
    torch.testing.assert_close(r["y"][1:], r1["y"], rtol=1e-4, atol=1e-5)      # cuDNN picks algorithms per batch size
    torch.testing.assert_close(r["x_hat"][1:], r1["x_hat"], rtol=1e-3, atol=1e-4)
    w = G.model_state(model_keys, 61)
    from oracle import ref_ops as R
    with torch.no_grad():
        me = R.alpha_pyramid(p["alpha"][:1])
        y0 = M.analysis(M._sub(w, "Encoder."), p["image"][:1], me[1], me[2])
    torch.testing.assert_close(r["y"][:1].cpu(), y0, rtol=RTOL, atol=ATOL)


def test_compress_decompress_round_trip(pkg, cuda_dev, model_keys):
    """models/AutoEncoderRGB_Journal.py:312-415: the decoder, working slice by slice from the strings alone, reproduces the
    encoder-side reconstruction bit for bit, and the strings are as long as the forward's rate estimate says"""
    cfg = next(iter(G.MODEL_CASES.values()))
    p = G.model_inputs(cfg)
    net = _codec(pkg, model_keys, cfg["seed"], cuda_dev)
    image = torch.cat([p["image"], p["image"].flip(3)], 0).to(cuda_dev)
    alpha = torch.cat([p["alpha"], p["alpha"].flip(3)], 0).to(cuda_dev)
    with torch.no_grad():
        out = net.compress(image, alpha)
        assert len(out["strings"][0]) == 2 and len(out["strings"][1]) == 2 and tuple(out["shape"]) == (3, 4)
        rec = net.decompress(out["strings"], out["shape"], alpha)["x_hat"]
        me = net.EncMakeMask(alpha)
        x_hat, _, bpp, _, _ = net(image, alpha, alpha, me[0], me[1], me[2], me[3])
    assert torch.equal(rec, x_hat.clamp(0, 1))
    nbits = 8 * sum(len(s) for group in out["strings"] for s in group)
    est = float(bpp) * image.shape[0] * image.shape[2] * image.shape[3]
    # the estimate is the continuous model's (likelihoods floored at 1e-9 = 30 bits a symbol); the coder works from the 64
    # quantised tables and codes far-out symbols through the escape, which is cheaper than that floor on a random-init
    # model: the stream must not be LONGER than the estimate, and not implausibly short (tests/test_entropy.py holds the
    # coder itself to 2 % of its tables' entropy)
    assert 0.5 * est < nbits < 1.05 * est + 4096, (nbits, est)
    # a single image: the reference's own structure (one y string, one z string)
    with torch.no_grad():
        one = net.compress(image[:1], alpha[:1])
        rec1 = net.decompress(one["strings"], one["shape"], alpha[:1])["x_hat"]
    assert torch.equal(rec1, rec[:1])


def test_host_pipeline_equals_direct_forward(pkg, cuda_dev, model_keys):
    """HostPipeline: copies on side streams around the forward; every batch's reconstruction lands in its pinned host
    buffer and equals the forward called directly"""
    cfg = next(iter(G.MODEL_CASES.values()))
    p = G.model_inputs(cfg)
    net = _codec(pkg, model_keys, cfg["seed"], cuda_dev)
    rgba = torch.cat([p["image"], p["alpha"]], 1)
    batches = [rgba.pin_memory(), rgba.flip(3).contiguous().pin_memory(), rgba.flip(2).contiguous().pin_memory()]
    outs = [torch.zeros(1, 3, cfg["H"], cfg["W"]).pin_memory() for _ in batches]
    pipe = pkg.HostPipeline(net)
    res = pipe.run(batches, outs)
    torch.cuda.synchronize()
    assert len(res) == 3
    with torch.no_grad():
        for hb, out in zip(batches, outs):
            d = hb.to(cuda_dev)
            me = net.EncMakeMask(d[:, 3:4])
            want = net(d[:, :3], d[:, 3:4], d[:, 3:4], me[0], me[1], me[2], me[3])[0]
            assert torch.equal(out, want.cpu())


def test_fused_rate_terms_match_the_torch_expressions(pkg, cuda_dev, model_keys, monkeypatch):
    """rate_forward (csrc/rate.cu) against the same forward with the rate / distortion tail as torch expressions
    (models/AutoEncoderRGB_Journal.py:36-64, :283-291; CompressAI's likelihood formulas): mse, bpp, y bpp, z bpp"""
    name = next(iter(G.MODEL_CASES))
    cfg = G.MODEL_CASES[name]
    p = G.model_inputs(cfg)
    net = _codec(pkg, model_keys, cfg["seed"], cuda_dev)
    with torch.no_grad():                       # a prior that is not at its initialisation: every parameter perturbed
        gen = torch.Generator().manual_seed(11)
        for n_, t in net.entropy_bottleneck.named_parameters():
            if n_ != "quantiles":
                t.add_((torch.randn(t.shape, generator=gen) * 0.3).to(cuda_dev))
    image, alpha, recon = (p[k].to(cuda_dev) for k in ("image", "alpha", "reconmask"))
    with torch.no_grad():
        me = net.EncMakeMask(alpha)
        fused = net(image, alpha, recon, me[0], me[1], me[2], me[3])
        monkeypatch.setattr(type(net), "_rate_terms_fused", lambda self, *a: None)
        plain = net(image, alpha, recon, me[0], me[1], me[2], me[3])
    assert torch.equal(fused[0], plain[0])
    for a, b, what in zip(fused[1:], plain[1:], ("mse", "bpp", "y bpp", "z bpp")):
        assert a.shape == b.shape == torch.Size([]), what
        torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-9, msg=lambda m: f"{what}: {m}")
    assert float(fused[3]) > 0 and float(fused[4]) > 0
