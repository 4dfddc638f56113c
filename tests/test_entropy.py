"""The bitstream side (csrc/rans.cu host coder + <package>/entropy.py tables): round trips and stream length against the
entropy of the model that produced the tables.  CompressAI's own coder is absent, so byte parity is unpinned; these are
the properties a decoder depends on.  Host code: runs without a GPU."""
import importlib

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def ent(pkg):
    return importlib.import_module(pkg.codec.__name__.rsplit(".", 1)[0] + ".entropy")


def test_quantised_cdf_has_no_empty_symbol(ent):
    rng = np.random.default_rng(0)
    for n in (2, 5, 300, 3000):
        pmf = rng.random(n) ** 8                      # many near-zero probabilities
        pmf[rng.integers(0, n)] = 0.0
        cdf = ent.pmf_to_quantized_cdf(pmf / pmf.sum())
        assert cdf[0] == 0 and cdf[-1] == 1 << 16 and np.all(np.diff(cdf) > 0) and len(cdf) == n + 1


def test_rans_round_trip_with_escapes_and_resumed_decoding(ent):
    table = ent.gaussian_table(ent.get_scale_table())
    assert table.cdf.shape[0] == 64 and int(table.sizes[0]) == 5 and int(table.offsets[0]) == -1
    rng = np.random.default_rng(1)
    n = 50000
    idx = rng.integers(0, 64, n).astype(np.int32)
    scales = ent.get_scale_table().numpy()[idx]
    sym = np.round(rng.normal(0, 1, n) * scales).astype(np.int32)
    sym[::97] += rng.integers(-70000, 70000, sym[::97].size)          # far outside every table: multi-digit escapes
    stream = table.encode(sym, idx)
    assert len(stream) % 4 == 0
    dec = table.decoder(stream)
    out = np.concatenate([dec.decode(idx[:1]), dec.decode(idx[1:12345]), dec.decode(idx[12345:])])
    assert np.array_equal(out, sym)
    # without escapes the stream is within 2 % of the coded distribution's entropy
    sym2 = np.clip(np.round(rng.normal(0, 1, n) * scales), -4 * scales, 4 * scales).astype(np.int32)
    bits = 8 * len(table.encode(sym2, idx))
    pos = sym2 - table.offsets[idx]
    p = (table.cdf[idx, pos + 1] - table.cdf[idx, pos]) / 65536.0
    assert abs(bits - float(-np.log2(p).sum())) < 0.02 * bits
    # an empty message and a truncated stream
    assert np.array_equal(table.decoder(table.encode(sym[:0], idx[:0])).decode(idx[:0]), sym[:0])
    with pytest.raises(Exception):
        table.decoder(stream[:len(stream) // 2 // 4 * 4]).decode(idx)


def test_build_indexes_picks_the_first_table_scale_not_below(ent):
    st = ent.get_scale_table()
    s = torch.tensor([0.0, 0.05, 0.11, 0.12, 1.0, 255.9, 256.0, 1e4])
    idx = ent.build_indexes(s, st)
    for v, i in zip(s.tolist(), idx.tolist()):
        v = max(v, 0.11)
        assert i == min(int((st < v).sum()), 63)


def test_entropy_bottleneck_round_trip(pkg):
    torch.manual_seed(0)
    eb = pkg.codec.EntropyBottleneck(6)
    with torch.no_grad():
        eb.quantiles[:, 0, 1] = torch.linspace(-0.3, 0.4, 6)
        eb.quantiles[:, 0, 0] = -4.0
        eb.quantiles[:, 0, 2] = 5.0
    z = torch.randn(3, 6, 5, 7) * 6                     # beyond the tabulated support on both sides
    strings = eb.compress(z)
    assert len(strings) == 3
    z_hat = eb.decompress(strings, (5, 7))
    assert torch.equal(z_hat, torch.round(z - eb._get_medians()) + eb._get_medians())
