"""Entropy-coding side of the codec: quantised CDF tables of the two entropy models and the
binding of the host rANS coder (csrc/rans.cu).

Reference call sites: models/AutoEncoderRGB_Journal.py:306-311 (update), :319-320 (entropy_bottleneck.compress /
decompress), :330-368 / :374-403 (BufferedRansEncoder / RansDecoder with the Gaussian conditional's tables).  The classes
behind them are CompressAI's (third-party, absent, unversioned): the table construction below restates their published
definitions (scale table exp(linspace(log 0.11, log 256, 64)), tail mass 1e-9, 16-bit CDFs with every symbol given a
non-zero frequency) -- enough for a self-consistent bitstream; byte parity with CompressAI is UNPINNED.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _abi

PRECISION = 16
SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256.0, 64
TAIL_MASS = 1e-9
# -norm.ppf(TAIL_MASS / 2): how many standard deviations the tabulated support reaches
TAIL_MULTIPLIER = 6.109410204869


def get_scale_table(lo=SCALES_MIN, hi=SCALES_MAX, levels=SCALES_LEVELS) -> torch.Tensor:
    """models/AutoEncoderRGB_Journal.py:23-27"""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def pmf_to_quantized_cdf(pmf: np.ndarray, precision: int = PRECISION) -> np.ndarray:
    """frequencies proportional to pmf summing to 2^precision, none of them zero (a zero is filled by taking one count from
    the smallest frequency above 1, as CompressAI's pmf_to_quantized_cdf does)"""
    freq = np.round(np.asarray(pmf, dtype=np.float64) * (1 << precision)).astype(np.int64)
    total = int(freq.sum())
    if total <= 0:
        raise ValueError("pmf_to_quantized_cdf: empty distribution")
    cdf = np.concatenate([[0], np.cumsum(freq)])
    cdf = ((1 << precision) * cdf) // total
    cdf[-1] = 1 << precision
    n = len(cdf) - 1
    for i in range(n):
        if cdf[i] == cdf[i + 1]:
            f = np.diff(cdf)
            cand = np.where(f > 1)[0]
            if cand.size == 0:
                raise ValueError("pmf_to_quantized_cdf: more symbols than counts")
            j = int(cand[np.argmin(f[cand])])
            if j < i:
                cdf[j + 1:i + 1] -= 1
            else:
                cdf[i + 1:j + 1] += 1
    assert np.all(np.diff(cdf) > 0) and cdf[0] == 0 and cdf[-1] == 1 << precision
    return cdf.astype(np.int32)


class CdfTable:
    """(ncdf, max_len) int32 quantised CDFs + per-row length and symbol offset, on the host (the coder runs there)"""

    def __init__(self, pmfs, tail_masses, lengths, offsets):
        n, max_len = len(lengths), int(max(lengths)) + 2
        self.cdf = np.zeros((n, max_len), dtype=np.int32)
        for i in range(n):
            prob = np.concatenate([np.asarray(pmfs[i][:lengths[i]], dtype=np.float64), [float(tail_masses[i])]])
            c = pmf_to_quantized_cdf(prob)
            self.cdf[i, :len(c)] = c
        self.sizes = (np.asarray(lengths, dtype=np.int32) + 2).astype(np.int32)
        self.offsets = np.asarray(offsets, dtype=np.int32)

    def _p(self, a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def encode(self, symbols: np.ndarray, indexes: np.ndarray) -> bytes:
        lib = _abi.load()
        symbols = np.ascontiguousarray(symbols, dtype=np.int32).reshape(-1)
        indexes = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
        cap = 16 + 8 * symbols.size + 64
        while True:
            out = np.empty(cap, dtype=np.uint8)
            n = lib.rans_encode_with_indexes(self._p(symbols), self._p(indexes), symbols.size, self._p(self.cdf),
                                             self.cdf.shape[1], self._p(self.sizes), self._p(self.offsets), self.cdf.shape[0],
                                             self._p(out), cap)
            if n == -4:                       # MWA_ERR_WORKSPACE: heavy use of the escape code; retry with more room
                cap *= 4
                continue
            if n < 0:
                _abi.check(int(n), "rans_encode_with_indexes")
            return out[:n].tobytes()

    def decoder(self, stream: bytes) -> "StreamDecoder":
        return StreamDecoder(self, stream)


class StreamDecoder:
    """RansDecoder.set_stream + decode_stream: successive calls continue on the same stream"""

    def __init__(self, table: CdfTable, stream: bytes):
        self.t, self.buf = table, np.frombuffer(stream, dtype=np.uint8).copy()
        self.state = np.zeros(4, dtype=np.int64)

    def decode(self, indexes: np.ndarray) -> np.ndarray:
        lib = _abi.load()
        t = self.t
        indexes = np.ascontiguousarray(indexes, dtype=np.int32).reshape(-1)
        out = np.empty(indexes.size, dtype=np.int32)
        _abi.check(lib.rans_decode_with_indexes(t._p(self.buf), self.buf.size, t._p(self.state), t._p(indexes), indexes.size,
                                                t._p(t.cdf), t.cdf.shape[1], t._p(t.sizes), t._p(t.offsets), t.cdf.shape[0],
                                                t._p(out)), "rans_decode_with_indexes")
        return out


def _std_cumulative(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * torch.erfc(-(2 ** -0.5) * x)


def gaussian_table(scale_table: torch.Tensor) -> CdfTable:
    """GaussianConditional.update_scale_table / update: one CDF per scale level, support +- ceil(6.1 scale)"""
    st = scale_table.double().cpu()
    center = torch.ceil(st * TAIL_MULTIPLIER).long()
    length = 2 * center + 1
    max_len = int(length.max())
    samples = torch.abs(torch.arange(max_len).double()[None, :] - center[:, None].double())
    s = st[:, None]
    upper, lower = _std_cumulative((0.5 - samples) / s), _std_cumulative((-0.5 - samples) / s)
    pmf = (upper - lower).numpy()
    tail = (2 * lower[:, 0]).numpy()
    return CdfTable(pmf, tail, length.tolist(), (-center).tolist())


def build_indexes(scales: torch.Tensor, scale_table: torch.Tensor) -> torch.Tensor:
    """GaussianConditional.build_indexes: index of the first table scale >= max(scale, lower bound)"""
    table = scale_table.to(scales.device)
    s = torch.clamp_min(scales, SCALES_MIN)
    idx = torch.full(s.shape, len(table) - 1, dtype=torch.int32, device=scales.device)
    for v in table[:-1].tolist():
        idx -= (s <= v).to(torch.int32)
    return idx
