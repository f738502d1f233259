"""CPU restatement of the entropy-coded `.ecdc` path -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Restates, citing /root/reference/encodec:
  * ``LMModel.forward`` (model.py:65-83) over ``StreamingTransformerEncoder`` (modules/transformer.py:62-119) and
    ``StreamingTransformerEncoderLayer`` (transformer.py:30-59; ``nn.TransformerEncoderLayer`` defaults: post-norm,
    exact GELU, eps 1e-5) driven one time step at a time as compress.py:69-78 / :131-152 do;
  * ``build_stable_quantized_cdf`` (quantization/ac.py:18-53) in the float32 arithmetic of the torch CPU kernels;
  * ``ArithmeticCoder`` / ``ArithmeticDecoder`` (ac.py:56-260) with Python integers, 1-bit BitPacker order
    (binary.py:55-89: least-significant bit of every byte first).

Pinned by tests/golden/lm_ac.npz, which oracle/make_golden_lm.py produced by importing the UNMODIFIED reference.
Floating point: the LM is fp32 in the reference; this restatement evaluates it in float64 (or float32 on request) and is
compared within a tolerance; everything downstream of a given pdf is integer / exactly rounded and compared bit for bit.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# language model
# ---------------------------------------------------------------------------------------------------------------------
def _layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * w + b


def _gelu(x):
    erf = np.vectorize(math.erf, otypes=[np.float64])
    return (0.5 * x * (1.0 + erf(x.astype(np.float64) / math.sqrt(2.0)))).astype(x.dtype)


def sin_embedding(positions: np.ndarray, dim: int, max_period: float, dtype) -> np.ndarray:
    """create_sin_embedding (transformer.py:16-27): [T] -> [T, dim] = cat(cos, sin)."""
    half = dim // 2
    adim = np.arange(half, dtype=np.float64)
    phase = positions.astype(np.float64)[:, None] / (max_period ** (adim / (half - 1)))[None, :]
    return np.concatenate([np.cos(phase), np.sin(phase)], axis=-1).astype(dtype)


def lm_probas(sd: Dict[str, np.ndarray], codes: np.ndarray, *, num_layers: int, num_heads: int, past_context: int,
              max_period: float = 10000.0, dtype=np.float64) -> np.ndarray:
    """Probabilities the reference's streaming loop sees for one frame: codes [K, T] -> probas [T, K, card].

    Step t is fed ``1 + codes[:, t-1]`` (zeros at t = 0, compress.py:69,78). The streaming state of a layer is its INPUT
    history, initialised with ONE all-zero row that is attended like a real position until the window of
    ``past_context`` rows pushes it out (transformer.py:103-104,116-117, mask :52-53 never excludes a stored row).
    Attention over the stored rows + the current one is what a causal windowed attention over rows
    [t - min(t + 1, past_context), t] computes, row -1 being the zero row.
    """
    K, T = codes.shape
    f = lambda name: sd[name].astype(dtype)
    dim = sd["transformer.norm_in.weight"].shape[0]
    hd = dim // num_heads
    idx = np.zeros((K, T), dtype=np.int64)
    idx[:, 1:] = 1 + codes[:, :-1]
    x = np.zeros((T, dim), dtype=dtype)
    for k in range(K):                                                   # model.py:79
        x += f(f"emb.{k}.weight")[idx[k]]
    x = _layer_norm(x, f("transformer.norm_in.weight"), f("transformer.norm_in.bias"))      # transformer.py:110
    x = x + sin_embedding(np.arange(T), dim, max_period, dtype)          # :106-111
    for i in range(num_layers):
        p = f"transformer.layers.{i}"
        w_in, b_in = f(p + ".self_attn.in_proj_weight"), f(p + ".self_attn.in_proj_bias")
        hist = np.concatenate([np.zeros((1, dim), dtype=dtype), x], axis=0)    # row j = position j - 1
        q = x @ w_in[:dim].T + b_in[:dim]
        kk = hist @ w_in[dim:2 * dim].T + b_in[dim:2 * dim]
        vv = hist @ w_in[2 * dim:].T + b_in[2 * dim:]
        att = np.zeros_like(x)
        for t in range(T):
            n_past = min(t + 1, past_context)
            lo = t + 1 - n_past                                          # first history row in the window
            for h in range(num_heads):
                sl = slice(h * hd, (h + 1) * hd)
                s = (kk[lo:t + 2, sl] @ q[t, sl]) / math.sqrt(hd)
                s = np.exp(s - s.max())
                att[t, sl] = (s / s.sum()) @ vv[lo:t + 2, sl]
        sa = att @ f(p + ".self_attn.out_proj.weight").T + f(p + ".self_attn.out_proj.bias")
        x = _layer_norm(x + sa, f(p + ".norm1.weight"), f(p + ".norm1.bias"))               # transformer.py:38
        ff = _gelu(x @ f(p + ".linear1.weight").T + f(p + ".linear1.bias")) @ f(p + ".linear2.weight").T + f(p + ".linear2.bias")
        x = _layer_norm(x + ff, f(p + ".norm2.weight"), f(p + ".norm2.bias"))               # :39
    out = np.zeros((T, K, sd["linears.0.bias"].shape[0]), dtype=dtype)
    for k in range(K):                                                   # model.py:81-82
        logits = x @ f(f"linears.{k}.weight").T + f(f"linears.{k}.bias")
        e = np.exp(logits - logits.max(axis=-1, keepdims=True))
        out[:, k] = e / e.sum(axis=-1, keepdims=True)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# pdf -> quantised cdf
# ---------------------------------------------------------------------------------------------------------------------
def build_stable_quantized_cdf(pdf: np.ndarray, total_range_bits: int = 24, roundoff: float = 1e-8,
                               min_range: int = 2) -> np.ndarray:
    """ac.py:18-53 with ``check=False`` (how compress.py calls it): float32 pdf [..., N] -> int64 cdf [..., N]."""
    pdf = pdf.astype(np.float32)
    r = np.float32(roundoff)
    pdf = np.floor(pdf / r) * r                                          # ac.py:37-38 (float32 tensor op scalar)
    total_range = 2 ** total_range_bits
    card = pdf.shape[-1]
    alpha = min_range * card / total_range
    assert alpha <= 1
    scale = np.float32((1 - alpha) * total_range)
    ranges = np.floor(scale * pdf).astype(np.int64) + min_range          # :44-45
    return np.cumsum(ranges, axis=-1)


# ---------------------------------------------------------------------------------------------------------------------
# arithmetic coder
# ---------------------------------------------------------------------------------------------------------------------
class _BitWriter:
    """BitPacker(bits=1) (binary.py:55-89): bit i of the stream is bit (i % 8) of byte i // 8."""
    def __init__(self):
        self.out = bytearray()
        self.cur = 0
        self.n = 0

    def push(self, bit: int):
        self.cur |= (bit & 1) << self.n
        self.n += 1
        if self.n == 8:
            self.out.append(self.cur)
            self.cur, self.n = 0, 0

    def flush(self):
        if self.n:
            self.out.append(self.cur)
            self.cur, self.n = 0, 0


class ArithmeticCoder:
    def __init__(self, total_range_bits: int = 24):
        self.bits = total_range_bits
        self.w = _BitWriter()
        self.low = 0
        self.high = 0
        self.max_bit = -1

    def push(self, symbol: int, cdf) -> None:                           # ac.py:127-157
        while self.high - self.low + 1 < 2 ** self.bits:
            self.low *= 2
            self.high = self.high * 2 + 1
            self.max_bit += 1
        range_low = 0 if symbol == 0 else int(cdf[symbol - 1])
        range_high = int(cdf[symbol]) - 1
        ratio = (self.high - self.low + 1) / (2 ** self.bits)
        eff_low = int(math.ceil(range_low * ratio))
        eff_high = int(math.floor(range_high * ratio))
        self.high = self.low + eff_high
        self.low = self.low + eff_low
        assert self.low <= self.high
        while self.max_bit >= 0:                                         # _flush_common_prefix, ac.py:109-125
            b1 = self.low >> self.max_bit
            b2 = self.high >> self.max_bit
            if b1 != b2:
                break
            self.low -= b1 << self.max_bit
            self.high -= b1 << self.max_bit
            self.max_bit -= 1
            self.w.push(b1)
        assert self.max_bit <= 61

    def finish(self) -> bytes:                                           # flush, ac.py:159-166
        while self.max_bit >= 0:
            self.w.push((self.low >> self.max_bit) & 1)
            self.max_bit -= 1
        self.w.flush()
        return bytes(self.w.out)


class ArithmeticDecoder:
    def __init__(self, data: bytes, total_range_bits: int = 24):
        self.bits = total_range_bits
        self.data = data
        self.pos = 0                                                     # in bits
        self.low = 0
        self.high = 0
        self.current = 0
        self.max_bit = -1

    @property
    def bytes_consumed(self) -> int:
        return (self.pos + 7) // 8

    def pull(self, cdf) -> Optional[int]:                                # ac.py:214-260
        while self.high - self.low + 1 < 2 ** self.bits:
            if self.pos >= 8 * len(self.data):
                return None
            bit = (self.data[self.pos >> 3] >> (self.pos & 7)) & 1
            self.pos += 1
            self.low *= 2
            self.high = self.high * 2 + 1
            self.current = self.current * 2 + bit
            self.max_bit += 1
        ratio = (self.high - self.low + 1) / (2 ** self.bits)
        lo_idx, hi_idx = 0, len(cdf) - 1
        while True:
            if hi_idx < lo_idx:
                raise RuntimeError("Binary search failed")
            mid = (lo_idx + hi_idx) // 2
            range_low = int(cdf[mid - 1]) if mid > 0 else 0
            range_high = int(cdf[mid]) - 1
            low = int(math.ceil(range_low * ratio)) + self.low
            high = int(math.floor(range_high * ratio)) + self.low
            if self.current >= low:
                if self.current <= high:
                    break
                lo_idx = mid + 1
            else:
                hi_idx = mid - 1
        self.low, self.high = low, high
        while self.max_bit >= 0:                                         # ac.py:195-212
            b1 = self.low >> self.max_bit
            b2 = self.high >> self.max_bit
            if b1 != b2:
                break
            self.low -= b1 << self.max_bit
            self.high -= b1 << self.max_bit
            self.current -= b1 << self.max_bit
            self.max_bit -= 1
        return mid


def encode_frame(codes: np.ndarray, cdfs: np.ndarray, total_range_bits: int = 24) -> bytes:
    """compress.py:66-87 for one frame with use_lm: codes [K, T], cdfs [T, K, card] -> the bytes of that frame."""
    coder = ArithmeticCoder(total_range_bits)
    K, T = codes.shape
    for t in range(T):
        for k in range(K):
            coder.push(int(codes[k, t]), cdfs[t, k])
    return coder.finish()
