"""Deterministic synthetic weights / audio for parity tests and the bench.

Everything here is plain numpy and reproducible bit-for-bit on any host: values come from a
counter-based splitmix64 hash, not from torch's or numpy's stateful generators, so the golden
fixtures under ``tests/golden`` (made in the build container by ``oracle/make_golden.py`` from the
live reference) can be re-derived on the GPU box, where ``/root/reference`` does not exist.

The key layout produced by :func:`make_state_dict` is the reference's ``state_dict`` layout
(reference ``encodec/modules/seanet.py:92-146,176-253``,
``encodec/modules/conv.py:109-163``, ``encodec/quantization/core_vq.py:128-135``).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _stream_key(seed: int, name: str) -> np.uint64:
    h = zlib.crc32(name.encode("utf-8")) & 0xFFFFFFFF
    k = (int(seed) * 0x100000001B3 + h * 0x9E3779B1 + 0x632BE59BD9B4E019) & 0xFFFFFFFFFFFFFFFF
    return _splitmix64(np.array([k], dtype=np.uint64))[0]


def hash_uniform(seed: int, name: str, n: int, offset: int = 0) -> np.ndarray:
    """``n`` float64 values in [0, 1), a pure function of (seed, name, offset + i)."""
    key = _stream_key(seed, name)
    with np.errstate(over="ignore"):
        idx = (np.arange(offset, offset + n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95)) ^ key
    z = _splitmix64(idx)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def hash_symmetric(seed: int, name: str, shape, bound: float) -> np.ndarray:
    """float32 tensor ~ U(-bound, bound)."""
    n = int(np.prod(shape))
    u = hash_uniform(seed, name, n)
    return ((2.0 * u - 1.0) * bound).astype(np.float32).reshape(shape)


def hash_normal(seed: int, name: str, shape, offset: int = 0) -> np.ndarray:
    """float32 tensor ~ N(0, 1) (Box-Muller on two hashed uniforms per value)."""
    n = int(np.prod(shape))
    u1 = hash_uniform(seed, name + "/a", n, offset)
    u2 = hash_uniform(seed, name + "/b", n, offset)
    r = np.sqrt(-2.0 * np.log1p(-u1))  # log(1-u1): u1 in [0,1) so the argument is in (0,1]
    return (r * np.cos(2.0 * math.pi * u2)).astype(np.float32).reshape(shape)


@dataclass
class CodecSpec:
    """Hyper-parameters of one EnCodec variant (reference ``model.py:286-326``)."""

    sample_rate: int = 24000
    channels: int = 1
    causal: bool = True
    norm: str = "weight_norm"  # or "time_group_norm" (48 kHz model) or "layer_norm" (the fork's 10 Hz configs)
    normalize: bool = False
    segment: Optional[float] = None
    overlap: float = 0.01
    ratios: List[int] = field(default_factory=lambda: [8, 5, 4, 2])
    n_filters: int = 32
    dimension: int = 128
    bins: int = 1024
    target_bandwidths: List[float] = field(default_factory=lambda: [1.5, 3.0, 6.0, 12.0, 24.0])
    kernel_size: int = 7
    last_kernel_size: int = 7
    residual_kernel_size: int = 3
    compress: int = 2
    lstm: int = 2

    @property
    def hop_length(self) -> int:
        return int(np.prod(self.ratios))

    @property
    def frame_rate(self) -> int:
        return math.ceil(self.sample_rate / self.hop_length)

    @property
    def n_q(self) -> int:
        # reference model.py:302
        return int(1000 * self.target_bandwidths[-1] // (math.ceil(self.sample_rate / self.hop_length) * 10))

    @property
    def segment_length(self) -> Optional[int]:
        return None if self.segment is None else int(self.segment * self.sample_rate)

    @property
    def segment_stride(self) -> Optional[int]:
        sl = self.segment_length
        return None if sl is None else max(1, int((1 - self.overlap) * sl))

    def n_q_for_bandwidth(self, bandwidth: Optional[float]) -> int:
        # reference quantization/vq.py:116-131
        bw_per_q = math.log2(self.bins) * self.frame_rate
        n_q = self.n_q
        if bandwidth and bandwidth > 0.0:
            n_q = int(max(1, math.floor(bandwidth * 1000 / bw_per_q)))
        return n_q


def spec_24khz() -> CodecSpec:
    return CodecSpec()


def spec_48khz() -> CodecSpec:
    return CodecSpec(sample_rate=48000, channels=2, causal=False, norm="time_group_norm", normalize=True,
                     segment=1.0, target_bandwidths=[3.0, 6.0, 12.0, 24.0])


def spec_fork10hz(ratios=(6, 5, 5, 2, 1)) -> CodecSpec:
    """The fork's own training configuration (reference ``params/091224_l1.yaml:65-92``): 10 Hz mono signals, causal,
    ``norm='layer_norm'`` (ConvLayerNorm), ratios incl. a stride-1 stage, dimension 256, 1024 bins, 0.08 kbps (n_q 8)."""
    return CodecSpec(sample_rate=10, channels=1, causal=True, norm="layer_norm", ratios=list(ratios), dimension=256,
                     bins=1024, target_bandwidths=[0.08])


def conv_layout(spec: CodecSpec):
    """List of (prefix, kind, c_in, c_out, k, stride) in reference module order.

    kind in {"conv", "convtr", "lstm"}; "conv_plain" marks decoder.model.<last> (norm='none', D7).
    """
    out = []
    nf = spec.n_filters
    rk = spec.residual_kernel_size

    def resblock(prefix, dim):
        hidden = dim // spec.compress
        out.append((f"{prefix}.block.1", "conv", dim, hidden, rk, 1))
        out.append((f"{prefix}.block.3", "conv", hidden, dim, 1, 1))
        out.append((f"{prefix}.shortcut", "conv", dim, dim, 1, 1))

    # encoder (ratios reversed, seanet.py:102)
    mult = 1
    out.append(("encoder.model.0", "conv", spec.channels, nf, spec.kernel_size, 1))
    idx = 1
    for ratio in reversed(spec.ratios):
        resblock(f"encoder.model.{idx}", mult * nf)
        out.append((f"encoder.model.{idx + 2}", "conv", mult * nf, mult * nf * 2, 2 * ratio, ratio))
        idx += 3
        mult *= 2
    if spec.lstm:
        out.append((f"encoder.model.{idx}", "lstm", mult * nf, mult * nf, spec.lstm, 0))
        idx += 1
    out.append((f"encoder.model.{idx + 1}", "conv", mult * nf, spec.dimension, spec.last_kernel_size, 1))
    # decoder
    mult = 2 ** len(spec.ratios)
    out.append(("decoder.model.0", "conv", spec.dimension, mult * nf, spec.kernel_size, 1))
    idx = 1
    if spec.lstm:
        out.append((f"decoder.model.{idx}", "lstm", mult * nf, mult * nf, spec.lstm, 0))
        idx += 1
    for ratio in spec.ratios:
        out.append((f"decoder.model.{idx + 1}", "convtr", mult * nf, mult * nf // 2, 2 * ratio, ratio))
        resblock(f"decoder.model.{idx + 2}", mult * nf // 2)
        idx += 3
        mult //= 2
    out.append((f"decoder.model.{idx + 1}", "conv_plain", nf, spec.channels, spec.last_kernel_size, 1))
    return out


def make_state_dict(spec: CodecSpec, seed: int = 0, codebooks: Optional[np.ndarray] = None,
                    shared_codebook: bool = False) -> Dict[str, np.ndarray]:
    """Random-init weights in the reference state_dict layout (float32 numpy arrays).

    Magnitudes mimic torch's default initialisers so that activations are in a realistic range;
    weight-norm gains and GroupNorm affines are perturbed so that folding them is really tested.
    ``codebooks`` is ``[n_q, bins, D]`` (or ``[bins, D]`` broadcast to every layer); if None a
    plain N(0, 0.05) codebook is used. ``shared_codebook`` reproduces the fork's aliasing (D6).
    """
    sd: Dict[str, np.ndarray] = {}
    wn = spec.norm == "weight_norm"
    for prefix, kind, c_in, c_out, k, stride in conv_layout(spec):
        if kind == "lstm":
            h = c_out
            bound = 1.0 / math.sqrt(h)
            for layer in range(k):
                for nm, shape in (("weight_ih", (4 * h, c_in if layer == 0 else h)), ("weight_hh", (4 * h, h)),
                                  ("bias_ih", (4 * h,)), ("bias_hh", (4 * h,))):
                    key = f"{prefix}.lstm.{nm}_l{layer}"
                    sd[key] = hash_symmetric(seed, key, shape, bound)
            continue
        if kind == "convtr":
            wshape = (c_in, c_out, k)
            fan_in = c_out * k  # torch computes fan_in from dim 1 of the stored weight
            base = f"{prefix}.convtr.convtr"
            normp = f"{prefix}.convtr.norm"
        else:
            wshape = (c_out, c_in, k)
            fan_in = c_in * k
            base = f"{prefix}.conv.conv"
            normp = f"{prefix}.conv.norm"
        bound = 1.0 / math.sqrt(fan_in)
        w = hash_symmetric(seed, base + ".w", wshape, bound)
        b = hash_symmetric(seed, base + ".bias", (c_out,), bound)
        if kind == "conv_plain" or not wn:
            sd[base + ".weight"] = w
            sd[base + ".bias"] = b
            if kind != "conv_plain":
                sd[normp + ".weight"] = (1.0 + hash_symmetric(seed, normp + ".weight", (c_out,), 0.3)).astype(np.float32)
                sd[normp + ".bias"] = hash_symmetric(seed, normp + ".bias", (c_out,), 0.1)
        else:
            nrm = np.sqrt((w.astype(np.float64) ** 2).sum(axis=(1, 2), keepdims=True))
            g = (nrm * (1.0 + hash_symmetric(seed, base + ".g", (wshape[0], 1, 1), 0.2))).astype(np.float32)
            sd[base + ".bias"] = b
            sd[base + ".weight_g"] = g
            sd[base + ".weight_v"] = w
    n_q = spec.n_q
    if codebooks is None:
        codebooks = 0.05 * hash_normal(seed, "codebook", (1 if shared_codebook else n_q, spec.bins, spec.dimension))
    codebooks = np.asarray(codebooks, dtype=np.float32)
    if codebooks.ndim == 2:
        codebooks = codebooks[None]
    for i in range(n_q):
        e = codebooks[0 if (shared_codebook or codebooks.shape[0] == 1) else i]
        p = f"quantizer.vq.layers.{i}._codebook"
        sd[p + ".inited"] = np.ones((1,), dtype=np.float32)
        sd[p + ".cluster_size"] = np.zeros((spec.bins,), dtype=np.float32)
        sd[p + ".embed"] = e
        sd[p + ".embed_avg"] = e.copy()
    return sd


def make_audio(seed: int, batch: int, channels: int, length: int) -> np.ndarray:
    """Synthetic audio: clipped Gaussian noise plus three sinusoids per item (float32 [B, C, T])."""
    x = 0.3 * hash_normal(seed, "audio", (batch, channels, length))
    t = np.arange(length, dtype=np.float64)
    for b in range(batch):
        for c in range(channels):
            f = hash_uniform(seed, f"audio/f{b}.{c}", 3)
            for j in range(3):
                x[b, c] += (0.15 * np.sin(2 * math.pi * (0.0005 + 0.02 * f[j]) * t + 6.28 * f[(j + 1) % 3])).astype(np.float32)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def calibrated_codebooks(seed: int, mean_vec: np.ndarray, scales: np.ndarray, bins: int) -> np.ndarray:
    """Codebooks matched to a model's residual statistics (SURVEY.md section 8c recipe, made portable).

    ``E_i = (i == 0) * mean_vec + scales[i] * N(0, 1)``. ``mean_vec`` ([D]) and ``scales`` ([n_q]) are
    *stored* calibration data (tests/golden), never recomputed, so every implementation sees the
    same codebooks bit-for-bit.
    """
    n_q = len(scales)
    d = mean_vec.shape[0]
    cb = hash_normal(seed, "calib-codebook", (n_q, bins, d)) * np.asarray(scales, np.float32)[:, None, None]
    cb[0] += mean_vec.astype(np.float32)[None, :]
    return cb.astype(np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# Language model of the entropy-coded .ecdc stream (reference model.py:45-83, modules/transformer.py:62-119)
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class LMSpec:
    """Constructor arguments of the reference's ``LMModel`` as ``EncodecModel.get_lm_model`` passes them (model.py:268-269)."""
    n_q: int = 32
    card: int = 1024
    dim: int = 200
    num_layers: int = 5
    num_heads: int = 8
    hidden_scale: float = 4.0
    past_context: int = 262           # int(3.5 * frame_rate): 262 at 75 Hz, 525 at 150 Hz
    max_period: float = 10000.0

    @property
    def hidden(self) -> int:
        return int(self.dim * self.hidden_scale)


def make_lm_state_dict(spec: LMSpec, seed: int = 0, logit_gain: float = 2.5) -> Dict[str, np.ndarray]:
    """Deterministic weights with the key layout of the reference's ``LMModel.state_dict()``. ``logit_gain`` sets the
    spread of the output logits (peaked distributions exercise the coder's minimum-range clamp)."""
    sd: Dict[str, np.ndarray] = {}
    d, h = spec.dim, spec.hidden

    def lin(name, n_out, n_in, gain=1.0):
        sd[name + ".weight"] = hash_symmetric(seed, name + ".weight", (n_out, n_in), gain * math.sqrt(3.0 / n_in))
        sd[name + ".bias"] = hash_symmetric(seed, name + ".bias", (n_out,), 0.1)

    def norm(name):
        sd[name + ".weight"] = (1.0 + hash_symmetric(seed, name + ".weight", (d,), 0.2)).astype(np.float32)
        sd[name + ".bias"] = hash_symmetric(seed, name + ".bias", (d,), 0.1)

    norm("transformer.norm_in")
    for i in range(spec.num_layers):
        p = f"transformer.layers.{i}"
        sd[p + ".self_attn.in_proj_weight"] = hash_symmetric(seed, p + ".in_proj_weight", (3 * d, d), 1.5 * math.sqrt(3.0 / d))
        sd[p + ".self_attn.in_proj_bias"] = hash_symmetric(seed, p + ".in_proj_bias", (3 * d,), 0.1)
        lin(p + ".self_attn.out_proj", d, d)
        lin(p + ".linear1", h, d)
        lin(p + ".linear2", d, h)
        norm(p + ".norm1")
        norm(p + ".norm2")
    for k in range(spec.n_q):
        sd[f"emb.{k}.weight"] = hash_normal(seed, f"emb.{k}.weight", (spec.card + 1, d))
        lin(f"linears.{k}", spec.card, d, gain=logit_gain)
    return sd
