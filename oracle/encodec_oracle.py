"""CPU oracle for the EnCodec codec forward pass -- TEST INFRASTRUCTURE ONLY.

A framework-free (numpy) restatement of the reference's algorithm for the hot path
``EncodecModel.encode / decode / forward``. Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
this file; the product package ``encodec_b200`` never does (it fails loudly without its CUDA library).

Pinning: the reference ships NO golden vectors or known-answer tests for this path (SURVEY.md
section 8c), so the oracle is pinned against outputs of the reference itself, produced in the build
container by ``oracle/make_golden.py`` (imports ``/root/reference`` unmodified) and committed under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the oracle against them everywhere.

Each function cites the reference lines it restates (paths relative to ``/root/reference/encodec``).
All arithmetic runs in ``dtype`` (float32 mirrors the reference; float64 gives a tighter truth for
tolerance tests). Tensors use the reference's boundary layout: ``[B, C, T]``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------
# elementary ops
# ----------------------------------------------------------------------------------------------
def elu(x: np.ndarray) -> np.ndarray:
    """nn.ELU(alpha=1.0) -- modules/seanet.py:43,50,126,138,207,226."""
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype)


def fold_weight_norm(g: np.ndarray, v: np.ndarray) -> np.ndarray:
    """torch.nn.utils.weight_norm, dim=0 -- modules/conv.py:28-29.

    ``w = v * (g / ||v||)`` with the norm over every dim but 0. For ConvTranspose1d dim 0 is C_in.
    """
    nrm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True)).astype(v.dtype)
    return (v * (g / nrm)).astype(v.dtype)


def get_extra_padding_for_conv1d(length: int, kernel_size: int, stride: int, padding_total: int) -> int:
    """modules/conv.py:55-62."""
    n_frames = (length - kernel_size + padding_total) / stride + 1
    ideal_length = (math.ceil(n_frames) - 1) * stride + (kernel_size - padding_total)
    return ideal_length - length


def pad1d_reflect(x: np.ndarray, left: int, right: int) -> np.ndarray:
    """modules/conv.py:80-97 (mode='reflect', including the zero-extension for short inputs)."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, extra)])
    padded = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(left, right)], mode="reflect")
    end = padded.shape[-1] - extra
    return padded[..., :end]


def conv1d(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int) -> np.ndarray:
    """nn.Conv1d (no padding, dilation 1) -- modules/conv.py:116,121. x [B,Ci,T], w [Co,Ci,K]."""
    bsz, ci, t = x.shape
    co, _, k = w.shape
    t_out = (t - k) // stride + 1
    # im2col: cols[b, ci, k, t_out]
    idx = np.arange(t_out)[None, :] * stride + np.arange(k)[:, None]  # [K, T_out]
    cols = x[:, :, idx]  # [B, Ci, K, T_out]
    y = np.einsum("ock,bckt->bot", w, cols, optimize=True).astype(x.dtype)
    if b is not None:
        y = y + b[None, :, None]
    return y.astype(x.dtype)


def conv_transpose1d(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int) -> np.ndarray:
    """nn.ConvTranspose1d -- modules/conv.py:156,161. x [B,Ci,L], w [Ci,Co,K] -> [B,Co,(L-1)s+K]."""
    bsz, ci, length = x.shape
    _, co, k = w.shape
    t_out = (length - 1) * stride + k
    y = np.zeros((bsz, co, t_out), dtype=x.dtype)
    contrib = np.einsum("bil,iok->bokl", x, w, optimize=True).astype(x.dtype)  # [B, Co, K, L]
    for kk in range(k):
        y[:, :, kk: kk + (length - 1) * stride + 1: stride] += contrib[:, :, kk, :]
    if b is not None:
        y = y + b[None, :, None]
    return y.astype(x.dtype)


def group_norm1(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.GroupNorm(1, C) -- modules/conv.py:50: per-sample statistics over (C, T), biased variance."""
    xd = x.astype(np.float64)
    mean = xd.mean(axis=(1, 2), keepdims=True)
    var = xd.var(axis=(1, 2), keepdims=True)
    y = (xd - mean) / np.sqrt(var + eps)
    return (y * gamma[None, :, None] + beta[None, :, None]).astype(x.dtype)


def layer_norm_c(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """ConvLayerNorm -- modules/norm.py:16-30 (used by modules/conv.py:44-46 for norm='layer_norm'): nn.LayerNorm over the
    channel axis of every time step separately (``b c t -> b t c``, normalise, back), biased variance, eps 1e-5."""
    xd = x.astype(np.float64)
    mean = xd.mean(axis=1, keepdims=True)
    var = xd.var(axis=1, keepdims=True)
    y = (xd - mean) / np.sqrt(var + eps)
    return (y * gamma[None, :, None] + beta[None, :, None]).astype(x.dtype)


def apply_norm(y: np.ndarray, gamma, beta, kind: str) -> np.ndarray:
    """get_norm_module -- modules/conv.py:38-52: Identity (weight_norm / none), GroupNorm(1, C) or ConvLayerNorm."""
    if gamma is None:
        return y
    return layer_norm_c(y, gamma, beta) if kind == "layer_norm" else group_norm1(y, gamma, beta)


# ----------------------------------------------------------------------------------------------
# parameter access in the reference state_dict layout
# ----------------------------------------------------------------------------------------------
class Params:
    """Folds weight-norm once and hands out (w, b, gamma, beta) per conv prefix."""

    def __init__(self, sd: Dict[str, np.ndarray], dtype=np.float32, norm: str = "time_group_norm"):
        self.sd = {k: np.asarray(v) for k, v in sd.items()}
        self.dtype = dtype
        self.norm = norm   # which module a '<conv>.norm.weight' entry belongs to (only matters when such entries exist)

    def conv(self, prefix: str, transposed: bool = False):
        base = f"{prefix}.convtr.convtr" if transposed else f"{prefix}.conv.conv"
        normp = f"{prefix}.convtr.norm" if transposed else f"{prefix}.conv.norm"
        sd = self.sd
        if base + ".weight_g" in sd:
            w = fold_weight_norm(sd[base + ".weight_g"].astype(np.float32), sd[base + ".weight_v"].astype(np.float32))
        else:
            w = sd[base + ".weight"]
        b = sd[base + ".bias"]
        gamma = sd.get(normp + ".weight")
        beta = sd.get(normp + ".bias")
        cast = lambda a: None if a is None else a.astype(self.dtype)
        return cast(w), cast(b), cast(gamma), cast(beta)

    def lstm(self, prefix: str, layer: int):
        sd = self.sd
        return tuple(sd[f"{prefix}.lstm.{n}_l{layer}"].astype(self.dtype)
                     for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"))


# ----------------------------------------------------------------------------------------------
# SConv1d / SConvTranspose1d / SLSTM / ResnetBlock
# ----------------------------------------------------------------------------------------------
def sconv1d(x: np.ndarray, p: Params, prefix: str, stride: int, causal: bool) -> np.ndarray:
    """SConv1d.forward -- modules/conv.py:202-221 (+ NormConv1d :120-128)."""
    w, b, gamma, beta = p.conv(prefix)
    k = w.shape[-1]
    padding_total = k - stride
    extra = get_extra_padding_for_conv1d(x.shape[-1], k, stride, padding_total)
    if causal:
        xp = pad1d_reflect(x, padding_total, extra)
    else:
        pr = padding_total // 2
        pl = padding_total - pr
        xp = pad1d_reflect(x, pl, pr + extra)
    y = conv1d(xp, w, b, stride)
    return apply_norm(y, gamma, beta, p.norm)


def sconvtr1d(x: np.ndarray, p: Params, prefix: str, stride: int, causal: bool,
              trim_right_ratio: float = 1.0) -> np.ndarray:
    """SConvTranspose1d.forward -- modules/conv.py:241-263: convtr -> norm (untrimmed) -> trim."""
    w, b, gamma, beta = p.conv(prefix, transposed=True)
    k = w.shape[-1]
    padding_total = k - stride
    y = conv_transpose1d(x, w, b, stride)
    y = apply_norm(y, gamma, beta, p.norm)
    if causal:
        pr = math.ceil(padding_total * trim_right_ratio)
        pl = padding_total - pr
    else:
        pr = padding_total // 2
        pl = padding_total - pr
    return y[..., pl: y.shape[-1] - pr]


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def slstm(x: np.ndarray, p: Params, prefix: str, num_layers: int) -> np.ndarray:
    """SLSTM.forward -- modules/lstm.py:22-28: nn.LSTM (gates i,f,g,o; zero state) + skip."""
    inp = np.transpose(x, (2, 0, 1))  # [T, B, C]
    seq = inp
    for layer in range(num_layers):
        w_ih, w_hh, b_ih, b_hh = p.lstm(prefix, layer)
        hdim = w_hh.shape[1]
        t_len, bsz, _ = seq.shape
        pre = (seq.reshape(t_len * bsz, -1) @ w_ih.T + b_ih).reshape(t_len, bsz, 4 * hdim)
        h = np.zeros((bsz, hdim), dtype=x.dtype)
        c = np.zeros((bsz, hdim), dtype=x.dtype)
        out = np.empty((t_len, bsz, hdim), dtype=x.dtype)
        for t in range(t_len):
            gates = pre[t] + (h @ w_hh.T + b_hh)
            i = _sigmoid(gates[:, 0 * hdim:1 * hdim])
            f = _sigmoid(gates[:, 1 * hdim:2 * hdim])
            g = np.tanh(gates[:, 2 * hdim:3 * hdim])
            o = _sigmoid(gates[:, 3 * hdim:4 * hdim])
            c = (f * c + i * g).astype(x.dtype)
            h = (o * np.tanh(c)).astype(x.dtype)
            out[t] = h
        seq = out
    y = seq + inp
    return np.ascontiguousarray(np.transpose(y, (1, 2, 0)))


def resnet_block(x: np.ndarray, p: Params, prefix: str, causal: bool) -> np.ndarray:
    """SEANetResnetBlock.forward -- modules/seanet.py:37-64 (true_skip=False => conv shortcut)."""
    h = sconv1d(elu(x), p, f"{prefix}.block.1", 1, causal)
    h = sconv1d(elu(h), p, f"{prefix}.block.3", 1, causal)
    return sconv1d(x, p, f"{prefix}.shortcut", 1, causal) + h


# ----------------------------------------------------------------------------------------------
# SEANet encoder / decoder
# ----------------------------------------------------------------------------------------------
def seanet_encoder(x: np.ndarray, p: Params, spec, taps: Optional[dict] = None) -> np.ndarray:
    """SEANetEncoder.forward -- modules/seanet.py:92-146."""
    causal = spec.causal
    y = sconv1d(x, p, "encoder.model.0", 1, causal)
    idx = 1
    for ratio in reversed(spec.ratios):
        y = resnet_block(y, p, f"encoder.model.{idx}", causal)
        if taps is not None:
            taps[f"encoder.model.{idx}"] = y
        y = sconv1d(elu(y), p, f"encoder.model.{idx + 2}", ratio, causal)
        if taps is not None:
            taps[f"encoder.model.{idx + 2}"] = y
        idx += 3
    if spec.lstm:
        y = slstm(y, p, f"encoder.model.{idx}", spec.lstm)
        if taps is not None:
            taps[f"encoder.model.{idx}"] = y
        idx += 1
    y = sconv1d(elu(y), p, f"encoder.model.{idx + 1}", 1, causal)
    return y


def seanet_decoder(z: np.ndarray, p: Params, spec, taps: Optional[dict] = None) -> np.ndarray:
    """SEANetDecoder.forward -- modules/seanet.py:176-253 (last conv has norm='none', :227-228)."""
    causal = spec.causal
    y = sconv1d(z, p, "decoder.model.0", 1, causal)
    idx = 1
    if spec.lstm:
        y = slstm(y, p, f"decoder.model.{idx}", spec.lstm)
        if taps is not None:
            taps[f"decoder.model.{idx}"] = y
        idx += 1
    for ratio in spec.ratios:
        y = sconvtr1d(elu(y), p, f"decoder.model.{idx + 1}", ratio, causal)
        if taps is not None:
            taps[f"decoder.model.{idx + 1}"] = y
        y = resnet_block(y, p, f"decoder.model.{idx + 2}", causal)
        if taps is not None:
            taps[f"decoder.model.{idx + 2}"] = y
        idx += 3
    y = sconv1d(elu(y), p, f"decoder.model.{idx + 1}", 1, causal)
    return y


# ----------------------------------------------------------------------------------------------
# residual vector quantiser
# ----------------------------------------------------------------------------------------------
def codebook_quantize(x: np.ndarray, embed: np.ndarray, chunk: int = 8192) -> np.ndarray:
    """EuclideanCodebook.quantize -- quantization/core_vq.py:178-194. x [N,D], embed [bins,D] -> int64 [N].

    ``dist = -(|x|^2 - 2 x E^T + |E|^2)`` in that association order; ``argmax`` = first maximum.
    """
    et = np.ascontiguousarray(embed.T)
    e2 = (embed ** 2).sum(axis=1)[None, :].astype(x.dtype)
    out = np.empty((x.shape[0],), dtype=np.int64)
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        x2 = (xs ** 2).sum(axis=1, keepdims=True).astype(x.dtype)
        dist = -((x2 - (2 * xs) @ et) + e2)
        out[s:s + chunk] = np.argmax(dist, axis=-1)
    return out


def rvq_forward(x: np.ndarray, codebooks: np.ndarray, n_q: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """ResidualVectorQuantization.forward -- quantization/core_vq.py:385-415 (eval mode).

    x [B,D,T]; codebooks [n_layers,bins,D]. Returns (quantized [B,D,T], codes [n_q,B,T] int64,
    quantized_stack [n_q,B,D,T]).
    """
    bsz, d, t = x.shape
    residual = np.ascontiguousarray(np.transpose(x, (0, 2, 1))).reshape(bsz * t, d)  # core_vq.py:305
    qout = np.zeros_like(residual)
    codes = []
    stack = []
    for i in range(n_q):
        emb = codebooks[i].astype(x.dtype)
        ind = codebook_quantize(residual, emb)
        q = emb[ind]  # F.embedding, core_vq.py:200-202
        residual = residual - q  # core_vq.py:402
        qout = qout + q  # core_vq.py:404
        codes.append(ind.reshape(bsz, t))
        stack.append(np.transpose(q.reshape(bsz, t, d), (0, 2, 1)))
    quantized = np.ascontiguousarray(np.transpose(qout.reshape(bsz, t, d), (0, 2, 1)))
    return quantized, np.stack(codes), np.stack(stack)


def rvq_encode(x: np.ndarray, codebooks: np.ndarray, n_q: int) -> np.ndarray:
    """ResidualVectorQuantization.encode -- quantization/core_vq.py:417-432."""
    return rvq_forward(x, codebooks, n_q)[1]


def rvq_decode(codes: np.ndarray, codebooks: np.ndarray, n_q: Optional[int] = None, dtype=np.float32) -> np.ndarray:
    """ResidualVectorQuantization.decode -- quantization/core_vq.py:434-445. codes [n_q,B,T] -> [B,D,T]."""
    if n_q is None:
        n_q = codebooks.shape[0]
    out = None
    for i in range(min(n_q, codes.shape[0])):
        q = codebooks[i].astype(dtype)[codes[i]]  # [B,T,D]
        out = q if out is None else out + q
    return np.ascontiguousarray(np.transpose(out, (0, 2, 1)))


def codebooks_from_state_dict(sd: Dict[str, np.ndarray], n_q: int) -> np.ndarray:
    return np.stack([np.asarray(sd[f"quantizer.vq.layers.{i}._codebook.embed"]) for i in range(n_q)])


# ----------------------------------------------------------------------------------------------
# EncodecModel orchestration
# ----------------------------------------------------------------------------------------------
def linear_overlap_add(frames: List[np.ndarray], stride: int) -> np.ndarray:
    """utils._linear_overlap_add -- utils.py:17-56."""
    dtype = frames[0].dtype
    shape = frames[0].shape[:-1]
    total = stride * (len(frames) - 1) + frames[-1].shape[-1]
    flen = frames[0].shape[-1]
    t = np.linspace(0, 1, flen + 2, dtype=dtype)[1:-1]
    weight = (0.5 - np.abs(t - 0.5)).astype(dtype)
    sum_w = np.zeros(total, dtype=dtype)
    out = np.zeros(shape + (total,), dtype=dtype)
    off = 0
    for fr in frames:
        n = fr.shape[-1]
        out[..., off:off + n] += weight[:n] * fr
        sum_w[off:off + n] += weight[:n]
        off += stride
    assert sum_w.min() > 0
    return (out / sum_w).astype(dtype)


def encode_frame(x: np.ndarray, p: Params, spec, codebooks: np.ndarray, n_q: int, taps=None) -> dict:
    """EncodecModel._encode_frame -- model.py:175-210."""
    if spec.normalize:
        mono = x.mean(axis=1, keepdims=True, dtype=x.dtype)
        volume = np.sqrt((mono ** 2).mean(axis=2, keepdims=True, dtype=x.dtype))
        scale = (1e-8 + volume).astype(x.dtype)
        x = x / scale
        scale = scale.reshape(-1, 1)
    else:
        scale = None
    emb = seanet_encoder(x, p, spec, taps)
    quantized, codes, _ = rvq_forward(emb, codebooks, n_q)
    return {"emb": emb, "quantized": quantized, "codes": np.transpose(codes, (1, 0, 2)), "scale": scale}


def decode_frame(frame: dict, p: Params, spec, taps=None) -> np.ndarray:
    """EncodecModel._decode_frame -- model.py:229-246 (decodes frame['quantized'], fork delta D3)."""
    out = seanet_decoder(frame["quantized"], p, spec, taps)
    if frame["scale"] is not None:
        out = out * frame["scale"].reshape(-1, 1, 1)
    return out.astype(frame["quantized"].dtype)


def encode(x: np.ndarray, sd: Dict[str, np.ndarray], spec, bandwidth: Optional[float], dtype=np.float32) -> List[dict]:
    """EncodecModel.encode -- model.py:146-173."""
    assert x.ndim == 3 and 0 < x.shape[1] <= 2
    p = Params(sd, dtype, spec.norm)
    n_q = spec.n_q_for_bandwidth(bandwidth)
    codebooks = codebooks_from_state_dict(sd, n_q)
    x = x.astype(dtype)
    length = x.shape[-1]
    seg = spec.segment_length
    if seg is None:
        seg, stride = length, length
    else:
        stride = spec.segment_stride
    return [encode_frame(x[:, :, off:off + seg], p, spec, codebooks, n_q) for off in range(0, length, stride)]


def decode(frames: List[dict], sd: Dict[str, np.ndarray], spec, dtype=np.float32) -> np.ndarray:
    """EncodecModel.decode -- model.py:212-227."""
    p = Params(sd, dtype, spec.norm)
    if spec.segment_length is None:
        assert len(frames) == 1
        return decode_frame(frames[0], p, spec)
    outs = [decode_frame(f, p, spec) for f in frames]
    return linear_overlap_add(outs, spec.segment_stride or 1)


def forward(x: np.ndarray, sd: Dict[str, np.ndarray], spec, bandwidth: Optional[float], dtype=np.float32):
    """EncodecModel.forward -- model.py:248-257. Returns (audio[B,C,T], codes[B,n_q,sum T_f], frames)."""
    frames = encode(x, sd, spec, bandwidth, dtype)
    codes = np.concatenate([f["codes"] for f in frames], axis=-1)
    audio = decode(frames, sd, spec, dtype)[:, :, :x.shape[-1]]
    return audio, codes, frames


# ----------------------------------------------------------------------------------------------
# scoring helpers (last row)
# ----------------------------------------------------------------------------------------------
def classify_code_mismatches(residual_in: np.ndarray, embed: np.ndarray, ref_codes: np.ndarray,
                             got_codes: np.ndarray, rel_tol: float = 1e-5) -> Tuple[int, int]:
    """Teacher-forced near-tie check for ONE layer.

    residual_in [N,D] is the *reference* residual entering the layer. Returns
    (n_near_tie, n_hard) over rows where the codes differ; distances in float64.
    """
    bad = np.nonzero(ref_codes != got_codes)[0]
    near = hard = 0
    e = embed.astype(np.float64)
    for r in bad:
        xr = residual_in[r].astype(np.float64)
        d_ref = ((xr - e[ref_codes[r]]) ** 2).sum()
        d_got = ((xr - e[got_codes[r]]) ** 2).sum()
        if abs(d_ref - d_got) <= rel_tol * min(d_ref, d_got):
            near += 1
        else:
            hard += 1
    return near, hard


def score_codes(emb: np.ndarray, codebooks: np.ndarray, ref_codes: np.ndarray, got_codes: np.ndarray,
                rel_tol: float = 1e-5) -> dict:
    """Compare code tensors [n_q, N] layer by layer, following only frames whose prefix matched.

    ``emb`` [N,D] is the reference encoder output. Returns counts: compared, mismatched, near_tie, hard.
    """
    n_q, n = ref_codes.shape
    alive = np.ones(n, dtype=bool)
    residual = emb.astype(np.float64).copy()
    res = {"compared": 0, "mismatched": 0, "near_tie": 0, "hard": 0}
    for i in range(n_q):
        idx = np.nonzero(alive)[0]
        res["compared"] += idx.size
        diff = idx[ref_codes[i, idx] != got_codes[i, idx]]
        if diff.size:
            near, hard = classify_code_mismatches(residual[diff], codebooks[i], ref_codes[i, diff], got_codes[i, diff], rel_tol)
            res["mismatched"] += int(diff.size)
            res["near_tie"] += near
            res["hard"] += hard
            alive[diff] = False
        residual -= codebooks[i].astype(np.float64)[ref_codes[i]]
    return res
