"""CPU baseline port of the reference's hot path on the SAME library kernels the reference uses -- TEST / BENCH
INFRASTRUCTURE ONLY (never imported by the product package).

The reference (ellen660/encodec) is pure Python over PyTorch: all of its arithmetic is ATen on the CPU (mkldnn conv /
mkldnn RNN / MKL sgemm). The numpy oracle (oracle/encodec_oracle.py) restates the algorithm without
any framework, which makes it the right CHECKER but a slow stand-in for the reference's speed. This file restates the
same functions with the ATen calls the reference makes -- F.pad(reflect) + F.conv1d (modules/conv.py:80-97,116-128),
F.conv_transpose1d (:156-163), F.group_norm (:50), F.elu (modules/seanet.py:43), nn.LSTM (modules/lstm.py:20-26), the
``-(|x|^2 - 2 x@E^T + |E|^2)`` / ``max(-1).indices`` / F.embedding quantiser (quantization/core_vq.py:178-202,385-415)
and the segment / overlap-add orchestration (model.py:146-257, utils.py:17-56) -- so that ``bench.py``'s cpu_baseline and
``--impl reference`` legs time what the reference itself would cost on the box's host cores. ``tests/test_torch_port.py``
pins it to the numpy oracle (and therefore to the reference's golden outputs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class TorchParams:
    """Weight-norm folded once (the reference re-folds on every forward, conv.py:28-29: ~3 % of its CPU time)."""

    def __init__(self, sd: Dict[str, np.ndarray], norm: str = "time_group_norm"):
        self.sd = sd
        self.norm = norm
        self._conv = {}
        self._lstm = {}

    def conv(self, prefix: str, transposed: bool = False):
        key = (prefix, transposed)
        if key not in self._conv:
            base = f"{prefix}.convtr.convtr" if transposed else f"{prefix}.conv.conv"
            normp = f"{prefix}.convtr.norm" if transposed else f"{prefix}.conv.norm"
            sd = self.sd
            if base + ".weight_g" in sd:
                v = _t(sd[base + ".weight_v"])
                g = _t(sd[base + ".weight_g"])
                w = v * (g / v.norm(2, dim=tuple(range(1, v.dim())), keepdim=True))
            else:
                w = _t(sd[base + ".weight"])
            gamma = _t(sd[normp + ".weight"]) if normp + ".weight" in sd else None
            beta = _t(sd[normp + ".bias"]) if normp + ".bias" in sd else None
            self._conv[key] = (w, _t(sd[base + ".bias"]), gamma, beta)
        return self._conv[key]

    def lstm(self, prefix: str, num_layers: int) -> torch.nn.LSTM:
        if prefix not in self._lstm:
            hdim = self.sd[f"{prefix}.lstm.weight_hh_l0"].shape[1]
            m = torch.nn.LSTM(hdim, hdim, num_layers)
            with torch.no_grad():
                for layer in range(num_layers):
                    for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                        getattr(m, f"{n}_l{layer}").copy_(_t(self.sd[f"{prefix}.lstm.{n}_l{layer}"]))
            self._lstm[prefix] = m.eval()
        return self._lstm[prefix]


def pad1d_reflect(x: torch.Tensor, left: int, right: int) -> torch.Tensor:
    """modules/conv.py:80-97."""
    length = x.shape[-1]
    max_pad = max(left, right)
    extra = 0
    if length <= max_pad:
        extra = max_pad - length + 1
        x = F.pad(x, (0, extra))
    padded = F.pad(x, (left, right), mode="reflect")
    return padded[..., :padded.shape[-1] - extra]


def _norm(y, gamma, beta, kind: str):
    """get_norm_module -- modules/conv.py:38-52; ConvLayerNorm = LayerNorm over channels per time step (norm.py:16-30)."""
    if gamma is None:
        return y
    if kind == "layer_norm":
        return F.layer_norm(y.transpose(1, 2), (y.shape[1],), gamma, beta, 1e-5).transpose(1, 2)
    return F.group_norm(y, 1, gamma, beta, 1e-5)


def sconv1d(x, p: TorchParams, prefix: str, stride: int, causal: bool):
    """SConv1d.forward -- modules/conv.py:202-221."""
    w, b, gamma, beta = p.conv(prefix)
    k = w.shape[-1]
    padding_total = k - stride
    n_frames = (x.shape[-1] - k + padding_total) / stride + 1
    extra = (math.ceil(n_frames) - 1) * stride + (k - padding_total) - x.shape[-1]
    if causal:
        xp = pad1d_reflect(x, padding_total, extra)
    else:
        pr = padding_total // 2
        xp = pad1d_reflect(x, padding_total - pr, pr + extra)
    y = F.conv1d(xp, w, b, stride=stride)
    return _norm(y, gamma, beta, p.norm)


def sconvtr1d(x, p: TorchParams, prefix: str, stride: int, causal: bool):
    """SConvTranspose1d.forward -- modules/conv.py:241-263."""
    w, b, gamma, beta = p.conv(prefix, transposed=True)
    padding_total = w.shape[-1] - stride
    y = F.conv_transpose1d(x, w, b, stride=stride)
    y = _norm(y, gamma, beta, p.norm)
    pr = padding_total if causal else padding_total // 2
    pl = padding_total - pr
    return y[..., pl: y.shape[-1] - pr]


def slstm(x, p: TorchParams, prefix: str, num_layers: int):
    """SLSTM.forward -- modules/lstm.py:22-28."""
    inp = x.permute(2, 0, 1)
    y, _ = p.lstm(prefix, num_layers)(inp)
    return (y + inp).permute(1, 2, 0)


def resnet_block(x, p: TorchParams, prefix: str, causal: bool):
    """SEANetResnetBlock.forward -- modules/seanet.py:37-64."""
    h = sconv1d(F.elu(x), p, f"{prefix}.block.1", 1, causal)
    h = sconv1d(F.elu(h), p, f"{prefix}.block.3", 1, causal)
    return sconv1d(x, p, f"{prefix}.shortcut", 1, causal) + h


def seanet_encoder(x, p: TorchParams, spec):
    causal = spec.causal
    y = sconv1d(x, p, "encoder.model.0", 1, causal)
    idx = 1
    for ratio in reversed(spec.ratios):
        y = resnet_block(y, p, f"encoder.model.{idx}", causal)
        y = sconv1d(F.elu(y), p, f"encoder.model.{idx + 2}", ratio, causal)
        idx += 3
    if spec.lstm:
        y = slstm(y, p, f"encoder.model.{idx}", spec.lstm)
        idx += 1
    return sconv1d(F.elu(y), p, f"encoder.model.{idx + 1}", 1, causal)


def seanet_decoder(z, p: TorchParams, spec):
    causal = spec.causal
    y = sconv1d(z, p, "decoder.model.0", 1, causal)
    idx = 1
    if spec.lstm:
        y = slstm(y, p, f"decoder.model.{idx}", spec.lstm)
        idx += 1
    for ratio in spec.ratios:
        y = sconvtr1d(F.elu(y), p, f"decoder.model.{idx + 1}", ratio, causal)
        y = resnet_block(y, p, f"decoder.model.{idx + 2}", causal)
        idx += 3
    return sconv1d(F.elu(y), p, f"decoder.model.{idx + 1}", 1, causal)


def rvq_forward(emb, codebooks: List[torch.Tensor], n_q: int):
    """ResidualVectorQuantization.forward -- quantization/core_vq.py:385-415 with EuclideanCodebook.quantize :178-194."""
    b, d, t = emb.shape
    residual = emb
    out = torch.zeros_like(emb)
    codes = []
    for i in range(n_q):
        e = codebooks[i]
        x = residual.permute(0, 2, 1).reshape(-1, d)
        embed = e.t()
        dist = -(x.pow(2).sum(1, keepdim=True) - 2 * x @ embed + embed.pow(2).sum(0, keepdim=True))
        ind = dist.max(dim=-1).indices
        q = F.embedding(ind, e).view(b, t, d).permute(0, 2, 1)
        residual = residual - q
        out = out + q
        codes.append(ind.view(b, t))
    return out, torch.stack(codes)


def linear_overlap_add(frames: List[torch.Tensor], stride: int):
    """utils._linear_overlap_add -- utils.py:17-56."""
    shape = frames[0].shape[:-1]
    total = stride * (len(frames) - 1) + frames[-1].shape[-1]
    flen = frames[0].shape[-1]
    t = torch.linspace(0, 1, flen + 2)[1:-1]
    weight = 0.5 - (t - 0.5).abs()
    sum_w = torch.zeros(total)
    out = torch.zeros(*shape, total)
    off = 0
    for fr in frames:
        n = fr.shape[-1]
        out[..., off:off + n] += weight[:n] * fr
        sum_w[off:off + n] += weight[:n]
        off += stride
    return out / sum_w


@torch.no_grad()
def forward(x: np.ndarray, sd: Dict[str, np.ndarray], spec, bandwidth: Optional[float], params: Optional[TorchParams] = None):
    """EncodecModel.forward -- model.py:146-257. Returns (audio [B,C,T], codes [B,n_q,sum T_f]) as numpy arrays."""
    p = params or TorchParams(sd, spec.norm)
    n_q = spec.n_q_for_bandwidth(bandwidth)
    cbs = [_t(sd[f"quantizer.vq.layers.{i}._codebook.embed"]) for i in range(n_q)]
    xt = _t(x)
    length = xt.shape[-1]
    seg = spec.segment_length
    stride = spec.segment_stride if seg is not None else length
    seg = seg if seg is not None else length
    outs, codes = [], []
    for off in range(0, length, stride):
        fr = xt[:, :, off:off + seg]
        scale = None
        if spec.normalize:
            mono = fr.mean(dim=1, keepdim=True)
            scale = 1e-8 + mono.pow(2).mean(dim=2, keepdim=True).sqrt()
            fr = fr / scale
        emb = seanet_encoder(fr, p, spec)
        q, c = rvq_forward(emb, cbs, n_q)
        codes.append(c.transpose(0, 1))
        y = seanet_decoder(q, p, spec)
        if scale is not None:
            y = y * scale.view(-1, 1, 1)
        outs.append(y)
    audio = outs[0] if spec.segment_length is None else linear_overlap_add(outs, stride)
    return audio[:, :, :length].numpy(), torch.cat(codes, dim=-1).numpy()
