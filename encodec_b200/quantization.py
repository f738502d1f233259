"""Residual vector quantiser drop-in (reference ``encodec/quantization/vq.py`` + ``core_vq.py``).

Inference only: the EMA / k-means / dead-code machinery of ``core_vq.py:143-176,240-252`` is training
state and is not reproduced; a codebook whose ``inited`` flag is 0 raises instead of silently running
k-means on the first batch ("first-call k-means trap").
"""
from __future__ import annotations

import math
import typing as tp
from dataclasses import dataclass, field

import torch
from torch import nn

from . import _native as nat


@dataclass
class QuantizedResult:
    """Same fields as the reference dataclass (quantization/vq.py:21-30)."""
    quantized: torch.Tensor
    codes: torch.Tensor
    bandwidth: torch.Tensor  # bandwidth in kb/s used, per batch item.
    soft_targets: tp.Optional[torch.Tensor] = None
    commit_loss: tp.Optional[torch.Tensor] = None
    codebook_loss: tp.Optional[torch.Tensor] = None
    latents: tp.Optional[torch.Tensor] = None
    metrics: dict = field(default_factory=dict)


class EuclideanCodebook(nn.Module):
    """Buffer holder with the reference's names (core_vq.py:128-135)."""

    def __init__(self, dim: int, codebook_size: int, kmeans_init: bool):
        super().__init__()
        self.codebook_size = codebook_size
        embed = torch.zeros(codebook_size, dim)
        if not kmeans_init:
            nn.init.kaiming_uniform_(embed)  # uniform_init, core_vq.py:58-61
        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())


class VectorQuantization(nn.Module):
    def __init__(self, dim: int, codebook_size: int, kmeans_init: bool):
        super().__init__()
        self._codebook = EuclideanCodebook(dim, codebook_size, kmeans_init)
        self.codebook_size = codebook_size

    @property
    def codebook(self):
        return self._codebook.embed


class ResidualVectorQuantization(nn.Module):
    """Layer container (core_vq.py:364-374). ``share_codebook=True`` reproduces the fork: every entry of
    ``layers`` is the SAME VectorQuantization object (SURVEY.md delta D6)."""

    def __init__(self, num_quantizers: int, dim: int, codebook_size: int, kmeans_init: bool, share_codebook: bool):
        super().__init__()
        if share_codebook:
            one = VectorQuantization(dim, codebook_size, kmeans_init)
            self.layers = nn.ModuleList([one for _ in range(num_quantizers)])
        else:
            self.layers = nn.ModuleList([VectorQuantization(dim, codebook_size, kmeans_init)
                                         for _ in range(num_quantizers)])


class ResidualVectorQuantizer(nn.Module):
    """Residual Vector Quantizer -- reference signature (quantization/vq.py:46-56) plus ``share_codebook``.

    ``forward / encode / decode / intermediate_results`` keep the reference's shapes: RVQ-level codes are
    ``[n_q, B, T]`` int64, tensors are ``[B, D, T]``.
    """

    def __init__(self, dimension: int = 256, n_q: int = 8, bins: int = 1024, codebook_dim: int = 8,
                 decay: float = 0.99, kmeans_init: bool = True, kmeans_iters: int = 50,
                 threshold_ema_dead_code: int = 2, share_codebook: bool = True):
        super().__init__()
        if codebook_dim != dimension:
            raise NotImplementedError("encodec_b200: project_in/project_out (codebook_dim != dimension) is not implemented")
        if dimension not in (128, 256):
            raise NotImplementedError("encodec_b200: only dimension 128 or 256 is implemented")
        if bins % 128 != 0:
            raise NotImplementedError("encodec_b200: bins must be a multiple of 128")
        self.n_q = n_q
        self.dimension = dimension
        self.bins = bins
        self.decay = decay
        self.kmeans_init = kmeans_init
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.vq = ResidualVectorQuantization(n_q, dimension, bins, kmeans_init, share_codebook)
        self.__dict__["_codec"] = None
        self.__dict__["_codec_sig"] = None

    # ---- native handle -------------------------------------------------------------------------
    def native(self) -> nat.Codec:
        embeds = [layer._codebook.embed for layer in self.vq.layers]
        if not embeds[0].is_cuda:
            raise RuntimeError("encodec_b200: codebooks must live on a CUDA device (no CPU fallback)")
        sig = tuple((e.data_ptr(), e._version, e.device) for e in embeds) + \
            tuple((l._codebook.inited.data_ptr(), l._codebook.inited._version) for l in self.vq.layers)
        if self._codec is None or sig != self._codec_sig:
            for i, layer in enumerate(self.vq.layers):
                if float(layer._codebook.inited.item()) == 0.0:
                    raise RuntimeError(
                        f"encodec_b200: codebook {i} is not initialised (inited == 0). The reference would run k-means "
                        "on the first batch (core_vq.py:143-153); load a state_dict or set embed and inited first.")
            spec = nat.make_spec(1, True, 0, 32, self.dimension, [8, 5, 4, 2], 7, 7, 3, 2, 2, self.bins, self.n_q)
            codec = nat.Codec(spec, embeds[0].device)
            codec.load({f"quantizer.vq.layers.{i}._codebook.embed": e for i, e in enumerate(embeds)})
            self.__dict__["_codec"] = codec
            self.__dict__["_codec_sig"] = sig
        return self._codec

    @property
    def codebooks(self):
        return {i: layer.codebook for i, layer in enumerate(self.vq.layers)}

    # ---- reference API -------------------------------------------------------------------------
    def get_bandwidth_per_quantizer(self, frame_rate: int):
        """quantization/vq.py:127-131."""
        return math.log2(self.bins) * frame_rate

    def get_num_quantizers_for_bandwidth(self, frame_rate: int, bandwidth: tp.Optional[float] = None) -> int:
        """quantization/vq.py:116-125."""
        bw_per_q = self.get_bandwidth_per_quantizer(frame_rate)
        n_q = self.n_q
        if bandwidth and bandwidth > 0.:
            n_q = int(max(1, math.floor(bandwidth * 1000 / bw_per_q)))
        return n_q

    @torch.no_grad()
    def quantize_frames(self, x: tp.Optional[torch.Tensor], x_frames: tp.Optional[torch.Tensor], batch: int,
                        n_frames: int, n_q: int, want_channels_first: bool = True, want_stack: bool = False):
        """Fused RVQ on ``x`` [B,D,T] or frames-major ``x_frames`` [B*T,D].

        Returns (codes [n_q,B,T] int64, quantized [B,D,T] or None, quantized_frames [B*T,D], stack or None).
        """
        codec = self.native()
        if not (1 <= n_q <= self.n_q):
            raise ValueError(f"n_q={n_q} out of range 1..{self.n_q}")
        src = x if x is not None else x_frames
        dev = src.device
        codes = torch.empty((n_q, batch, n_frames), dtype=torch.int64, device=dev)
        quantized = torch.empty((batch, self.dimension, n_frames), dtype=torch.float32, device=dev) \
            if want_channels_first else None
        qf = torch.empty((batch * n_frames, self.dimension), dtype=torch.float32, device=dev)
        stack = torch.empty((n_q, batch, self.dimension, n_frames), dtype=torch.float32, device=dev) if want_stack else None
        with torch.cuda.device(dev):
            nbytes = nat.lib.ecb_codec_rvq_workspace_bytes(codec.handle, batch, n_frames)
            ws = nat.shared_workspace(dev, nbytes)
            nat.check(nat.lib.ecb_codec_rvq_forward(codec.handle, nat.ptr(x), nat.ptr(x_frames), batch, n_frames, n_q,
                                                    nat.ptr(codes), nat.ptr(quantized), nat.ptr(qf), nat.ptr(stack),
                                                    nat.ptr(ws), ws.numel(), nat.stream_ptr(dev)))
        return codes, quantized, qf, stack

    def _losses(self, n_q: int, device) -> torch.Tensor:
        # eval mode: every layer's loss is torch.tensor([0.0]) (core_vq.py:333) stacked to [n_q, 1]
        return torch.zeros((n_q, 1), dtype=torch.float32, device=device)

    def forward(self, x: torch.Tensor, frame_rate: int, bandwidth: tp.Optional[float] = None) -> QuantizedResult:
        """quantization/vq.py:91-114."""
        nat.require_cuda(x, "ResidualVectorQuantizer input")
        bw_per_q = self.get_bandwidth_per_quantizer(frame_rate)
        n_q = self.get_num_quantizers_for_bandwidth(frame_rate, bandwidth)
        x = x.contiguous()
        codes, quantized, _, _ = self.quantize_frames(x, None, x.shape[0], x.shape[2], n_q)
        commit_loss = self._losses(n_q, x.device)
        bw = torch.tensor(n_q * bw_per_q).to(x)
        return QuantizedResult(quantized, codes, bw, None, commit_loss, commit_loss)

    def intermediate_results(self, x: torch.Tensor, n_q: int):
        """quantization/vq.py:80-89."""
        nat.require_cuda(x, "ResidualVectorQuantizer input")
        x = x.contiguous()
        n_q = n_q or self.n_q
        codes, quantized, _, stack = self.quantize_frames(x, None, x.shape[0], x.shape[2], n_q, want_stack=True)
        return {"quantized": quantized, "codes": codes, "commit_loss": self._losses(n_q, x.device),
                "quantized_stack": stack}

    def encode(self, x: torch.Tensor, frame_rate: int, bandwidth: tp.Optional[float] = None) -> torch.Tensor:
        """quantization/vq.py:133-140 -> codes [n_q, B, T]."""
        nat.require_cuda(x, "ResidualVectorQuantizer input")
        n_q = self.get_num_quantizers_for_bandwidth(frame_rate, bandwidth)
        x = x.contiguous()
        return self.quantize_frames(x, None, x.shape[0], x.shape[2], n_q, want_channels_first=False)[0]

    @torch.no_grad()
    def decode_frames(self, codes: torch.Tensor, n_q: tp.Optional[int] = None, want_channels_first: bool = True):
        codec = self.native()
        nat.require_cuda(codes, "codes", torch.int64)
        codes = codes.contiguous()
        k, batch, n_frames = codes.shape
        n_q = min(k, self.n_q if n_q is None else n_q)
        dev = codes.device
        qf = torch.empty((batch * n_frames, self.dimension), dtype=torch.float32, device=dev)
        quantized = torch.empty((batch, self.dimension, n_frames), dtype=torch.float32, device=dev) \
            if want_channels_first else None
        with torch.cuda.device(dev):
            nat.check(nat.lib.ecb_codec_rvq_decode(codec.handle, nat.ptr(codes), batch, n_frames, n_q,
                                                   nat.ptr(quantized), nat.ptr(qf), nat.stream_ptr(dev)))
        return quantized, qf

    def decode(self, codes: torch.Tensor, n_q=None) -> torch.Tensor:
        """quantization/vq.py:142-147: codes [n_q, B, T] -> quantized [B, D, T]."""
        return self.decode_frames(codes, n_q)[0]
