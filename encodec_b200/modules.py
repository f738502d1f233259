"""SEANet encoder / decoder drop-ins (reference ``encodec/modules/seanet.py``, ``conv.py``, ``lstm.py``).

The classes keep the reference's constructor signatures, attributes and ``state_dict`` key layout
 but hold parameters only: ``forward`` hands raw device pointers to the CUDA
library (``include/encodec_b200.h``), which folds weight-norm once per weight update and runs the whole
stack as hand-written sm_100a kernels. Hyper-parameter points the kernels do not cover raise
``NotImplementedError``; there is no PyTorch / CPU fallback.
"""
from __future__ import annotations

import math
import typing as tp

import numpy as np
import torch
from torch import nn

from . import _native as nat

_SUPPORTED_NORMS = ("weight_norm", "time_group_norm", "layer_norm")
_AFFINE_NORMS = ("time_group_norm", "layer_norm")
_NORM_CODE = {"weight_norm": 0, "time_group_norm": 1, "layer_norm": 2}


# ------------------------------------------------------------------------------------------------
# parameter holders (same attribute paths as the reference modules => same state_dict keys)
# ------------------------------------------------------------------------------------------------
class _ConvParams(nn.Module):
    """Parameters of nn.Conv1d / nn.ConvTranspose1d (+ old-style weight_norm), reference conv.py:26-35."""

    def __init__(self, c_in: int, c_out: int, k: int, transposed: bool, weight_norm: bool):
        super().__init__()
        shape = (c_in, c_out, k) if transposed else (c_out, c_in, k)
        w = torch.empty(shape)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))  # torch's default conv init
        bound = 1.0 / math.sqrt(shape[1] * k)
        b = torch.empty(c_out).uniform_(-bound, bound)
        if weight_norm:
            self.weight_g = nn.Parameter(w.flatten(1).norm(dim=1).view(-1, 1, 1).clone())
            self.weight_v = nn.Parameter(w)
        else:
            self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(b)


class _AffineParams(nn.Module):
    """Affine parameters of nn.GroupNorm(1, C) (reference conv.py:50) or ConvLayerNorm (conv.py:44-46, norm.py:16-30)."""

    def __init__(self, channels: int):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class NormConv1d(nn.Module):
    def __init__(self, c_in, c_out, k, norm: str):
        super().__init__()
        self.conv = _ConvParams(c_in, c_out, k, False, norm == "weight_norm")
        self.norm = _AffineParams(c_out) if norm in _AFFINE_NORMS else nn.Identity()
        self.norm_type = norm


class NormConvTranspose1d(nn.Module):
    def __init__(self, c_in, c_out, k, norm: str):
        super().__init__()
        self.convtr = _ConvParams(c_in, c_out, k, True, norm == "weight_norm")
        self.norm = _AffineParams(c_out) if norm in _AFFINE_NORMS else nn.Identity()
        self.norm_type = norm


class SConv1d(nn.Module):
    """Parameter holder for reference conv.py:182-221 (the arithmetic runs inside the fused stack)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, norm="none", causal=False):
        super().__init__()
        self.conv = NormConv1d(in_channels, out_channels, kernel_size, norm)
        self.causal = causal
        self.kernel_size, self.stride = kernel_size, stride


class SConvTranspose1d(nn.Module):
    """Parameter holder for reference conv.py:224-263."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, norm="none", causal=False):
        super().__init__()
        self.convtr = NormConvTranspose1d(in_channels, out_channels, kernel_size, norm)
        self.causal = causal
        self.kernel_size, self.stride = kernel_size, stride


class SEANetResnetBlock(nn.Module):
    """Parameter holder for reference seanet.py:22-64 (kernel_sizes [3, 1], conv shortcut)."""

    def __init__(self, dim: int, residual_kernel_size: int, norm: str, causal: bool, compress: int):
        super().__init__()
        hidden = dim // compress
        self.block = nn.ModuleDict({
            "1": SConv1d(dim, hidden, residual_kernel_size, norm=norm, causal=causal),
            "3": SConv1d(hidden, dim, 1, norm=norm, causal=causal),
        })
        self.shortcut = SConv1d(dim, dim, 1, norm=norm, causal=causal)


class SLSTM(nn.Module):
    """Parameter holder for reference lstm.py:12-28; ``self.lstm`` is never called, it only owns the
    ``weight_ih_l*`` / ``weight_hh_l*`` / ``bias_*`` tensors with torch's default initialisation."""

    def __init__(self, dimension: int, num_layers: int = 2, skip: bool = True):
        super().__init__()
        self.skip = skip
        self.lstm = nn.LSTM(dimension, dimension, num_layers)


class _Placeholder(nn.Module):
    """Stands where the reference Sequential has a parameter-free nn.ELU, to keep module indices."""


def _check_common(activation, activation_params, norm, norm_params, n_residual_layers, pad_mode, true_skip,
                  kernel_size, last_kernel_size, residual_kernel_size, compress, n_filters, dimension, causal):
    if activation != "ELU" or float(activation_params.get("alpha", 1.0)) != 1.0:
        raise NotImplementedError("encodec_b200: only activation='ELU' with alpha=1.0 is implemented")
    if norm not in _SUPPORTED_NORMS:
        raise NotImplementedError(f"encodec_b200: norm={norm!r} is not implemented (supported: {_SUPPORTED_NORMS})")
    if norm == "time_group_norm" and causal:
        raise ValueError("GroupNorm doesn't support causal evaluation.")  # reference conv.py:47-48
    if norm_params:
        raise NotImplementedError("encodec_b200: norm_params are not supported")
    if n_residual_layers != 1 or true_skip or pad_mode != "reflect":
        raise NotImplementedError("encodec_b200: needs n_residual_layers=1, true_skip=False, pad_mode='reflect'")
    if (kernel_size, last_kernel_size, residual_kernel_size, compress) != (7, 7, 3, 2):
        raise NotImplementedError("encodec_b200: needs kernel_size=7, last_kernel_size=7, residual_kernel_size=3, compress=2")
    if n_filters != 32 or dimension not in (128, 256):
        raise NotImplementedError("encodec_b200: needs n_filters=32 and dimension 128 or 256")


class _NativeStack(nn.Module):
    """Shared machinery: lazily (re)build the native codec handle when parameters change."""

    _prefix = ""

    def _init_native(self):
        self.__dict__["_codec"] = None
        self.__dict__["_codec_sig"] = None

    def _signature(self):
        """(data_ptr, version, device) of every parameter / buffer. The tensors are read from each sub-module's own dicts on
        every call (a replaced or re-assigned Parameter is seen); only the LIST of sub-modules is cached -- walking
        ``self.parameters()`` was ~0.3 ms per stack and call, on the critical path of a 1 ms forward -- and rebuilt on
        ``_apply`` (.to / .cuda / .float), ``load_state_dict`` and every 256th call."""
        mods = self.__dict__.get("_sig_modules")
        calls = self.__dict__.get("_sig_calls", 0)
        if mods is None or (calls & 255) == 0:
            mods = list(self.modules())
            self.__dict__["_sig_modules"] = mods
        self.__dict__["_sig_calls"] = calls + 1
        sig = []
        for md in mods:
            for t in md._parameters.values():
                if t is not None:
                    sig.append((t.data_ptr(), t._version, t.device))
            for t in md._buffers.values():
                if t is not None:
                    sig.append((t.data_ptr(), t._version, t.device))
        return tuple(sig)

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_sig_modules"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__["_sig_modules"] = None
        return super().load_state_dict(*args, **kwargs)

    def _spec(self) -> nat.EcbSpec:
        return nat.make_spec(self.channels, self.causal, _NORM_CODE[self.norm], self.n_filters, self.dimension,
                             self._dec_ratios, 7, 7, 3, 2, self.lstm_layers, 128, 1)

    def native(self) -> nat.Codec:
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("encodec_b200: module parameters must live on a CUDA device (no CPU fallback); "
                               "call .cuda() first")
        sig = self._signature()
        if self._codec is None or sig != self._codec_sig:
            codec = nat.Codec(self._spec(), p.device)
            sd = {self._prefix + k: v for k, v in self.state_dict().items()}
            codec.load(sd)
            self.__dict__["_codec"] = codec
            self.__dict__["_codec_sig"] = sig
        return self._codec


class SEANetEncoder(_NativeStack):
    """SEANet encoder -- same constructor and ``forward(x[B,C,T]) -> [B,D,ceil(T/hop)]`` as the reference
    (modules/seanet.py:67-146)."""

    _prefix = "encoder."

    def __init__(self, channels: int = 1, dimension: int = 128, n_filters: int = 32, n_residual_layers: int = 1,
                 ratios: tp.List[int] = [8, 5, 4, 2], activation: str = 'ELU', activation_params: dict = {'alpha': 1.0},
                 norm: str = 'weight_norm', norm_params: tp.Dict[str, tp.Any] = {}, kernel_size: int = 7,
                 last_kernel_size: int = 7, residual_kernel_size: int = 3, dilation_base: int = 2, causal: bool = False,
                 pad_mode: str = 'reflect', true_skip: bool = False, compress: int = 2, lstm: int = 2):
        super().__init__()
        _check_common(activation, activation_params, norm, norm_params, n_residual_layers, pad_mode, true_skip,
                      kernel_size, last_kernel_size, residual_kernel_size, compress, n_filters, dimension, causal)
        self.channels = channels
        self.dimension = dimension
        self.n_filters = n_filters
        self._dec_ratios = list(ratios)
        self.ratios = list(reversed(ratios))  # reference seanet.py:102
        self.n_residual_layers = n_residual_layers
        self.hop_length = int(np.prod(self.ratios))
        self.causal = causal
        self.norm = norm
        self.lstm_layers = lstm

        model: tp.Dict[str, nn.Module] = {}
        mult, idx = 1, 1
        model["0"] = SConv1d(channels, mult * n_filters, kernel_size, norm=norm, causal=causal)
        for ratio in self.ratios:
            model[str(idx)] = SEANetResnetBlock(mult * n_filters, residual_kernel_size, norm, causal, compress)
            model[str(idx + 1)] = _Placeholder()
            model[str(idx + 2)] = SConv1d(mult * n_filters, mult * n_filters * 2, ratio * 2, stride=ratio, norm=norm,
                                          causal=causal)
            idx += 3
            mult *= 2
        if lstm:
            model[str(idx)] = SLSTM(mult * n_filters, num_layers=lstm)
            idx += 1
        model[str(idx)] = _Placeholder()
        model[str(idx + 1)] = SConv1d(mult * n_filters, dimension, last_kernel_size, norm=norm, causal=causal)
        self.model = nn.ModuleDict(model)
        self._init_native()

    @torch.no_grad()
    def encode_items(self, x: torch.Tensor, n_items: int, n_seg: int, length: int, batch_stride: int, seg_stride: int,
                     chan_stride: int, want_scale: bool, want_channels_first: bool = True):
        """Run the stack on ``n_items`` windows of ``x`` read in place (see ecb_encoder_forward).

        Returns (emb [n_items, D, T_f] or None, emb_frames [n_items*T_f, D], scale [n_items] or None).
        """
        codec = self.native()
        dev = x.device
        t_f = -(-length // self.hop_length)
        emb = torch.empty((n_items, self.dimension, t_f), dtype=torch.float32, device=dev) if want_channels_first else None
        frames = torch.empty((n_items * t_f, self.dimension), dtype=torch.float32, device=dev)
        scale = torch.empty((n_items,), dtype=torch.float32, device=dev) if want_scale else None
        with torch.cuda.device(dev):
            nbytes = nat.lib.ecb_encoder_workspace_bytes(codec.handle, n_items, length)
            ws = nat.shared_workspace(dev, nbytes)
            nat.check(nat.lib.ecb_encoder_forward(codec.handle, nat.ptr(x), n_items, n_seg, length, batch_stride,
                                                  seg_stride, chan_stride, nat.ptr(scale), nat.ptr(emb), nat.ptr(frames),
                                                  nat.ptr(ws), ws.numel(), nat.stream_ptr(dev)))
        return emb, frames, scale

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        nat.require_cuda(x, "SEANetEncoder input")
        assert x.dim() == 3 and x.shape[1] == self.channels, (x.shape, self.channels)
        x = x.contiguous()
        b, c, t = x.shape
        emb, _, _ = self.encode_items(x, b, 1, t, c * t, 0, t, False)
        return emb


class SEANetDecoder(_NativeStack):
    """SEANet decoder -- same constructor and ``forward(z[B,D,T_f]) -> [B,C,T_f*hop]`` as the reference
    (modules/seanet.py:149-253; the last conv has norm='none' as in the fork, :227-228)."""

    _prefix = "decoder."

    def __init__(self, channels: int = 1, dimension: int = 128, n_filters: int = 32, n_residual_layers: int = 1,
                 ratios: tp.List[int] = [8, 5, 4, 2], activation: str = 'ELU', activation_params: dict = {'alpha': 1.0},
                 final_activation: tp.Optional[str] = None, final_activation_params: tp.Optional[dict] = None,
                 norm: str = 'weight_norm', norm_params: tp.Dict[str, tp.Any] = {}, kernel_size: int = 7,
                 last_kernel_size: int = 7, residual_kernel_size: int = 3, dilation_base: int = 2, causal: bool = False,
                 pad_mode: str = 'reflect', true_skip: bool = False, compress: int = 2, lstm: int = 2,
                 trim_right_ratio: float = 1.0):
        super().__init__()
        _check_common(activation, activation_params, norm, norm_params, n_residual_layers, pad_mode, true_skip,
                      kernel_size, last_kernel_size, residual_kernel_size, compress, n_filters, dimension, causal)
        if final_activation is not None:
            raise NotImplementedError("encodec_b200: final_activation is not implemented")
        if trim_right_ratio != 1.0:
            raise NotImplementedError("encodec_b200: trim_right_ratio != 1.0 is not implemented")
        if not lstm:
            raise NotImplementedError("encodec_b200: a decoder without LSTM is not implemented")
        self.dimension = dimension
        self.channels = channels
        self.n_filters = n_filters
        self.ratios = list(ratios)
        self._dec_ratios = list(ratios)
        self.n_residual_layers = n_residual_layers
        self.hop_length = int(np.prod(self.ratios))
        self.causal = causal
        self.norm = norm
        self.lstm_layers = lstm

        model: tp.Dict[str, nn.Module] = {}
        mult = int(2 ** len(self.ratios))
        model["0"] = SConv1d(dimension, mult * n_filters, kernel_size, norm=norm, causal=causal)
        idx = 1
        if lstm:
            model[str(idx)] = SLSTM(mult * n_filters, num_layers=lstm)
            idx += 1
        for ratio in self.ratios:
            model[str(idx)] = _Placeholder()
            model[str(idx + 1)] = SConvTranspose1d(mult * n_filters, mult * n_filters // 2, ratio * 2, stride=ratio,
                                                   norm=norm, causal=causal)
            model[str(idx + 2)] = SEANetResnetBlock(mult * n_filters // 2, residual_kernel_size, norm, causal, compress)
            idx += 3
            mult //= 2
        model[str(idx)] = _Placeholder()
        model[str(idx + 1)] = SConv1d(n_filters, channels, last_kernel_size, norm="none", causal=causal)  # fork delta D7
        self.model = nn.ModuleDict(model)
        self._init_native()

    #: Operand scheme of the decoder's tensor-core convs. ``None`` (default) = automatic: weight-norm models run ONE TF32
    #: pass -- nothing downstream of the decoder is discrete, the decoded audio stays within ~1e-4 max-abs / 2.5e-5 RMS of
    #: the reference (bar 1e-3 / 1e-4; tests on the golden cases and on real speech at three loudness levels), decoder convs
    #: ~1.6x faster -- while GroupNorm (48 kHz) and LayerNorm models keep split operands (their O(1) output puts TF32's
    #: ~3e-4 relative error outside the absolute RMS bar). ``False``: split operands (3xTF32, fp32-accurate) everywhere, as
    #: the encoder and the quantiser always use. ``True``: TF32 also in the >= 128-channel convs of a GroupNorm decoder.
    tf32: tp.Optional[bool] = None

    @torch.no_grad()
    def decode_items(self, z: tp.Optional[torch.Tensor], z_frames: tp.Optional[torch.Tensor], n_items: int,
                     n_frames: int, scale: tp.Optional[torch.Tensor], out: tp.Optional[torch.Tensor] = None):
        codec = self.native()
        nat.check(nat.lib.ecb_codec_set_decoder_precision(codec.handle, -1 if self.tf32 is None else int(bool(self.tf32))))
        src = z if z is not None else z_frames
        dev = src.device
        if out is None:
            out = torch.empty((n_items, self.channels, n_frames * self.hop_length), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            nbytes = nat.lib.ecb_decoder_workspace_bytes(codec.handle, n_items, n_frames)
            ws = nat.shared_workspace(dev, nbytes)
            nat.check(nat.lib.ecb_decoder_forward(codec.handle, nat.ptr(z), nat.ptr(z_frames), n_items, n_frames,
                                                  nat.ptr(scale), nat.ptr(out), nat.ptr(ws), ws.numel(),
                                                  nat.stream_ptr(dev)))
        return out

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        nat.require_cuda(z, "SEANetDecoder input")
        assert z.dim() == 3 and z.shape[1] == self.dimension, (z.shape, self.dimension)
        z = z.contiguous()
        return self.decode_items(z, None, z.shape[0], z.shape[2], None)
