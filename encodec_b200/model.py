"""EncodecModel drop-in (reference ``encodec/model.py:86-382``).

The fork's API is reproduced exactly (deltas D1-D4): ``encode`` returns a list of
dicts, ``decode`` consumes ``frame['quantized']``, ``forward`` returns a 4-tuple. As a superset
``decode`` also accepts the upstream ``(codes, scale)`` tuples and dicts without ``'quantized'``
(SURVEY.md section 8f row 1): those are de-quantised from the codes first.

Differences in mechanism, not in results: all segments of the 48 kHz model (reference Python loop
``model.py:168-170``) and all batch items go through ONE batched launch sequence per segment length; the
per-segment loudness scale, reflect padding, ELU, residual adds, weight-norm and GroupNorm are fused into
the CUDA kernels; overlap-add is one gather kernel.
"""
from __future__ import annotations

import math
import typing as tp
import warnings

import numpy as np
import torch
from torch import nn

from . import _native as nat
from .modules import SEANetDecoder, SEANetEncoder
from .quantization import QuantizedResult, ResidualVectorQuantizer  # noqa: F401

EncodedFrame = tp.Dict[str, tp.Optional[torch.Tensor]]


def _version_of(t):
    """Version counter of a tensor (None for None; -1 where autograd keeps none, e.g. inference tensors)."""
    if t is None:
        return None
    try:
        return t._version
    except RuntimeError:
        return -1


def _untouched(frame, tag) -> bool:
    _, _, q, qv, sc, scv = tag
    if qv == -1 or scv == -1:
        return False
    return frame.get('quantized') is q and _version_of(q) == qv and frame.get('scale') is sc and _version_of(sc) == scv


class _Batched:
    """Private side-channel from encode() to decode(): the batched, frames-major tensors of one call."""

    def __init__(self):
        self.groups = []  # list of dict(seg_ids, length, n_frames, quantized_frames, scale, batch)


class _Frame(dict):
    """A plain dict with the reference's six keys; the batched tensors ride along as an attribute."""
    __slots__ = ("_batched",)


ROOT_URL = 'https://dl.fbaipublicfiles.com/encodec/v0/'   # reference model.py:30


class EncodecModel(nn.Module):
    """EnCodec model operating on the raw waveform (same constructor as reference model.py:99-110). Inference only."""
    _warned_training = False

    def __init__(self,
                 encoder: SEANetEncoder,
                 decoder: SEANetDecoder,
                 quantizer: ResidualVectorQuantizer,
                 target_bandwidths: tp.List[float],
                 sample_rate: int,
                 channels: int,
                 normalize: bool = False,
                 segment: tp.Optional[float] = None,
                 overlap: float = 0.01,
                 name: str = 'unset'):
        super().__init__()
        self.bandwidth: tp.Optional[float] = None
        self.target_bandwidths = target_bandwidths
        self.encoder = encoder
        self.quantizer = quantizer
        self.decoder = decoder
        self.sample_rate = sample_rate
        self.channels = channels
        self.normalize = normalize
        self.segment = segment
        self.overlap = overlap
        self.frame_rate = math.ceil(self.sample_rate / np.prod(self.encoder.ratios))
        self.name = name
        self.bits_per_codebook = int(math.log2(self.quantizer.bins))
        self.n_q = quantizer.n_q
        assert 2 ** self.bits_per_codebook == self.quantizer.bins, \
            "quantizer bins must be a power of 2."
        # largest number of (batch item, segment) windows pushed through the stacks in one launch sequence;
        # bounds the activation workspace (n_items * length * 32 floats * 3-4 buffers). 72 GB of a B200's 180 GB: large
        # launches matter, the LSTM recurrence switches to its tensor-core form from 320 items per launch
        self.max_items_bytes = 72 << 30

    # ---- reference properties ------------------------------------------------------------------
    @property
    def segment_length(self) -> tp.Optional[int]:
        if self.segment is None:
            return None
        return int(self.segment * self.sample_rate)

    @property
    def segment_stride(self) -> tp.Optional[int]:
        segment_length = self.segment_length
        if segment_length is None:
            return None
        return max(1, int((1 - self.overlap) * segment_length))

    @property
    def codebooks(self):
        return self.quantizer.codebooks

    # ---- helpers -------------------------------------------------------------------------------
    def _segments(self, length: int):
        """(offset, length) of every segment, reference model.py:157-170."""
        seg = self.segment_length
        if seg is None:
            return [(0, length)], length
        stride = self.segment_stride
        assert stride is not None
        return [(off, min(seg, length - off)) for off in range(0, length, stride)], stride

    def _batch_chunk(self, n_seg: int, length: int, batch: int) -> int:
        """Clips per launch sequence: bounded by the activation workspace (``max_items_bytes``) and by the native limits on
        items per launch -- 65 535 (grid dimension; codec.cu setup_ctx) and, for inputs shorter than 32 latent frames (the
        CUDA-core path with its shared-memory resident LSTM state), 256."""
        n_seg = max(1, n_seg)
        per_item = (length + 2 * self.encoder.hop_length) * 32 * 4 * 4 * n_seg
        cap = 65535 // n_seg
        if -(-length // self.encoder.hop_length) < 32:
            cap = min(cap, max(1, 256 // n_seg))
        return max(1, min(batch, cap, int(self.max_items_bytes // per_item)))

    def _n_q(self) -> int:
        return self.quantizer.get_num_quantizers_for_bandwidth(self.frame_rate, self.bandwidth)

    # ---- encode --------------------------------------------------------------------------------
    @torch.no_grad()
    def _encode_batched(self, x: torch.Tensor, seg_range: tp.Optional[tp.Tuple[int, int]] = None):
        """All segments of all batch items (or segments ``seg_range = (s0, s1)`` of every item: the segments of the 48 kHz
        model are independent, reference model.py:157-170, which is what multi-GPU partitioning by segment relies on).
        Returns per segment-length group the batched results."""
        nat.require_cuda(x, "EncodecModel input")
        assert x.dim() == 3
        batch, channels, length = x.shape
        assert channels > 0 and channels <= 2
        x = x.contiguous()
        segs, stride = self._segments(length)
        if seg_range is not None:
            assert self.segment is not None and 0 <= seg_range[0] < seg_range[1] <= len(segs), (seg_range, len(segs))
            segs = segs[seg_range[0]:seg_range[1]]
        n_q = self._n_q()
        hop = self.encoder.hop_length
        dim = self.encoder.dimension
        # group consecutive segments of equal length: [full ... full] [short]? [shorter]?
        groups = []
        i = 0
        while i < len(segs):
            j = i
            while j + 1 < len(segs) and segs[j + 1][1] == segs[i][1]:
                j += 1
            groups.append((i, j - i + 1, segs[i][1]))
            i = j + 1
        out_groups = []
        for first, n_seg, seg_len in groups:
            assert self.segment is None or seg_len / self.sample_rate <= 1e-5 + self.segment  # model.py:178
            t_f = -(-seg_len // hop)
            n_items = batch * n_seg
            bc = self._batch_chunk(n_seg, seg_len, batch)
            parts = []
            for b0 in range(0, batch, bc):
                b1 = min(batch, b0 + bc)
                ni = (b1 - b0) * n_seg
                xv = x[b0:, :, segs[first][0]:]
                _, emb_frames, sc = self.encoder.encode_items(
                    xv, ni, n_seg, seg_len, channels * length, stride if n_seg > 1 else 0, length,
                    self.normalize, want_channels_first=False)
                c, q, f, _ = self.quantizer.quantize_frames(None, emb_frames, ni, t_f, n_q)
                parts.append((c, q, f, sc))
            if len(parts) == 1:
                codes, quant, qf, scale = parts[0]
            else:
                codes = torch.cat([p[0] for p in parts], dim=1)
                quant = torch.cat([p[1] for p in parts], dim=0)
                qf = torch.cat([p[2] for p in parts], dim=0)
                scale = torch.cat([p[3] for p in parts], dim=0) if self.normalize else None
            out_groups.append(dict(first=first, n_seg=n_seg, length=seg_len, n_frames=t_f, codes=codes,
                                   quantized=quant, quantized_frames=qf, scale=scale, batch=batch))
        return out_groups, n_q

    def encode_segments(self, x: torch.Tensor, s0: int, s1: int) -> tp.List[EncodedFrame]:
        """``encode`` restricted to segments ``s0 .. s1 - 1`` of every clip (segmented models). ``decode_segments`` turns the
        frames into per-segment audio; the overlap-add happens wherever all segments of a clip meet
        (``encodec_b200.dist.forward_sharded_segments``)."""
        return self.encode(x, _seg_range=(s0, s1))

    def encode(self, x: torch.Tensor, _seg_range: tp.Optional[tp.Tuple[int, int]] = None) -> tp.List[EncodedFrame]:
        """Same contract as reference model.py:146-210: one dict per segment with keys
        ``quantized [B,D,T_f]``, ``codes [B,K,T_f]``, ``soft_targets``, ``commit_loss [K,1]``,
        ``codebook_loss`` (the same tensor object) and ``scale [B,1]`` (or None).

        Inference only: eval semantics whatever ``self.training`` says (commit loss zero, no EMA / k-means / code expiry, no
        autograd graph) -- a training loop pointed at this class would run and learn nothing, hence the warning."""
        if self.training and torch.is_grad_enabled() and not EncodecModel._warned_training:
            EncodecModel._warned_training = True
            warnings.warn("encodec_b200.EncodecModel is inference-only: it always runs the reference's eval() forward (zero commit "
                          "loss, frozen codebooks, no gradients), also in train mode with autograd enabled", RuntimeWarning, stacklevel=2)
        groups, n_q = self._encode_batched(x, _seg_range)
        batched = _Batched()
        batched.groups = groups
        frames: tp.List[EncodedFrame] = []
        for g in groups:
            b, n_seg, t_f = g["batch"], g["n_seg"], g["n_frames"]
            codes = g["codes"].view(n_q, b, n_seg, t_f)
            quant = g["quantized"].view(b, n_seg, -1, t_f)
            scale = g["scale"].view(b, n_seg) if g["scale"] is not None else None
            for s in range(n_seg):
                loss = torch.zeros((n_q, 1), dtype=torch.float32, device=x.device)
                fr = _Frame({
                    'quantized': quant[:, s],
                    'codes': codes[:, :, s].transpose(0, 1),   # [B, K, T], as model.py:193
                    'soft_targets': None,
                    'commit_loss': loss,
                    'codebook_loss': loss,
                    'scale': scale[:, s:s + 1] if scale is not None else None,
                })
                # side channel for decode(): valid only while 'quantized' / 'scale' are still these tensors, unmodified
                fr._batched = (batched, len(frames), fr['quantized'], _version_of(fr['quantized']), fr['scale'],
                               _version_of(fr['scale']))
                frames.append(fr)
        return frames

    # ---- decode --------------------------------------------------------------------------------
    def _frame_fields(self, frame):
        """Accept the fork's dicts and, as a superset, upstream (codes, scale) tuples."""
        if isinstance(frame, dict):
            quantized, codes, scale = frame.get('quantized'), frame.get('codes'), frame.get('scale')
        else:
            codes, scale = frame
            quantized = None
        if quantized is None:
            assert codes is not None, "frame has neither 'quantized' nor 'codes'"
            quantized = self.quantizer.decode(codes.transpose(0, 1).contiguous())
        return quantized, scale

    @torch.no_grad()
    def decode(self, encoded_frames: tp.List[EncodedFrame]) -> torch.Tensor:
        """Same contract as reference model.py:212-246 (decodes ``frame['quantized']``, fork delta D3)."""
        segment_length = self.segment_length
        if segment_length is None:
            assert len(encoded_frames) == 1
        outs = self._decode_groups(encoded_frames)
        if segment_length is None:
            return outs[0][1]
        frames, lens = self._stack_segments(outs)
        return self._overlap_add(frames, lens)

    @torch.no_grad()
    def decode_segments(self, encoded_frames: tp.List[EncodedFrame]) -> tp.Tuple[torch.Tensor, tp.List[int]]:
        """Decoded segments WITHOUT the overlap-add: ``(frames [B, n, C, segment_length] zero-padded, lengths)``."""
        assert self.segment_length is not None
        return self._stack_segments(self._decode_groups(encoded_frames), self.segment_length)

    @torch.no_grad()
    def _decode_groups(self, encoded_frames: tp.List[EncodedFrame]):
        segment_length = self.segment_length
        # fast path: frames straight from our own encode() -> reuse the batched frames-major tensors
        # (the fork decodes frame['quantized'], delta D3, so editing or replacing the latents between encode and decode is
        # a supported use: the cached tensors are only used while the dict still holds the very tensors encode() put
        # there, with unchanged version counters)
        tag = [getattr(f, '_batched', None) for f in encoded_frames]
        groups = None
        if tag and all(t is not None for t in tag) and all(t[0] is tag[0][0] for t in tag) and \
                [t[1] for t in tag] == list(range(len(tag))) and \
                sum(g["n_seg"] for g in tag[0][0].groups) == len(tag) and \
                all(_untouched(f, t) for f, t in zip(encoded_frames, tag)):
            groups = tag[0][0].groups
        else:
            groups = []
            i = 0
            fields = [self._frame_fields(f) for f in encoded_frames]
            while i < len(fields):
                j = i
                while j + 1 < len(fields) and fields[j + 1][0].shape == fields[i][0].shape:
                    j += 1
                q = torch.stack([fields[k][0] for k in range(i, j + 1)], dim=1)  # [B, n_seg, D, T_f]
                b, n_seg, d, t_f = q.shape
                nat.require_cuda(q, "frame['quantized']")
                sc = None
                if fields[i][1] is not None:
                    sc = torch.stack([fields[k][1].reshape(-1) for k in range(i, j + 1)], dim=1).reshape(-1).contiguous()
                groups.append(dict(first=i, n_seg=n_seg, n_frames=t_f, quantized=q.reshape(b * n_seg, d, t_f).contiguous(),
                                   quantized_frames=None, scale=sc, batch=b))
                i = j + 1
        hop = self.decoder.hop_length
        outs = []
        for g in groups:
            b, n_seg, t_f = g["batch"], g["n_seg"], g["n_frames"]
            out = torch.empty((b * n_seg, self.channels, t_f * hop), dtype=torch.float32,
                              device=g["quantized"].device)
            bc = self._batch_chunk(n_seg, t_f * hop, b)
            for b0 in range(0, b, bc):
                b1 = min(b, b0 + bc)
                sl = slice(b0 * n_seg, b1 * n_seg)
                zf = g["quantized_frames"]
                z = None
                if zf is not None:
                    zf = zf[b0 * n_seg * t_f: b1 * n_seg * t_f]
                else:
                    z = g["quantized"][sl]
                sc = g["scale"][sl] if g["scale"] is not None else None
                self.decoder.decode_items(z, zf, (b1 - b0) * n_seg, t_f, sc, out[sl])
            outs.append((g, out))
        return outs

    def _stack_segments(self, outs, seg_len: tp.Optional[int] = None) -> tp.Tuple[torch.Tensor, tp.List[int]]:
        """Per segment-length group outputs -> one [B, n_seg, C, seg_len] tensor (short segments zero-padded) + lengths.
        ``seg_len`` defaults to the first frame's length, which is what the reference's overlap-add builds its window
        from (utils.py:40-44)."""
        g0, o0 = outs[0]
        b = g0["batch"]
        if seg_len is None:
            seg_len = o0.shape[-1]
        n_seg_total = sum(g["n_seg"] for g, _ in outs)
        dev = o0.device
        if len(outs) == 1 and o0.shape[-1] == seg_len:
            frames = o0.view(b, n_seg_total, self.channels, seg_len)  # [B * n_seg, C, seg_len] == [B, n_seg, C, seg_len]
            lens = [seg_len] * n_seg_total
        else:
            frames = torch.zeros((b, n_seg_total, self.channels, seg_len), dtype=torch.float32, device=dev)
            lens = []
            pos = 0
            for g, o in outs:
                n = o.shape[-1]
                frames[:, pos:pos + g["n_seg"], :, :n] = o.view(b, g["n_seg"], self.channels, n)
                lens += [n] * g["n_seg"]
                pos += g["n_seg"]
        return frames, lens

    def _overlap_add(self, frames: torch.Tensor, lens: tp.Sequence[int]) -> torch.Tensor:
        """utils._linear_overlap_add (reference utils.py:17-56) as one kernel over ``frames [B, n_seg, C, segment_length]``."""
        stride = self.segment_stride or 1
        b, n_seg_total, _, seg_len = frames.shape
        dev = frames.device
        total = stride * (n_seg_total - 1) + lens[-1]
        seg_lens = torch.tensor(list(lens), dtype=torch.int32, device=dev)
        out = torch.empty((b, self.channels, total), dtype=torch.float32, device=dev)
        frames = frames.contiguous()
        with torch.cuda.device(dev):
            nat.check(nat.lib.ecb_overlap_add(nat.ptr(frames), nat.ptr(seg_lens), b, self.channels, n_seg_total, seg_len,
                                              stride, nat.ptr(out), total, nat.stream_ptr(dev)))
        return out

    # ---- forward -------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        """Reference model.py:248-257: (audio[:, :, :T], codes [B,K,sum T_f], commit_loss, codebook_loss)."""
        frames = self.encode(x)
        codes = torch.cat([frame['codes'] for frame in frames], dim=-1)
        commit_loss = torch.cat([frame['commit_loss'] for frame in frames], dim=-1)
        codebook_loss = torch.cat([frame['codebook_loss'] for frame in frames], dim=-1)
        return self.decode(frames)[:, :, :x.shape[-1]], codes, commit_loss, codebook_loss

    def set_target_bandwidth(self, bandwidth: float):
        if bandwidth not in self.target_bandwidths:
            raise ValueError(f"This model doesn't support the bandwidth {bandwidth}. "
                             f"Select one of {self.target_bandwidths}.")
        self.bandwidth = bandwidth

    def get_lm_model(self, state_dict: tp.Optional[tp.Dict[str, torch.Tensor]] = None):
        """Return the associated LM (reference model.py:264-284): ``LMModel(n_q, bins, num_layers=5, dim=200,
        past_context=int(3.5 * frame_rate))`` on the model's device, in eval mode. The reference downloads the pre-trained
        checkpoint named after ``self.name``; pass ``state_dict`` to load weights held locally instead (no network here)."""
        from .lm import LMModel
        device = next(self.parameters()).device
        lm = LMModel(self.quantizer.n_q, self.quantizer.bins, num_layers=5, dim=200,
                     past_context=int(3.5 * self.frame_rate)).to(device)
        if state_dict is None:
            checkpoints = {'encodec_24khz': 'encodec_lm_24khz-1608e3c0.th', 'encodec_48khz': 'encodec_lm_48khz-7add9fc3.th'}
            try:
                checkpoint_name = checkpoints[self.name]
            except KeyError:
                raise RuntimeError("No LM pre-trained for the current Encodec model.")
            state_dict = torch.hub.load_state_dict_from_url(ROOT_URL + checkpoint_name, map_location='cpu', check_hash=True)
        lm.load_state_dict(state_dict)
        lm.eval()
        return lm

    # ---- factories (reference model.py:286-382) ------------------------------------------------------
    @staticmethod
    def _get_model(target_bandwidths: tp.List[float],
                   sample_rate: int = 10,
                   channels: int = 1,
                   causal: bool = True,
                   model_norm: str = 'weight_norm',
                   audio_normalize: bool = False,
                   segment: tp.Optional[float] = None,
                   name: str = 'breathing_model',
                   ratios=[8, 5, 4, 2],
                   bins=256,
                   dimension=128,
                   codebook_dim=32,
                   share_codebook: bool = True):
        encoder = SEANetEncoder(channels=channels, norm=model_norm, causal=causal, ratios=ratios, dimension=dimension)
        decoder = SEANetDecoder(channels=channels, norm=model_norm, causal=causal, ratios=ratios, dimension=dimension)
        n_q = int(1000 * target_bandwidths[-1] // (math.ceil(sample_rate / encoder.hop_length) * 10))
        quantizer = ResidualVectorQuantizer(
            dimension=encoder.dimension,
            n_q=n_q,
            bins=bins,
            codebook_dim=encoder.dimension,  # the reference ignores its own codebook_dim argument (model.py:303-308)
            share_codebook=share_codebook,
        )
        return EncodecModel(encoder, decoder, quantizer, target_bandwidths, sample_rate, channels,
                            normalize=audio_normalize, segment=segment, name=name)

    @staticmethod
    def encodec_model_24khz(pretrained: bool = False, repository=None, bins: int = 256, share_codebook: bool = True):
        """Causal 24 kHz model (reference model.py:344-362). No network here: ``pretrained`` must be False.
        Like the fork, the default builds 256-entry codebooks (delta D5); pass ``bins=1024`` for upstream."""
        if pretrained or repository:
            raise NotImplementedError("encodec_b200: pretrained checkpoints cannot be fetched (no network); "
                                      "build the model and call load_state_dict")
        model = EncodecModel._get_model([1.5, 3., 6, 12., 24.], 24_000, 1, causal=True, model_norm='weight_norm',
                                        audio_normalize=False, name='unset', bins=bins, share_codebook=share_codebook)
        model.eval()
        return model

    @staticmethod
    def encodec_model_48khz(pretrained: bool = False, repository=None, bins: int = 256, share_codebook: bool = True):
        """Non-causal stereo 48 kHz model (reference model.py:364-382)."""
        if pretrained or repository:
            raise NotImplementedError("encodec_b200: pretrained checkpoints cannot be fetched (no network); "
                                      "build the model and call load_state_dict")
        model = EncodecModel._get_model([3., 6., 12., 24.], 48_000, 2, causal=False, model_norm='time_group_norm',
                                        audio_normalize=True, segment=1., name='unset', bins=bins,
                                        share_codebook=share_codebook)
        model.eval()
        return model
