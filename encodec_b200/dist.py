"""Multi-GPU plumbing: one process per GPU, clips sharded across ranks, one gather per (micro-)batch.

The codec path has no exchange step: batch items and the 48 kHz model's segments
are independent, so ranks never talk while encoding/decoding. The only collective is the gather of codes
and audio to one rank (north_star), done with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the
CPU tests). Codes and audio of a (micro-)batch travel as ONE packed byte buffer per rank -- codes as int16 when
the codebook size allows it (10-bit codes), else int32, widened to the reference's int64 on the root -- and the
gather is issued on a side stream behind an event, so the transfer of micro-batch i hides under the kernels of
micro-batch i + 1 (``GatherQueue``).
"""
from __future__ import annotations

import math
import typing as tp

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Contiguous split of ``n`` clips: rank r gets [lo, hi); the first ``n % world`` ranks get one more."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_counts(n: int, world: int) -> tp.List[int]:
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def _world() -> tp.Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _code_wire_dtype(bins: tp.Optional[int]) -> torch.dtype:
    return torch.int16 if bins is not None and bins <= 32768 else torch.int32


def pack_results(codes: torch.Tensor, audio: tp.Optional[torch.Tensor], rows: int, bins: tp.Optional[int] = None) -> torch.Tensor:
    """One contiguous uint8 buffer holding ``rows`` clips: [codes (wire dtype) | audio (fp32)], zero-padded when this
    rank holds fewer than ``rows`` clips."""
    wd = _code_wire_dtype(bins)
    c = codes.to(wd).contiguous()
    per_c = math.prod(c.shape[1:]) * c.element_size()
    per_a = math.prod(audio.shape[1:]) * 4 if audio is not None else 0
    off_a = (rows * per_c + 15) // 16 * 16   # 16-byte aligned audio section
    buf = torch.zeros(off_a + rows * per_a, dtype=torch.uint8, device=codes.device)
    n = c.shape[0]
    if n:
        buf[:n * per_c].copy_(c.view(-1).view(torch.uint8))
        if audio is not None:
            buf[off_a:off_a + n * per_a].copy_(audio.contiguous().view(-1).view(torch.uint8))
    return buf


def unpack_results(buf: torch.Tensor, n: int, rows: int, code_shape: tp.Sequence[int], audio_shape: tp.Optional[tp.Sequence[int]],
                   bins: tp.Optional[int] = None):
    """Inverse of ``pack_results`` for the first ``n`` of ``rows`` clips: (codes int64 [n, *code_shape], audio fp32 or None)."""
    wd = _code_wire_dtype(bins)
    per_c = math.prod(code_shape) * torch.empty((), dtype=wd).element_size()
    off_a = (rows * per_c + 15) // 16 * 16
    codes = buf[:n * per_c].view(wd).view((n,) + tuple(code_shape)).to(torch.int64)
    audio = None
    if audio_shape is not None:
        per_a = math.prod(audio_shape) * 4
        audio = buf[off_a:off_a + n * per_a].view(torch.float32).view((n,) + tuple(audio_shape)).clone()
    return codes, audio


def gather_results(codes: torch.Tensor, audio: tp.Optional[torch.Tensor], dst: int = 0,
                   counts: tp.Optional[tp.Sequence[int]] = None, bins: tp.Optional[int] = None):
    """Gather per-rank ``codes [b_r, K, T]`` (int64) and ``audio [b_r, C, L]`` on ``dst`` with ONE collective.

    ``counts`` (clips per rank) allows ragged shards; by default every rank holds the same number of clips.
    Returns ``(codes_all, audio_all)`` on ``dst`` and ``(None, None)`` elsewhere. Issued on the current stream.
    """
    rank, world = _world()
    if world == 1:
        return codes, audio
    if counts is None:
        counts = [codes.shape[0]] * world
    rows = max(counts)
    buf = pack_results(codes, audio, rows, bins)
    recv = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, recv, dst=dst)
    if rank != dst:
        return None, None
    a_shape = tuple(audio.shape[1:]) if audio is not None else None
    parts = [unpack_results(b, n, rows, tuple(codes.shape[1:]), a_shape, bins) for b, n in zip(recv, counts)]
    codes_all = torch.cat([p[0] for p in parts], dim=0)
    audio_all = torch.cat([p[1] for p in parts], dim=0) if audio is not None else None
    return codes_all, audio_all


class GatherQueue:
    """Gathers of successive (micro-)batches on a side stream: ``submit`` returns at once, the collective waits (on the
    device) for the kernels that produced its inputs and runs while the next batch computes; ``finish`` makes the current
    stream wait for all of them and returns the per-batch results on ``dst``."""

    def __init__(self, device: torch.device, dst: int = 0, bins: tp.Optional[int] = None):
        self.device = device
        self.dst = dst
        self.bins = bins
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self._results: tp.List[tp.Tuple[tp.Optional[torch.Tensor], tp.Optional[torch.Tensor]]] = []

    def submit(self, codes: torch.Tensor, audio: tp.Optional[torch.Tensor], counts: tp.Optional[tp.Sequence[int]] = None):
        if self.stream is None:
            self._results.append(gather_results(codes, audio, self.dst, counts, self.bins))
            return
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            self._results.append(gather_results(codes, audio, self.dst, counts, self.bins))
        for t in (codes, audio):   # the producers' memory must outlive the side stream's reads
            if t is not None:
                t.record_stream(self.stream)

    def keep_last(self, n: int):
        """Drop all but the last ``n`` gathered results (a benchmark loop that only wants the traffic, not the data)."""
        if len(self._results) > n:
            del self._results[:len(self._results) - n]

    def finish(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
        out, self._results = self._results, []
        return out


def output_shapes(model, x: torch.Tensor) -> tp.Tuple[tp.Tuple[int, ...], tp.Tuple[int, ...]]:
    """Per-clip shapes of ``model(x)``'s codes and audio, from the model's geometry alone (every rank can size the
    collective's buffers without having run anything -- ranks with an empty shard included)."""
    length = x.shape[-1]
    hop = model.encoder.hop_length
    segs, _ = model._segments(length)
    n_frames = sum(-(-n // hop) for _, n in segs)
    return (model._n_q(), n_frames), (x.shape[1], length)


def forward_sharded(model, x: torch.Tensor, dst: int = 0, micro_batch: tp.Optional[int] = None, overlap: bool = True):
    """``model.forward`` on this rank's contiguous shard of the clips ``x [B, C, T]`` (every rank passes the same ``x``, or
    at least its own rows), in micro-batches of ``micro_batch`` clips, each followed by one gather -- on a side stream when
    ``overlap`` (the transfer hides under the next micro-batch). Returns the full ``(audio, codes)`` on ``dst``, in clip
    order, and ``(None, None)`` elsewhere."""
    rank, world = _world()
    lo, hi = shard_range(x.shape[0], rank, world)
    return forward_shard(model, x[lo:hi], shard_counts(x.shape[0], world), dst, micro_batch, overlap)


def forward_shard(model, x_local: torch.Tensor, counts: tp.Sequence[int], dst: int = 0, micro_batch: tp.Optional[int] = None,
                  overlap: bool = True):
    """Same as ``forward_sharded`` for callers that hold only their own rows: ``x_local [counts[rank], C, T]``, ``counts`` =
    clips per rank (known to everybody; shards may be ragged or empty)."""
    rank, world = _world()
    assert len(counts) == world and x_local.shape[0] == counts[rank], (x_local.shape, counts, rank)
    mb = micro_batch or max(1, max(counts))
    bins = getattr(getattr(model, "quantizer", None), "bins", None)
    if world == 1:
        outs = [model(x_local[i:i + mb])[:2] for i in range(0, counts[0], mb)]
        if not outs:
            return None, None
        return torch.cat([o[0] for o in outs], dim=0), torch.cat([o[1] for o in outs], dim=0)
    code_shape, audio_shape = output_shapes(model, x_local)
    queue = GatherQueue(x_local.device, dst, bins) if overlap else None
    results = []
    rounds = -(-max(counts) // mb)
    for r in range(rounds):
        part = [max(0, min(mb, c - r * mb)) for c in counts]   # clips every rank contributes in this round
        if part[rank] > 0:
            audio, codes, _, _ = model(x_local[r * mb:r * mb + part[rank]])
        else:   # an empty shard still takes part in the collective, with zero-row tensors
            codes = torch.zeros((0,) + code_shape, dtype=torch.int64, device=x_local.device)
            audio = torch.zeros((0,) + audio_shape, dtype=torch.float32, device=x_local.device)
        if queue is not None:
            queue.submit(codes, audio, part)
        else:
            results.append(gather_results(codes, audio, dst, part, bins))
    if queue is not None:
        results = queue.finish()
    if rank != dst:
        return None, None
    # round r holds, for every rank, its r-th micro-batch: restore clip order (rank-major)
    per_rank_codes = [[] for _ in range(world)]
    per_rank_audio = [[] for _ in range(world)]
    for r, (c_all, a_all) in enumerate(results):
        part = [max(0, min(mb, c - r * mb)) for c in counts]
        pos = 0
        for k in range(world):
            per_rank_codes[k].append(c_all[pos:pos + part[k]])
            per_rank_audio[k].append(a_all[pos:pos + part[k]])
            pos += part[k]
    codes_all = torch.cat([t for k in range(world) for t in per_rank_codes[k]], dim=0)
    audio_all = torch.cat([t for k in range(world) for t in per_rank_audio[k]], dim=0)
    return audio_all, codes_all


def segment_shards(n_clips: int, n_seg: int, world: int) -> tp.List[tp.List[tp.Tuple[int, int, int]]]:
    """Contiguous split of the flattened (clip, segment) list over ``world`` ranks: for every rank a list of
    ``(clip, first segment, end segment)`` pieces (when there are fewer clips than GPUs the segments
    of a clip, which are independent, are what gets sharded)."""
    out = []
    for r in range(world):
        lo, hi = shard_range(n_clips * n_seg, r, world)
        pieces = []
        while lo < hi:
            clip, s0 = divmod(lo, n_seg)
            s1 = min(n_seg, s0 + (hi - lo))
            pieces.append((clip, s0, s1))
            lo += s1 - s0
        out.append(pieces)
    return out


def encode_decode_pieces(model, x: torch.Tensor, pieces: tp.Sequence[tp.Tuple[int, int, int]]):
    """This rank's part of ``forward_sharded_segments``: for every ``(clip, s0, s1)`` piece the codes of its segments
    (``[n, K, T_f max]`` int64, zero-padded) and their decoded audio (``[n, C, segment_length]``, zero-padded)."""
    c = x.shape[1]
    hop, seg_len, n_q = model.encoder.hop_length, model.segment_length, model._n_q()
    tf_max = -(-seg_len // hop)
    dev = x.device
    codes_parts, audio_parts = [], []
    for clip, s0, s1 in pieces:
        frames = model.encode_segments(x[clip:clip + 1], s0, s1)
        seg_audio, _ = model.decode_segments(frames)                        # [1, n, C, seg_len]
        cp = torch.zeros((s1 - s0, n_q, tf_max), dtype=torch.int64, device=dev)
        for k, f in enumerate(frames):
            cp[k, :, :f["codes"].shape[-1]] = f["codes"][0]
        codes_parts.append(cp)
        audio_parts.append(seg_audio[0])
    if codes_parts:
        return torch.cat(codes_parts, dim=0), torch.cat(audio_parts, dim=0)
    return (torch.zeros((0, n_q, tf_max), dtype=torch.int64, device=dev),
            torch.zeros((0, c, seg_len), dtype=torch.float32, device=dev))


def assemble_segments(model, codes_all: torch.Tensor, audio_all: torch.Tensor, b: int, c: int, length: int):
    """The destination rank's part: ``codes_all [B * n_seg, K, T_f max]`` / ``audio_all [B * n_seg, C, segment_length]`` in
    (clip, segment) order -> ``(audio [B, C, T], codes [B, K, sum T_f])`` with the reference's linear overlap-add."""
    segs, _ = model._segments(length)
    n_seg, hop, seg_len = len(segs), model.encoder.hop_length, model.segment_length
    lens = [n for _, n in segs]
    t_f = [-(-n // hop) for n in lens]
    codes_all = codes_all.view(b, n_seg, codes_all.shape[1], codes_all.shape[2])
    codes = torch.cat([codes_all[:, s, :, :t_f[s]] for s in range(n_seg)], dim=-1)
    audio = model._overlap_add(audio_all.view(b, n_seg, c, seg_len), lens)
    return audio[:, :, :length], codes


def forward_sharded_segments(model, x: torch.Tensor, dst: int = 0):
    """``model.forward`` of a SEGMENTED model (48 kHz: 1 s segments, 1 % overlap) with the (clip, segment) list sharded over
    the ranks -- the partition for fewer clips than GPUs. Every rank encodes and decodes its segments
    (``encode_segments`` / ``decode_segments``), codes and decoded segments are gathered on ``dst`` (one collective), and the
    linear overlap-add (reference utils.py:17-56) runs there over the complete segment list of every clip. Returns
    ``(audio [B, C, T], codes [B, K, sum T_f])`` on ``dst`` -- bit-identical to ``model(x)`` -- and ``(None, None)`` elsewhere."""
    rank, world = _world()
    assert model.segment_length is not None, "forward_sharded_segments needs a segmented model"
    b, c, length = x.shape
    n_seg = len(model._segments(length)[0])
    shards = segment_shards(b, n_seg, world)
    counts = [sum(s1 - s0 for _, s0, s1 in pieces) for pieces in shards]
    codes_r, audio_r = encode_decode_pieces(model, x, shards[rank])
    codes_all, audio_all = gather_results(codes_r, audio_r, dst, counts, getattr(model.quantizer, "bins", None))
    if rank != dst:
        return None, None
    return assemble_segments(model, codes_all, audio_all, b, c, length)
