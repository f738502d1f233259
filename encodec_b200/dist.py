"""Multi-GPU plumbing: one process per GPU, clips sharded across ranks, one gather at the end.

The codec path has no exchange step: batch items and the 48 kHz model's segments
are independent, so ranks never talk while encoding/decoding. The only collective is the gather of codes
and audio to one rank (north_star), done with ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the
CPU tests). Codes travel as int32 and are widened to the reference's int64 on the root.
"""
from __future__ import annotations

import typing as tp

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tp.Tuple[int, int]:
    """Contiguous split of ``n`` clips: rank r gets [lo, hi); the first ``n % world`` ranks get one more."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world() -> tp.Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_results(codes: torch.Tensor, audio: tp.Optional[torch.Tensor], dst: int = 0,
                   counts: tp.Optional[tp.Sequence[int]] = None):
    """Gather per-rank ``codes [b_r, K, T]`` (int64) and ``audio [b_r, C, L]`` on ``dst``.

    ``counts`` (clips per rank) allows ragged shards; by default every rank holds the same number of clips.
    Returns ``(codes_all, audio_all)`` on ``dst`` and ``(None, None)`` elsewhere.
    """
    rank, world = _world()
    if world == 1:
        return codes, audio
    wire = codes.to(torch.int32).contiguous()  # neither NCCL nor gloo moves int16; int32 halves the int64 bytes
    if counts is None:
        counts = [codes.shape[0]] * world
    b_max = max(counts)

    def padded(t):
        if t.shape[0] == b_max:
            return t
        pad = torch.zeros((b_max - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        return torch.cat([t, pad], dim=0)

    wire = padded(wire)
    code_list = [torch.empty_like(wire) for _ in range(world)] if rank == dst else None
    dist.gather(wire, code_list, dst=dst)
    audio_list = None
    if audio is not None:
        a = padded(audio.contiguous())
        audio_list = [torch.empty_like(a) for _ in range(world)] if rank == dst else None
        dist.gather(a, audio_list, dst=dst)
    if rank != dst:
        return None, None
    codes_all = torch.cat([c[:n] for c, n in zip(code_list, counts)], dim=0).to(torch.int64)
    audio_all = torch.cat([a[:n] for a, n in zip(audio_list, counts)], dim=0) if audio_list is not None else None
    return codes_all, audio_all


def forward_sharded(model, x: torch.Tensor, dst: int = 0):
    """``model.forward`` on this rank's contiguous shard of the clips ``x [B, C, T]`` (every rank passes the
    same ``x``, or at least its own rows), then one gather. Returns the full ``(audio, codes)`` on ``dst``."""
    rank, world = _world()
    lo, hi = shard_range(x.shape[0], rank, world)
    counts = [shard_range(x.shape[0], r, world) for r in range(world)]
    counts = [b - a for a, b in counts]
    if hi > lo:
        audio, codes, _, _ = model(x[lo:hi])
    else:
        audio = codes = None
    if world == 1:
        return audio, codes
    # ranks with an empty shard still take part in the collective with zero-row tensors
    shapes = [None]
    if rank == dst:
        shapes = [(tuple(codes.shape[1:]), tuple(audio.shape[1:]))]
    dist.broadcast_object_list(shapes, src=dst)
    if audio is None:
        dev = x.device
        codes = torch.zeros((0,) + shapes[0][0], dtype=torch.int64, device=dev)
        audio = torch.zeros((0,) + shapes[0][1], dtype=torch.float32, device=dev)
    codes_all, audio_all = gather_results(codes, audio, dst=dst, counts=counts)
    return audio_all, codes_all
