"""`.ecdc` streams without entropy coding -- the caller right after ``EncodecModel.encode``.

Mirrors ``compress_to_file / decompress_from_file / compress / decompress`` of the reference (compress.py:28-185) for
``use_lm=False``: header (binary.py:23-52) + per frame an optional big-endian float32 scale + the codes packed
``bits_per_codebook`` bits each, time-major, LSB first (binary.BitPacker, binary.py:55-89). The packing / unpacking runs
on the GPU (csrc/bitpack.cu) and is byte-exact with the reference's BitPacker. Differences: ``use_lm=True`` (LM +
arithmetic coder, out of scope) raises; ``decompress*`` takes the model instead of building a pretrained one from its
name (the fork's factories cannot load the upstream checkpoints, SURVEY.md delta D11).
"""
from __future__ import annotations

import io
import json
import math
import struct
import typing as tp

import torch

from . import _native as nat

_HEADER = struct.Struct("!4sBI")   # magic, protocol version, header size (binary.py:19-20)
_MAGIC = b"ECDC"


def write_ecdc_header(fo: tp.IO[bytes], metadata: tp.Any) -> None:
    """binary.write_ecdc_header (binary.py:23-29)."""
    meta = json.dumps(metadata).encode("utf-8")
    fo.write(_HEADER.pack(_MAGIC, 0, len(meta)))
    fo.write(meta)
    fo.flush()


def _read_exactly(fo: tp.IO[bytes], size: int) -> bytes:
    buf = b""
    while len(buf) < size:
        new = fo.read(size - len(buf))
        if not new:
            raise EOFError(f"Impossible to read enough data from the stream, {size - len(buf)} bytes remaining.")
        buf += new
    return buf


def read_ecdc_header(fo: tp.IO[bytes]):
    """binary.read_ecdc_header (binary.py:45-52)."""
    magic, version, size = _HEADER.unpack(_read_exactly(fo, _HEADER.size))
    if magic != _MAGIC:
        raise ValueError("File is not in ECDC format.")
    if version != 0:
        raise ValueError("Version not supported.")
    return json.loads(_read_exactly(fo, size).decode("utf-8"))


def pack_codes(codes: torch.Tensor, bits: int) -> torch.Tensor:
    """codes [K, T] int64 (CUDA, any strides) -> uint8 [ceil(K*T*bits/8)]: what BitPacker writes for one frame."""
    nat.require_cuda(codes, "codes", torch.int64)
    assert codes.dim() == 2
    k, t = codes.shape
    out = torch.empty(int(nat.lib.ecb_packed_bytes(k, t, bits)), dtype=torch.uint8, device=codes.device)
    with torch.cuda.device(codes.device):
        nat.check(nat.lib.ecb_pack_codes(nat.ptr(codes), codes.stride(0), codes.stride(1), k, t, bits, nat.ptr(out),
                                         nat.stream_ptr(codes.device)))
    return out


def unpack_codes(packed: torch.Tensor, n_codebooks: int, n_frames: int, bits: int) -> torch.Tensor:
    """uint8 stream (CUDA) -> codes [K, T] int64: what BitUnpacker.pull yields, K values per time step."""
    nat.require_cuda(packed, "packed", torch.uint8)
    codes = torch.empty((n_codebooks, n_frames), dtype=torch.int64, device=packed.device)
    with torch.cuda.device(packed.device):
        nat.check(nat.lib.ecb_unpack_codes(nat.ptr(packed), packed.numel(), n_codebooks, n_frames, bits, nat.ptr(codes),
                                           codes.stride(0), codes.stride(1), nat.stream_ptr(packed.device)))
    return codes


def compress_to_file(model, wav: torch.Tensor, fo: tp.IO[bytes], use_lm: bool = False) -> None:
    """compress.compress_to_file (compress.py:28-89) for one waveform ``wav [C, T]`` on the model's device."""
    assert wav.dim() == 2, "Only single waveform can be encoded."
    if use_lm:
        raise NotImplementedError("encodec_b200: entropy coding with the language model is not implemented")
    with torch.no_grad():
        frames = model.encode(wav[None])
    metadata = {"m": model.name, "al": wav.shape[-1], "nc": int(frames[0]["codes"].shape[1]), "lm": False}
    write_ecdc_header(fo, metadata)
    for frame in frames:
        if frame["scale"] is not None:
            fo.write(struct.pack("!f", frame["scale"].cpu().item()))
        fo.write(pack_codes(frame["codes"][0], model.bits_per_codebook).cpu().numpy().tobytes())


def decompress_from_file(fo: tp.IO[bytes], model) -> tp.Tuple[torch.Tensor, int]:
    """compress.decompress_from_file (compress.py:92-156): returns ``(wav [C, T], sample_rate)``."""
    metadata = read_ecdc_header(fo)
    audio_length, num_codebooks = metadata["al"], metadata["nc"]
    assert isinstance(audio_length, int) and isinstance(num_codebooks, int)
    if metadata["lm"]:
        raise NotImplementedError("encodec_b200: entropy-coded streams are not supported")
    device = next(model.parameters()).device
    bits = model.bits_per_codebook
    frames = []
    segment_length = model.segment_length or audio_length
    segment_stride = model.segment_stride or audio_length
    for offset in range(0, audio_length, segment_stride):
        this_len = min(audio_length - offset, segment_length)
        frame_length = int(math.ceil(this_len * model.frame_rate / model.sample_rate))
        scale = None
        if model.normalize:
            scale_f, = struct.unpack("!f", _read_exactly(fo, 4))
            scale = torch.tensor(scale_f, device=device).view(1)
        raw = _read_exactly(fo, (num_codebooks * frame_length * bits + 7) // 8)
        packed = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)
        frames.append((unpack_codes(packed, num_codebooks, frame_length, bits)[None], scale))
    with torch.no_grad():
        wav = model.decode(frames)
    return wav[0, :, :audio_length], model.sample_rate


def compress(model, wav: torch.Tensor, use_lm: bool = False) -> bytes:
    """compress.compress (compress.py:159-173)."""
    fo = io.BytesIO()
    compress_to_file(model, wav, fo, use_lm=use_lm)
    return fo.getvalue()


def decompress(compressed: bytes, model) -> tp.Tuple[torch.Tensor, int]:
    """compress.decompress (compress.py:176-185)."""
    return decompress_from_file(io.BytesIO(compressed), model)
