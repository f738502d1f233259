"""`.ecdc` streams -- the caller right after ``EncodecModel.encode``.

Mirrors ``compress_to_file / decompress_from_file / compress / decompress`` of the reference (compress.py:28-185):
header (binary.py:23-52) + per frame an optional big-endian float32 scale + the codes, either packed
``bits_per_codebook`` bits each, time-major, LSB first (``use_lm=False``: binary.BitPacker, binary.py:55-89; packing /
unpacking on the GPU, csrc/bitpack.cu, byte-exact with the reference's BitPacker) or entropy-coded (``use_lm=True``:
language model + arithmetic coder, ``encodec_b200.lm`` / csrc/lm.cu -- compression runs the LM over all steps of all frames
in one batched pass and codes on the host; decompression keeps the LM step / cdf / decoder-pull loop on the device).
Differences: ``decompress*`` takes the model (and the LM) instead of building pretrained ones from the stream's model name
(the fork's factories cannot load the upstream checkpoints, SURVEY.md delta D11, and there is no network); ``lm=`` lets the
caller supply the language model that ``model.get_lm_model()`` would otherwise have to download.
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


def _entropy_code_frames(lm, frames) -> tp.List[bytes]:
    """The coder's bytes of every frame (compress.py:66-87 with use_lm): frames of equal length share one batched LM pass."""
    out: tp.List[tp.Optional[bytes]] = [None] * len(frames)
    by_len: tp.Dict[int, tp.List[int]] = {}
    for i, frame in enumerate(frames):
        by_len.setdefault(int(frame["codes"].shape[-1]), []).append(i)
    for idxs in by_len.values():
        codes = torch.cat([frames[i]["codes"][:1] for i in idxs], dim=0).contiguous()      # [n, K, T]
        for i, data in zip(idxs, lm.encode_frames(codes)):
            out[i] = data
    return out  # type: ignore[return-value]


def compress_to_file(model, wav: torch.Tensor, fo: tp.IO[bytes], use_lm: bool = False, lm=None) -> None:
    """compress.compress_to_file (compress.py:28-89) for one waveform ``wav [C, T]`` on the model's device. ``lm``: the
    language model to use with ``use_lm=True`` (default ``model.get_lm_model()``, compress.py:49-50)."""
    assert wav.dim() == 2, "Only single waveform can be encoded."
    if use_lm and lm is None:
        lm = model.get_lm_model()
    with torch.no_grad():
        frames = model.encode(wav[None])
    metadata = {"m": model.name, "al": wav.shape[-1], "nc": int(frames[0]["codes"].shape[1]), "lm": bool(use_lm)}
    write_ecdc_header(fo, metadata)
    coded = _entropy_code_frames(lm, frames) if use_lm else None
    for i, frame in enumerate(frames):
        if frame["scale"] is not None:
            fo.write(struct.pack("!f", frame["scale"].cpu().item()))
        if use_lm:
            fo.write(coded[i])
        else:
            fo.write(pack_codes(frame["codes"][0], model.bits_per_codebook).cpu().numpy().tobytes())


def _decompress_entropy_coded(fo: tp.IO[bytes], model, lm, audio_length: int, num_codebooks: int):
    """compress.py:114-152 with use_lm: the frames follow each other without a length field, so they are decoded in order;
    within a frame the whole step loop runs on the device (``LMModel.decode_frame``)."""
    device = next(model.parameters()).device
    raw = fo.read()
    data = torch.frombuffer(bytearray(raw) if raw else bytearray(1), dtype=torch.uint8).to(device)
    pos = 0
    frames = []
    segment_length = model.segment_length or audio_length
    segment_stride = model.segment_stride or audio_length
    for offset in range(0, audio_length, segment_stride):
        this_len = min(audio_length - offset, segment_length)
        frame_length = int(math.ceil(this_len * model.frame_rate / model.sample_rate))
        scale = None
        if model.normalize:
            if pos + 4 > len(raw):
                raise EOFError(f"Impossible to read enough data from the stream, {pos + 4 - len(raw)} bytes remaining.")
            scale_f, = struct.unpack("!f", raw[pos: pos + 4])
            scale = torch.tensor(scale_f, device=device).view(1)
            pos += 4
        if pos >= len(raw):
            raise EOFError("The stream ended sooner than expected.")
        codes, pos = lm.decode_frame(data[: len(raw)], pos, num_codebooks, frame_length)
        frames.append((codes[None], scale))
    return frames


def decompress_from_file(fo: tp.IO[bytes], model, lm=None) -> tp.Tuple[torch.Tensor, int]:
    """compress.decompress_from_file (compress.py:92-156): returns ``(wav [C, T], sample_rate)``."""
    metadata = read_ecdc_header(fo)
    audio_length, num_codebooks = metadata["al"], metadata["nc"]
    assert isinstance(audio_length, int) and isinstance(num_codebooks, int)
    if metadata["lm"]:
        if lm is None:
            lm = model.get_lm_model()
        frames = _decompress_entropy_coded(fo, model, lm, audio_length, num_codebooks)
        with torch.no_grad():
            wav = model.decode(frames)
        return wav[0, :, :audio_length], model.sample_rate
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


def compress(model, wav: torch.Tensor, use_lm: bool = False, lm=None) -> bytes:
    """compress.compress (compress.py:159-173)."""
    fo = io.BytesIO()
    compress_to_file(model, wav, fo, use_lm=use_lm, lm=lm)
    return fo.getvalue()


def decompress(compressed: bytes, model, lm=None) -> tp.Tuple[torch.Tensor, int]:
    """compress.decompress (compress.py:176-185)."""
    return decompress_from_file(io.BytesIO(compressed), model, lm=lm)
