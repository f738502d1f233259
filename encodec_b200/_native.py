"""ctypes binding of libencodec_b200.so (the C ABI declared in include/encodec_b200.h).

There is no CPU fallback: if the library is missing this module raises at import time, and every op
raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libencodec_b200.so")

MAX_RATIOS = 8


class EcbSpec(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("causal", C.c_int32), ("group_norm", C.c_int32), ("n_filters", C.c_int32),
        ("dimension", C.c_int32), ("n_ratios", C.c_int32), ("ratios", C.c_int32 * MAX_RATIOS),
        ("kernel_size", C.c_int32), ("last_kernel_size", C.c_int32), ("residual_kernel_size", C.c_int32),
        ("compress", C.c_int32), ("lstm_layers", C.c_int32), ("bins", C.c_int32), ("n_q", C.c_int32),
    ]


class EcbLmSpec(C.Structure):
    _fields_ = [("n_q", C.c_int32), ("card", C.c_int32), ("dim", C.c_int32), ("n_layers", C.c_int32),
                ("n_heads", C.c_int32), ("hidden", C.c_int32), ("past_context", C.c_int32), ("max_period", C.c_float)]


class EcbProfEntry(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "ecb_last_error": (C.c_char_p, []),
    "ecb_version": (C.c_int, []),
    "ecb_launch_count": (C.c_int64, []),
    "ecb_debug_tap": (None, [C.c_void_p, C.c_int64, C.c_int32]),
    "ecb_debug_lstm_trace": (None, [C.c_void_p]),
    "ecb_f16_saturation_count": (C.c_int64, [C.c_int32]),
    "ecb_debug_lstm_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_debug_lstm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ecb_debug_tc_conv": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "ecb_profile_begin": (None, []),
    "ecb_profile_end": (C.c_int, [C.POINTER(EcbProfEntry), C.c_int]),
    "ecb_codec_create": (C.c_int, [C.POINTER(EcbSpec), C.POINTER(C.c_void_p)]),
    "ecb_codec_destroy": (None, [C.c_void_p]),
    "ecb_codec_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecb_codec_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ecb_codec_set_decoder_precision": (C.c_int, [C.c_void_p, C.c_int32]),
    "ecb_encoder_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_encoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    "ecb_decoder_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_decoder_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ecb_rvq_prepare": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "ecb_rvq_encode_frames": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ecb_rvq_decode_frames": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p]),
    "ecb_codec_rvq_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_codec_rvq_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]),
    "ecb_codec_rvq_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "ecb_overlap_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                  C.c_void_p, C.c_int64, C.c_void_p]),
    "ecb_packed_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int32]),
    "ecb_pack_codes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "ecb_unpack_codes": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.c_int64,
                                   C.c_void_p]),
    "ecb_lm_create": (C.c_int, [C.POINTER(EcbLmSpec), C.POINTER(C.c_void_p)]),
    "ecb_lm_destroy": (None, [C.c_void_p]),
    "ecb_lm_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ecb_lm_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ecb_lm_cache_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_lm_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int64, C.c_int64]),
    "ecb_lm_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_size_t, C.c_void_p]),
    "ecb_lm_decode_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ecb_ac_decode_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "ecb_quantized_cdf": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ecb_ac_encode": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "ecb_ac_decode": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                C.POINTER(C.c_int64)]),
    "ecb_transpose_bct_to_btc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "ecb_transpose_btc_to_bct": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: encodec_b200 has no CPU or PyTorch fallback. Build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python encodec_b200/_build.py).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.ecb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"encodec_b200: {last_error()}")


def launch_count() -> int:
    return int(lib.ecb_launch_count())


def stream_ptr(device: torch.device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else int(t.data_ptr())


def require_cuda(t: torch.Tensor, what: str, dtype=torch.float32) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"encodec_b200: {what} must be a CUDA tensor (there is no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"encodec_b200: {what} must be {dtype}, got {t.dtype}")


def make_spec(channels: int, causal: bool, group_norm: int, n_filters: int, dimension: int, ratios, kernel_size: int,
              last_kernel_size: int, residual_kernel_size: int, compress: int, lstm_layers: int, bins: int,
              n_q: int) -> EcbSpec:
    if len(ratios) > MAX_RATIOS:
        raise NotImplementedError(f"at most {MAX_RATIOS} ratios are supported")
    s = EcbSpec()
    s.channels, s.causal, s.group_norm = int(channels), int(bool(causal)), int(group_norm)   # 0 wn, 1 GroupNorm, 2 LayerNorm
    s.n_filters, s.dimension, s.n_ratios = int(n_filters), int(dimension), len(ratios)
    for i, r in enumerate(ratios):
        s.ratios[i] = int(r)
    s.kernel_size, s.last_kernel_size = int(kernel_size), int(last_kernel_size)
    s.residual_kernel_size, s.compress, s.lstm_layers = int(residual_kernel_size), int(compress), int(lstm_layers)
    s.bins, s.n_q = int(bins), int(n_q)
    return s


class Codec:
    """Owns one native ``ecb_codec`` handle (prepared weights on one device)."""

    def __init__(self, spec: EcbSpec, device: torch.device):
        self.device = torch.device(device)
        self.spec = spec
        h = C.c_void_p()
        check(lib.ecb_codec_create(C.byref(spec), C.byref(h)))
        self._h = h
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                lib.ecb_codec_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    def load(self, tensors: dict) -> None:
        """tensors: reference state_dict key -> CUDA float32 tensor; folds / repacks afterwards."""
        with torch.cuda.device(self.device):
            st = stream_ptr(self.device)
            keep = []
            for key, t in tensors.items():
                t = t.detach()
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != self.device:
                    t = t.to(device=self.device, dtype=torch.float32).contiguous()
                keep.append(t)
                check(lib.ecb_codec_load_tensor(self._h, key.encode(), ptr(t), t.numel(), st))
            check(lib.ecb_codec_finalize(self._h, st))
            torch.cuda.current_stream(self.device).synchronize()  # the staging copies in `keep` may now die

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws


_WORKSPACES: dict = {}


def shared_workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One growing scratch buffer per device, shared by every op (all work is ordered on the current stream,
    like the single-stream reference; callers using several streams must serialise their calls)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        ws = None
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def release_workspaces() -> None:
    _WORKSPACES.clear()


def f16_saturation_count(reset: bool = False) -> int:
    """Operand tiles in which the fp16 pair conversion of the fp32-accurate convs clipped a value (|a| > 65504) since the last
    reset; 0 unless a model's activations leave the fp16 range (then run with ECB_F16_PAIR=0)."""
    return int(lib.ecb_f16_saturation_count(1 if reset else 0))


def profile_begin() -> None:
    lib.ecb_profile_begin()


def profile_end() -> dict:
    """name -> dict(launches, ms, flops, bytes), summed over the launches since profile_begin()."""
    arr = (EcbProfEntry * 16)()
    n = lib.ecb_profile_end(arr, 16)
    return {arr[i].name.decode(): dict(launches=int(arr[i].launches), ms=float(arr[i].ms), flops=float(arr[i].flops),
                                       bytes=float(arr[i].bytes)) for i in range(n)}
