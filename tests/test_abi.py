"""The C-ABI library loads on a CPU-only box and exports every symbol include/encodec_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "encodec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("ecb_encoder_forward", "ecb_decoder_forward", "ecb_rvq_encode_frames", "ecb_rvq_decode_frames",
                 "ecb_codec_load_tensor", "ecb_codec_finalize", "ecb_overlap_add", "ecb_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from encodec_b200 import _native
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_native.SIGNATURES) == declared_symbols()


def test_error_convention_without_gpu():
    """Bad arguments return non-zero + a message; nothing aborts (no compute is attempted)."""
    from encodec_b200 import _native as nat
    spec = nat.make_spec(3, True, False, 32, 128, [8, 5, 4, 2], 7, 7, 3, 2, 2, 1024, 8)  # 3 channels: unsupported
    h = ctypes.c_void_p()
    assert nat.lib.ecb_codec_create(ctypes.byref(spec), ctypes.byref(h)) != 0
    assert "channels" in nat.last_error()
    spec = nat.make_spec(2, True, True, 32, 128, [8, 5, 4, 2], 7, 7, 3, 2, 2, 1024, 8)  # GroupNorm + causal
    assert nat.lib.ecb_codec_create(ctypes.byref(spec), ctypes.byref(h)) != 0
    assert "causal" in nat.last_error()
    spec = nat.make_spec(1, True, False, 32, 128, [8, 5, 4, 2], 7, 7, 3, 2, 2, 1024, 8)
    assert nat.lib.ecb_codec_create(ctypes.byref(spec), ctypes.byref(h)) == 0
    assert nat.lib.ecb_encoder_workspace_bytes(h, 2, 24000) > 2 * 24000 * 32 * 4 * 3
    assert nat.lib.ecb_codec_finalize(h, None) != 0 and "no tensors" in nat.last_error()
    nat.lib.ecb_codec_destroy(h)
