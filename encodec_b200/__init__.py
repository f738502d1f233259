"""B200-native EnCodec codec forward pass -- a drop-in for the reference's ``encodec.model`` path.

Public surface (same names / signatures as ellen660/encodec): ``EncodecModel``, ``SEANetEncoder``,
``SEANetDecoder``, ``ResidualVectorQuantizer``, ``QuantizedResult``, ``LMModel`` (the entropy-coded ``.ecdc`` stream lives in
``encodec_b200.compress``). Everything numerical runs in
``lib/libencodec_b200.so`` (hand-written sm_100a CUDA behind the C ABI of ``include/encodec_b200.h``).
Importing the model classes without that library raises: there is no CPU / PyTorch fallback.
``encodec_b200.synth`` (pure numpy, no CUDA) can be imported on its own.
"""
__version__ = "0.1.0"

_LAZY = {
    "EncodecModel": ("model", "EncodecModel"),
    "SEANetEncoder": ("modules", "SEANetEncoder"),
    "SEANetDecoder": ("modules", "SEANetDecoder"),
    "ResidualVectorQuantizer": ("quantization", "ResidualVectorQuantizer"),
    "QuantizedResult": ("quantization", "QuantizedResult"),
    "HostPipeline": ("pipeline", "HostPipeline"),
    "LMModel": ("lm", "LMModel"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)


__all__ = list(_LAZY)
