"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

import encodec_b200 as eb
from encodec_b200 import _native as nat


def build_model(spec, sd, bandwidth, distinct=True, device="cuda"):
    m = eb.EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                   model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment,
                                   name="unset", ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension,
                                   share_codebook=not distinct)
    missing = m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    m = m.to(device).eval()
    if bandwidth is not None:
        m.set_target_bandwidth(bandwidth)
    return m


def tap_stage(fn, stage, numel, device="cuda"):
    """Run fn() with the diagnostic tap armed; returns the tapped activation (flat float32 numpy)."""
    buf = torch.zeros(numel, dtype=torch.float32, device=device)
    nat.lib.ecb_debug_tap(buf.data_ptr(), numel, stage)
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        nat.lib.ecb_debug_tap(None, 0, -1)
    return buf.cpu().numpy()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
