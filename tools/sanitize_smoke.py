#!/usr/bin/env python
"""Small forwards of every model family in one process (diagnostic: a quick fault check after kernel changes; results are
checked by the parity tests). compute-sanitizer is closed on this pool, so this only catches faults the driver reports."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import encodec_b200 as eb  # noqa: E402
from encodec_b200 import synth  # noqa: E402

for name, spec, batch, length, bw in (("24k", synth.spec_24khz(), 2, 12345, 6.0),
                                      ("48k", synth.spec_48khz(), 1, 47520 + 3000, 6.0),
                                      ("fork10hz", synth.spec_fork10hz(), 2, 300 * 34 + 7, 0.08),
                                      ("24k-short", synth.spec_24khz(), 1, 700, 6.0)):
    sd = synth.make_state_dict(spec, seed=5)
    m = eb.EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                   model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment, name="unset",
                                   ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension, share_codebook=False)
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    m = m.cuda().eval()
    m.set_target_bandwidth(bw)
    x = torch.from_numpy(synth.make_audio(3, batch, spec.channels, length)).cuda()
    with torch.no_grad():
        audio, codes, _, _ = m(x)
    torch.cuda.synchronize()
    print(name, tuple(audio.shape), tuple(codes.shape), float(audio.abs().max()), flush=True)
print("done")
