#!/usr/bin/env python
"""Times the SLSTM (two layers: projections + recurrences) alone for several batch sizes, persistent FFMA kernel vs the
step-wise tensor-core form, back to back on one GPU. Diagnostic; bench.py is the contract benchmark."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
import encodec_b200 as eb  # noqa: E402
from encodec_b200 import _native as nat, synth  # noqa: E402

spec = synth.spec_24khz()
sd = synth.make_state_dict(spec, seed=0)
m = eb.EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=True, model_norm="weight_norm",
                               audio_normalize=False, segment=None, name="unset", ratios=spec.ratios, bins=spec.bins,
                               dimension=spec.dimension, share_codebook=False)
m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
m = m.cuda().eval()
codec = m.encoder.native()
dev = torch.device("cuda")
st = nat.stream_ptr(dev)
T, H = int(os.environ.get("LSTM_T", "150")), 512
print(f"{'B':>5s} {'mode':>10s} {'ms':>8s} {'us/step':>8s}")
for B in [int(v) for v in os.environ.get("LSTM_B", "64,128,256,512,960").split(",")]:
    x = torch.randn(B, T, H, device=dev)
    out = torch.empty_like(x)
    ws = torch.empty(nat.lib.ecb_debug_lstm_workspace_bytes(codec.handle, B, T), dtype=torch.uint8, device=dev)
    ref = None
    for mode, lo, tcm, form in (("0", "1", "0", "0"), ("0", "1", "2", "0"), ("1", "1", "0", "0"), ("0", "1", "2", "16"), ("0", "1", "2", "2")):
        if form == "2" and B > 128:
            continue
        os.environ["ECB_LSTM_STEPWISE"] = mode
        os.environ["ECB_LSTM_LO_TMA"] = lo
        os.environ["ECB_LSTM_TC"] = tcm
        os.environ["ECB_LSTM_FORM"] = form

        def run():
            nat.check(nat.lib.ecb_debug_lstm(codec.handle, x.data_ptr(), out.data_ptr(), B, T, ws.data_ptr(), ws.numel(), st))
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if ref is None:
            ref = out.clone()
            err = 0.0
        else:
            err = float((out - ref).abs().max())
        print(f"{B:5d} {'stepwise' if mode == '1' else (('tensor' if form == '0' else 'tensor16' if form == '16' else 'wavefront') if tcm == '2' else 'ffma'):>11s} {ms:8.3f} {ms * 1e3 / (2 * T):8.2f}   max |diff| vs persistent {err:.2e}",
              flush=True)
