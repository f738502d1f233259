#!/usr/bin/env python
"""Prints the per-role clock stamps the persistent tensor-core LSTM kernel records for CTA 0 (steps 20..27): where a step's
time goes (poll, TMA, MMA, cell epilogue, publish). Diagnostic."""
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
B, T, H = int(os.environ.get("LSTM_B", "64")), 64, 512
os.environ["ECB_LSTM_TC"] = "2"
os.environ["ECB_LSTM_STEPWISE"] = "0"
x = torch.randn(B, T, H, device=dev)
out = torch.empty_like(x)
ws = torch.empty(nat.lib.ecb_debug_lstm_workspace_bytes(codec.handle, B, T), dtype=torch.uint8, device=dev)
trace = torch.zeros(3 * 8 * 16, dtype=torch.int64, device=dev)
for i in range(3):
    if i == 2:
        nat.lib.ecb_debug_lstm_trace(trace.data_ptr())
    nat.check(nat.lib.ecb_debug_lstm(codec.handle, x.data_ptr(), out.data_ptr(), B, T, ws.data_ptr(), ws.numel(), st))
    torch.cuda.synchronize()
nat.lib.ecb_debug_lstm_trace(None)
tr = trace.cpu().numpy().reshape(3, 8, 16)
t0 = tr[0, 0, 0]
names = {0: ["poll"] + [f"tma{j}" for j in range(8)] + ["polled", "fenced"],
         1: [f"full{j}" for j in range(8)] + ["commit"] + [f"issued{j}" for j in range(7)],
         2: ["wait", "accf", "ld", "h_stored", "rec", "math", "published"]}
for step in range(8):
    print(f"--- step {20 + step} (second layer of the SLSTM; cycles since the loader's first stamp)")
    for role, rn in ((0, "loader"), (1, "mma"), (2, "cell")):
        print(f"  {rn:7s}", "  ".join(f"{n}={int(tr[role, step, i] - t0)}" for i, n in enumerate(names[role]) if n))
