#!/usr/bin/env python
"""The entropy-coded path between cudaProfilerStart / cudaProfilerStop, for ncu (see tools/ncu_steps_quick.sh for the metrics):
one batched (compression) pass over a 32 x 750 frame, then the decoding loop of a 32 x 12 frame launched from the host
(ECB_LM_GRAPH=0: under ncu every launch is serialised anyway).

    ncu --profile-from-start off --metrics <M> --clock-control none --csv --log-file gpurun_out/r02_step_ecdc_lm.csv python tools/lm_step_ncu.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["ECB_LM_GRAPH"] = "0"
from encodec_b200 import synth  # noqa: E402
from encodec_b200.lm import LMModel  # noqa: E402

K, T, TD = 32, 750, 12
spec = synth.LMSpec(n_q=32, card=1024, past_context=262)
lm = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads, past_context=spec.past_context)
lm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_lm_state_dict(spec, 3).items()})
lm = lm.cuda().eval()
u = synth.hash_uniform(4, "tp-codes", K * T).reshape(1, K, T)
codes = torch.from_numpy(np.minimum((u * spec.card).astype(np.int64), spec.card - 1)).cuda()
short = codes[:, :, :TD].contiguous()
data = lm.encode_frames(short)[0]
buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
lm.coder_ranges(codes)
lm.decode_frame(buf, 0, K, TD)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
lm.coder_ranges(codes)
got, _ = lm.decode_frame(buf, 0, K, TD)
torch.cuda.synchronize()
rt.cudaProfilerStop()
assert torch.equal(got, short[0])
print("ok")
