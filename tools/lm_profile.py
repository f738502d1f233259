"""Per-kernel-class device time of the entropy-coded path (library profiler: CUDA events around every launch).

    python tools/lm_profile.py [K] [T]

Prints, for n_q = K codebooks and T steps: the batched (compression) pass and the decoding loop (launched from the host
while profiling; the production loop replays one captured step), microseconds per step and class."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from encodec_b200 import _native as nat, synth  # noqa: E402
from encodec_b200.lm import LMModel  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 750
    spec = synth.LMSpec(n_q=32, card=1024, past_context=262)
    lm = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads,
                 past_context=spec.past_context)
    lm.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_lm_state_dict(spec, 3).items()})
    lm = lm.cuda().eval()
    u = synth.hash_uniform(4, "tp-codes", K * T).reshape(1, K, T)
    codes = torch.from_numpy(np.minimum((u * spec.card).astype(np.int64), spec.card - 1)).cuda()
    data = lm.encode_frames(codes)[0]
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    lm.decode_frame(buf, 0, K, T)
    torch.cuda.synchronize()
    for what in ("batched pass", "decoding loop"):
        nat.profile_begin()
        t0 = time.time()
        if what == "batched pass":
            lm.coder_ranges(codes)
        else:
            got, _ = lm.decode_frame(buf, 0, K, T)
            assert torch.equal(got, codes[0])
        torch.cuda.synchronize()
        dt = time.time() - t0
        prof = nat.profile_end()
        n = 1 if what == "batched pass" else T
        total = sum(v["ms"] for v in prof.values())
        print(f"{what}: K = {K}, T = {T}: {1e3 * dt:.2f} ms wall while profiling, kernels {total:.3f} ms"
              + ("" if n == 1 else f" = {1e3 * total / n:.1f} us per step"))
        for name, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"  {name:12s} {v['launches']:6d} launches  {v['ms']:9.3f} ms  {1e3 * v['ms'] / v['launches']:8.2f} us per launch"
                  + ("" if n == 1 else f"  {1e3 * v['ms'] / n:7.1f} us per step"))
    t0 = time.time()
    lm.decode_frame(buf, 0, K, T)
    dt = time.time() - t0
    print(f"production decoding loop (captured step): {1e3 * dt:.1f} ms = {1e6 * dt / T:.0f} us per step")


if __name__ == "__main__":
    main()
