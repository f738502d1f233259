#!/usr/bin/env python
"""Times the tensor-core conv kernel on the layer shapes of BASELINE config 2 (24 kHz, 64 x 10 s) through the
library's own per-launch CUDA-event profiler. Diagnostic only; bench.py is the contract benchmark."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from encodec_b200 import _native as nat  # noqa: E402

HALO = 16
# name, T_in, C0, taps, stride, pad_left, N, C1, zero_pad, dual
LAYERS = [
    ("res32.b1", 240000, 32, 3, 1, 2, 32, 0, 0, 0),
    ("res32.b3+sc", 240000, 32, 1, 1, 0, 32, 32, 0, 0),
    ("down32", 240000, 32, 4, 2, 2, 64, 0, 0, 1),
    ("res64.b1", 120000, 64, 3, 1, 2, 32, 0, 0, 0),
    ("res64.b3+sc", 120000, 32, 1, 1, 0, 64, 64, 0, 0),
    ("down64", 120000, 64, 8, 4, 4, 128, 0, 0, 1),
    ("res128.b1", 30000, 128, 3, 1, 2, 64, 0, 0, 0),
    ("res128.b3+sc", 30000, 64, 1, 1, 0, 128, 128, 0, 0),
    ("down128", 30000, 128, 10, 5, 5, 256, 0, 0, 1),
    ("res256.b1", 6000, 256, 3, 1, 2, 128, 0, 0, 0),
    ("res256.b3+sc", 6000, 128, 1, 1, 0, 256, 256, 0, 0),
    ("down256", 6000, 256, 16, 8, 8, 512, 0, 0, 0),
    ("lstm.proj", 750, 512, 1, 1, 0, 2048, 0, 1, 0),
    ("enc_out.k7", 750, 512, 7, 1, 6, 128, 0, 0, 0),
    ("up512", 750, 512, 2, 1, 1, 8 * 256, 0, 1, 1),
    ("up256", 6000, 256, 2, 1, 1, 5 * 128, 0, 1, 1),
    ("up128", 30000, 128, 2, 1, 1, 4 * 64, 0, 1, 1),
    ("up64", 120000, 64, 2, 1, 1, 2 * 32, 0, 1, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda")
    st = nat.stream_ptr(dev)
    print(f"{'layer':14s} {'split':>5s} {'ms':>8s} {'TFLOP/s':>8s} {'GB/s':>8s}")
    for (name, T, C0, taps, s, pl, N, C1, zp, dual) in LAYERS:
        if args.only and args.only not in name:
            continue
        items = args.items
        M = T if zp else -(-T // s)
        rows_in = T if zp else T + 2 * HALO
        x = torch.randn(items, rows_in, C0, device=dev)
        x1 = torch.randn(items, M, C1, device=dev) if C1 else None
        ktot = taps * C0 + C1
        w = torch.randn(ktot, N, device=dev) / ktot ** 0.5
        b = torch.randn(N, device=dev)
        o1 = torch.empty(items, M + 2 * HALO, N, device=dev)
        o2 = torch.empty(items, M + 2 * HALO, N, device=dev) if dual else None
        for split in (3, 2, 1):
            def run():
                nat.check(nat.lib.ecb_debug_tc_conv(
                    x.data_ptr(), rows_in * C0, C0, 0 if zp else -HALO, rows_in, taps, s, pl,
                    x1.data_ptr() if C1 else None, M * C1, C1, M, w.data_ptr(), b.data_ptr(), N, M, items,
                    o1.data_ptr() + HALO * N * 4, (o2.data_ptr() + HALO * N * 4) if dual else None,
                    (M + 2 * HALO) * N, HALO if M > HALO else 0, 0, split, st))
            run()
            nat.profile_begin()
            for _ in range(args.reps):
                run()
            res = nat.profile_end()
            prof = res.get("tc_conv_narrow") or res["tc_conv_wide"]
            ms = prof["ms"] / args.reps
            print(f"{name:14s} {split:5d} {ms:8.3f} {prof['flops'] / args.reps / ms / 1e9:8.1f} "
                  f"{prof['bytes'] / args.reps / ms / 1e6:8.0f}", flush=True)
        del x, x1, w, o1, o2
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
