#!/usr/bin/env python
"""One step of a bench workload between cudaProfilerStart / cudaProfilerStop, for ncu:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file profiles/r02_step_dram_cfg2.csv python tools/profile_step.py cfg2
    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:rvq_tc -c 1 -o ... python tools/profile_step.py cfg2

Warm-up forwards run unprofiled first (function attributes, lazy module loading, CUDA-graph capture)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    wl["batch"] = int(sys.argv[2])
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rt = torch.cuda.cudart()
if wl["model"] == "rvq":
    import encodec_b200 as eb
    from encodec_b200 import synth
    cbs = synth.hash_normal(4, "cfg4-codebooks", (wl["n_q"], wl["bins"], wl["dim"]))
    q = eb.ResidualVectorQuantizer(dimension=wl["dim"], n_q=wl["n_q"], bins=wl["bins"], codebook_dim=wl["dim"], share_codebook=False)
    for i, layer in enumerate(q.vq.layers):
        layer._codebook.embed.copy_(torch.from_numpy(cbs[i]))
        layer._codebook.inited.fill_(1)
    q = q.to(dev).eval()
    x = torch.randn(256, wl["dim"], 1000, device=dev)
    with torch.no_grad():
        q.encode(x, 75, None)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        codes = q.encode(x, 75, None)
        q.decode(codes)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
else:
    spec = bench.make_spec(wl["model"])
    model, _ = bench.build_model(spec, wl, dev)
    length = int(wl["seconds"] * spec.sample_rate)
    x = (0.3 * torch.randn(wl["batch"], spec.channels, length, device=dev)).clamp_(-1, 1)
    with torch.no_grad():
        for _ in range(2):
            model(x)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        audio, codes, _, _ = model(x)
        if os.environ.get("PROFILE_ECDC"):
            from encodec_b200 import compress
            blob = compress.compress(model, x[0])
            compress.decompress(blob, model)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
print("profiled one step of", name, "batch", wl.get("batch"))
