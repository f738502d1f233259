#!/usr/bin/env python
"""Headline benchmark: audio-seconds encoded+decoded per wall-second (EncodecModel.forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4|cfg5|fork10hz|ecdc_lm]

N > 1 is launched by torchrun (one rank per GPU); clips are sharded across ranks with no data-path collective, and the
only exchange is an NCCL gather of codes and audio to rank 0 (north_star), inside the timed region, issued on a side
stream so that it overlaps the next step's kernels. Rank 0 prints ONE JSON line (see the task contract): `value` is
measured with inputs resident in HBM, `e2e` through the public API from pinned host buffers with H2D/D2H copies inside
the timed region, `roofline` describes the dominant kernel (timed with CUDA events on its stream by the library's own
profiler hooks; bytes = SURVEY.md section 8d's fused-block bytes), `cpu_baseline` is the UNMODIFIED reference
(baseline/_ref, see baseline/reference_arm.py) timed on the host cores.

Workloads (BASELINE.json configs): cfg2 (default, the headline: 24 kHz, 24 kbps, 64 x 10 s per GPU, weak scaling), cfg1
(latency of 1 x 1 s), cfg3 (48 kHz stereo, 32 x 30 s per GPU), cfg4 (RVQ only, 1 M frames x 32 layers, frames/s and
near-tie counts against the reference quantiser), cfg5 (8 192 x 10 s clips IN TOTAL, sharded over the GPUs: strong
scaling). The default line also carries short cfg5 / cfg3 / cfg4 / cfg1 runs of the same launch under `also`.

`--impl reference` times the reference's own CPU forward (`kind: "reference"`) on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import csv
import glob
import json
import math
import os
import re
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "cfg2": dict(model="24k", bandwidth=24.0, batch=64, seconds=10.0, scaling="weak",
                 desc="EnCodec 24 kHz causal mono, 24 kbps (n_q=32), batch 64 x 10 s per GPU, random-init weights"),
    "cfg1": dict(model="24k", bandwidth=6.0, batch=1, seconds=1.0, scaling="weak",
                 desc="EnCodec 24 kHz causal mono, 6 kbps (n_q=8), 1 x 1 s"),
    "cfg3": dict(model="48k", bandwidth=24.0, batch=32, seconds=30.0, scaling="weak",
                 desc="EnCodec 48 kHz stereo, 24 kbps (n_q=16), 1 s segments 1% overlap, batch 32 x 30 s per GPU"),
    "cfg4": dict(model="rvq", frames=1_000_000, n_q=32, bins=1024, dim=128, scaling="weak",
                 desc="RVQ only: 1 000 000 frames x 128 dims, n_q=32 distinct 1024-entry codebooks per GPU, codes checked "
                      "against the reference quantiser on a 65 536-frame sample"),
    "cfg5": dict(model="24k", bandwidth=6.0, batch=512, seconds=10.0, total_clips=8192, scaling="strong",
                 desc="EnCodec 24 kHz causal mono, 6 kbps (n_q=8), 8192 x 10 s clips in total sharded by clip across the GPUs, "
                      "micro-batches of 512 clips, codes+audio gathered to rank 0"),
    # one micro-batch of cfg5 per GPU and step (weak scaling), with the per-kernel profile: explains the cfg5 number
    "cfg5shard": dict(model="24k", bandwidth=6.0, batch=512, seconds=10.0, scaling="weak",
                      desc="EnCodec 24 kHz causal mono, 6 kbps (n_q=8), 512 x 10 s clips per GPU and step (one micro-batch of cfg5)"),
    # the fork's own training configuration (params/091224_l1.yaml): 32 x 4 hours of a 10 Hz signal
    "fork10hz": dict(model="fork10hz", bandwidth=0.08, batch=32, seconds=14400.0, scaling="weak",
                     desc="the fork's 10 Hz model (layer_norm, ratios 6,5,5,2,1, dimension 256, 1024-wide LSTM), 0.08 kbps (n_q=8), "
                          "batch 32 x 144000 samples (4 h each)"),
    # the entropy-coded body of a .ecdc stream (LM + arithmetic coder), codes -> bytes -> codes
    "ecdc_lm": dict(model="lm", n_q=32, frames=750, cpu_frames=300, scaling="weak",
                    desc="entropy-coded .ecdc body of one 10 s clip of the 24 kHz model at 24 kbps per GPU: LMModel (5 layers, dim 200, "
                         "past_context 262) + arithmetic coder over 32 codebooks x 750 latent frames, codes -> bytes -> codes, "
                         "random-init LM weights"),
}
METRIC = "audio-sec/sec encode+decode"
LM_METRIC = "latent frames/sec entropy-coded and decoded (LM + arithmetic coder, n_q=32)"


def make_spec(kind):
    from encodec_b200 import synth
    return {"24k": synth.spec_24khz, "48k": synth.spec_48khz, "fork10hz": synth.spec_fork10hz}[kind]()


def cpu_sample(wl):
    """Bounded sample of the workload for the CPU legs: (clips, seconds per clip)."""
    if wl["model"] == "24k":
        return min(8, wl["batch"]), min(wl["seconds"], 10.0)
    if wl["model"] == "fork10hz":
        return min(2, wl["batch"]), wl["seconds"]          # 2 x 144000 samples: a few seconds of CPU work
    return min(2, wl["batch"]), min(wl["seconds"], 10.0)


def config_of(name, wl, world):
    """The `config` object: identical in the `ours` and `reference` arms of the same launch."""
    if wl["model"] == "rvq":
        return {"workload": wl["desc"], "frames_per_gpu": wl["frames"], "n_q": wl["n_q"], "parallelism": f"dp{world} (frames sharded)",
                "reference_sample": "65536 frames x 32 layers per step on the CPU arm"}
    if wl["model"] == "lm":
        return {"workload": wl["desc"], "name": name, "frames_per_gpu": wl["frames"], "n_q": wl["n_q"],
                "parallelism": f"dp{world} (one independent stream per GPU, no collective)",
                "reference_sample": f"a frame of {wl['cpu_frames']} latent frames x {wl['n_q']} codebooks per step on the CPU arm"}
    clips, seconds = cpu_sample(wl)
    total = wl.get("total_clips")
    return {"workload": wl["desc"], "name": name,
            "global_batch": total if total else wl["batch"] * world, "clip_seconds": wl["seconds"],
            "parallelism": f"dp{world} (clips sharded, one NCCL gather of codes+audio to rank 0 per step, on a side stream)",
            "reference_sample": f"{clips} clip(s) x {seconds:g} s of this workload per step on the CPU arm"}


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback (B200_PROFILING.md)")


def measure_tf32_peak(dev, seconds=0.4):
    """Dense TF32 GEMM rate of this GPU (cuBLAS through torch, 8192^3, run for ~`seconds`): context for the 3xTF32
    kernels -- MEASURED_PEAKS.json holds the bf16 figure only ('TF32 peak ... measure it')."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 0
        e0.record()
        t0 = time.perf_counter()
        while True:
            for _ in range(5):
                torch.matmul(a, b, out=c)
            reps += 5
            if time.perf_counter() - t0 > seconds and reps >= 20:
                break
        e1.record()
        torch.cuda.synchronize(dev)
        return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as ex:  # NVML missing: report that instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(ex).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------
# CPU legs: the unmodified reference when baseline/_ref travelled with the snapshot, else the ATen port
# ---------------------------------------------------------------------------------------------------------
_PORT_PARAMS = {}


def cpu_throughput(spec, sd, bandwidth, clips, seconds):
    """(audio-s/s, seconds spent, kind, how) of the reference's forward on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm as ra
    cores = os.cpu_count() or 1
    if ra.available() or ra.install():
        v, spent = ra.throughput(spec, sd, bandwidth, clips, seconds)
        return v, spent, "reference", (f"the unmodified reference from baseline/_ref (EncodecModel._get_model(...).eval(), model(x), "
                                       f"torch.no_grad, {cores} threads)")
    import torch
    from encodec_b200 import synth
    from oracle import torch_port as port
    torch.set_num_threads(cores)
    x = synth.make_audio(999, clips, spec.channels, int(seconds * spec.sample_rate))
    params = _PORT_PARAMS.setdefault(id(sd), port.TorchParams(sd, spec.norm))
    t0 = time.perf_counter()
    port.forward(x, sd, spec, bandwidth, params)
    dt = time.perf_counter() - t0
    return clips * seconds / dt, dt, "port", (f"oracle/torch_port.py (baseline/_ref is missing: the reference's forward restated on the "
                                              f"ATen CPU kernels it uses, {cores} threads)")


def cpu_rvq_throughput(frames, cbs):
    """(frames/s, seconds, kind, codes) of the reference quantiser (core_vq.ResidualVectorQuantization.encode) on CPU."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm as ra
    t0 = time.perf_counter()
    if ra.available() or ra.install():
        codes = ra.rvq_encode(frames, cbs)
        kind = "reference"
    else:
        from oracle import encodec_oracle as orc
        codes = orc.rvq_forward(np.ascontiguousarray(frames.T)[None], cbs, cbs.shape[0])[1][:, 0]
        kind = "port"
    dt = time.perf_counter() - t0
    return frames.shape[0] / dt, dt, kind, codes


def lm_codes(wl, n_frames, seed):
    from encodec_b200 import synth
    u = synth.hash_uniform(seed, "ecdc-lm-codes", wl["n_q"] * n_frames).reshape(wl["n_q"], n_frames)
    return np.minimum((u ** 2 * 1024).astype(np.int64), 1023)     # skewed symbols


def cpu_lm_throughput(wl, n_frames):
    """(frames/s, seconds, kind, how): the reference's LM + arithmetic coder loops (compress.py:66-87,125-152) on one frame."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import reference_arm as ra
    from encodec_b200 import synth
    spec = synth.LMSpec(n_q=wl["n_q"], card=1024, past_context=262)
    sd = _LM_SD.setdefault("sd", synth.make_lm_state_dict(spec, 3))
    codes = lm_codes(wl, n_frames, 5)
    cores = os.cpu_count() or 1
    if ra.available() or ra.install():
        tc, td, nbytes, out = ra.lm_entropy_round_trip(spec, sd, codes)
        assert np.array_equal(out, codes)
        return n_frames / (tc + td), tc + td, "reference", (f"the unmodified reference from baseline/_ref: LMModel step by step, "
                                                            f"build_stable_quantized_cdf, ArithmeticCoder.push / ArithmeticDecoder.pull, "
                                                            f"{cores} threads; coding {tc:.2f} s + decoding {td:.2f} s")
    from oracle import lm_oracle as lo
    t0 = time.perf_counter()
    p = lo.lm_probas(sd, codes, num_layers=spec.num_layers, num_heads=spec.num_heads, past_context=spec.past_context, dtype=np.float32)
    cdfs = lo.build_stable_quantized_cdf(p)
    data = lo.encode_frame(codes, cdfs)
    dec = lo.ArithmeticDecoder(data)
    for t in range(n_frames):
        for k in range(wl["n_q"]):
            assert dec.pull(cdfs[t, k]) == codes[k, t]
    dt = time.perf_counter() - t0
    return n_frames / dt, dt, "port", "oracle/lm_oracle.py (baseline/_ref is missing)"


_LM_SD = {}


def run_reference_arm(args, name, wl, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from encodec_b200 import synth
    cores = os.cpu_count() or 1
    if wl["model"] == "lm":
        for _ in range(max(min(args.warmup, 1), 0)):
            cpu_lm_throughput(wl, 2)
        spent = 0.0
        for _ in range(args.steps):
            _, sec, kind, how = cpu_lm_throughput(wl, wl["cpu_frames"])
            spent += sec                    # the coding + decoding loops only (model construction is outside, as in our arm)
        dt = spent / args.steps
        value, unit, metric = wl["cpu_frames"] / dt, "frames/s", LM_METRIC
        sample = f"a frame of {wl['cpu_frames']} latent frames x {wl['n_q']} codebooks per step"
    elif wl["model"] == "rvq":
        frames = synth.hash_normal(4, "cfg4-frames", (65536, wl["dim"]))
        cbs = synth.hash_normal(4, "cfg4-codebooks", (wl["n_q"], wl["bins"], wl["dim"]))
        for _ in range(max(args.warmup, 0)):
            cpu_rvq_throughput(frames[:4096], cbs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, kind, _ = cpu_rvq_throughput(frames, cbs)
        dt = (time.perf_counter() - t0) / args.steps
        value, unit, metric = frames.shape[0] / dt, "frames/s", "RVQ frames/sec (n_q=32)"
        sample = "65536 frames x 32 layers per step"
        how = "core_vq.ResidualVectorQuantization.encode" if kind == "reference" else "oracle numpy restatement"
    else:
        spec = make_spec(wl["model"])
        sd = synth.make_state_dict(spec, seed=0)
        clips, seconds = cpu_sample(wl)
        for _ in range(max(args.warmup, 0)):
            cpu_throughput(spec, sd, wl["bandwidth"], 1, min(seconds, 1.0) if wl["model"] != "fork10hz" else 1000.0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, _, kind, how = cpu_throughput(spec, sd, wl["bandwidth"], clips, seconds)
        dt = (time.perf_counter() - t0) / args.steps
        value, unit, metric = clips * seconds / dt, "audio-s/s", METRIC
        sample = f"{clips} clip(s) x {seconds:g} s of the workload per step"
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(name, wl, world),
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample, "how": how},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# algorithmic work per item ("one fused kernel per SConv / ResBlock / LSTM, input read once,
# output written once"), attributed to the kernel class that runs each block in this build
# ---------------------------------------------------------------------------------------------------------
def algorithmic_elems(spec, seg_len, fused_res32=True):
    """fp32 elements moved per item (one clip, or one segment of the 48 kHz model) by kernel class."""
    nf, C, dim = 32, spec.channels, spec.dimension
    cls = {}

    def add(name, v):
        cls[name] = cls.get(name, 0.0) + v

    def conv_class(c_in_row, n_out):   # tc_conv's own narrow / wide split (tc_conv.cu: launch_tc_conv)
        return "tc_conv_narrow" if c_in_row <= 64 and n_out <= 64 else "tc_conv_wide"

    def res_class(ch):
        if ch == 32 and fused_res32 and spec.norm == "weight_norm":
            return "tc_res"
        return "tc_conv_narrow" if ch <= 64 else "tc_conv_wide"

    ratios_enc = list(reversed(spec.ratios))
    t, ch = seg_len, nf
    add("conv_in", t * C + t * nf)
    ts = [t]
    for r in ratios_enc:
        add(res_class(ch), 2.0 * t * ch)
        t2 = -(-t // r)
        add(conv_class(ch, 2 * ch), t * ch + t2 * 2 * ch)
        t, ch = t2, 2 * ch
        ts.append(t)
    add("lstm_recurrent", 2.0 * t * ch)                 # the SLSTM block: x in, y out (its projections are inside the block)
    add("tc_conv_wide", t * ch + t * dim)               # final conv
    add("rvq_encode", 2.0 * t * dim)                    # frames in, quantized out (+ codes, negligible)
    add("tc_conv_wide", t * dim + t * ch)               # decoder.model.0
    add("lstm_recurrent", 2.0 * t * ch)
    for r in spec.ratios:
        t2 = t * r
        add(conv_class(ch, r * ch // 2), t * ch + t2 * ch // 2)   # transposed conv as a 2-tap GEMM with N = r * C_out
        t, ch = t2, ch // 2
        add(res_class(ch), 2.0 * t * ch)
    add("conv_out", t * nf + t * C)
    return cls


def parse_step_dram(path_glob):
    """DRAM bytes per kernel name from a committed `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,
    gpu__time_duration.sum --csv` capture of ONE step of the default workload (profiles/r02_step_dram_*.csv)."""
    files = sorted(glob.glob(path_glob))
    if not files:
        return None, None
    per = {}
    with open(files[-1], newline="") as f:
        rows = [r for r in csv.reader(f) if len(r) > 10]
    if not rows:
        return None, files[-1]
    head = rows[0]
    try:
        i_name, i_metric, i_unit, i_val = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Unit"), head.index("Metric Value")
    except ValueError:
        return None, files[-1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[1:]:
        if not r[i_metric].startswith("dram__bytes"):
            continue
        m = re.search(r"([A-Za-z_]\w*_kernel)", r[i_name])   # this library's kernels all end in _kernel
        name = m.group(1) if m else "other"
        per[name] = per.get(name, 0.0) + float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    return per, files[-1]


KERNEL_OF_CLASS = {"tc_conv_narrow": "tc_conv_kernel", "tc_conv_wide": "tc_conv_kernel", "tc_res": "tc_res_kernel",
                   "lstm_recurrent": "lstm_tc_kernel", "rvq_encode": "rvq_tc_kernel", "conv_in": "conv_in_kernel",
                   "conv_out": "conv_out_kernel"}
TENSOR_CLASSES = ("tc_conv_wide", "conv_gemm", "rvq_encode", "lstm_recurrent")


def build_model(spec, wl, dev):
    import torch
    import encodec_b200 as eb
    from encodec_b200 import synth
    sd = synth.make_state_dict(spec, seed=0)
    model = eb.EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                       model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment,
                                       name="unset", ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension,
                                       share_codebook=False)
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    model = model.to(dev).eval()
    model.set_target_bandwidth(wl["bandwidth"])
    return model, sd


class Timer:
    """CUDA-event timing of a region on the current stream, barrier + synchronize on both sides, max over ranks."""

    def __init__(self, dev, world):
        import torch
        self.torch, self.dev, self.world = torch, dev, world
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def start(self):
        self.sync()
        self.e0.record()

    def stop(self):
        self.e1.record()
        self.sync()
        ms = self.e0.elapsed_time(self.e1)
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([ms], device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def run_codec(name, wl, dev, rank, world, steps, warmup, full):
    """One codec workload with weak scaling (every rank its own batch). `full` adds the e2e leg, the per-kernel profile and the
    decoder-precision variant; the short form (sub-results of the default line) reports value / ms only."""
    import torch
    from encodec_b200 import _native as nat, dist as ebdist
    spec = make_spec(wl["model"])
    model, sd = build_model(spec, wl, dev)
    batch = wl["batch"]
    length = int(wl["seconds"] * spec.sample_rate)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    n_rot = 3  # rotate inputs so that no step re-reads the previous step's input from L2
    xs = [(0.3 * torch.randn(batch, spec.channels, length, generator=gen, device=dev)).clamp_(-1, 1) for _ in range(n_rot)]
    audio_s = batch * wl["seconds"] * world
    queue = ebdist.GatherQueue(dev, 0, spec.bins) if world > 1 else None
    tm = Timer(dev, world)

    def step(i, gather=True):
        audio, codes, _, _ = model(xs[i % n_rot])
        if queue is not None and gather:
            queue.submit(codes, audio)     # side stream: overlaps the next step's kernels
            queue.keep_last(2)             # keep at most two steps of gathered results alive
        return audio, codes

    out = {}
    with torch.no_grad():
        for i in range(max(warmup, 1)):
            step(i)
        if queue is not None:
            queue.finish()
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        launches0 = nat.launch_count()
        tm.start()
        for i in range(steps):
            step(i)
        if queue is not None:
            queue.finish()                 # the last gather is inside the timed region
        ms = tm.stop()
        out["clocks"] = sampler.stop()
        out["gpu_launches"] = nat.launch_count() - launches0
        out["ms_per_step"] = ms / steps
        out["value"] = audio_s / (ms / steps / 1e3)
        out["audio_seconds_per_step"] = audio_s
        if not full:
            return out, model, spec, sd

        # ---- end to end through the public API from pinned host memory ---------------------------------
        from encodec_b200.pipeline import HostPipeline
        x_host = xs[0].cpu().pin_memory()
        x_dev = torch.empty_like(xs[0])
        a0, c0 = step(0, gather=False)
        audio_host = torch.empty(a0.shape, dtype=a0.dtype).pin_memory()
        codes_host = torch.empty(c0.shape, dtype=c0.dtype).pin_memory()
        h2d = x_host.numel() * 4
        d2h = audio_host.numel() * 4 + codes_host.numel() * 8
        # (a) the plain call sequence a user of the reference writes, on one stream: copy in, forward, copy out
        # (one untimed round first: the first DMA to / from freshly pinned pages is not representative)
        x_dev.copy_(x_host, non_blocking=True)
        audio_host.copy_(a0, non_blocking=True)
        codes_host.copy_(c0, non_blocking=True)
        torch.cuda.synchronize()
        tm.start()
        for i in range(steps):
            x_dev.copy_(x_host, non_blocking=True)
            audio, codes, _, _ = model(x_dev)
            audio_host.copy_(audio, non_blocking=True)
            codes_host.copy_(codes, non_blocking=True)
            if queue is not None:
                queue.submit(codes, audio)
                queue.keep_last(2)
        if queue is not None:
            queue.finish()
        ms_serial = tm.stop()
        # (b) the package's host pipeline: the same copies and the same forward per step on three streams
        pipe = HostPipeline(model, depth=2)

        def after(a, c):
            if queue is not None:
                queue.submit(c, a)
                queue.keep_last(2)
        for _ in pipe.run([x_host] * 2, after_forward=after):
            pass
        # three timed runs of `steps` steps each, the MEDIAN is reported (all three are listed): the host side of this path
        # (PCIe DMA, Python threads) is noisier from box to box than the device-timed value above
        runs = []
        for _ in range(3):
            tm.start()
            for _ in pipe.run([x_host] * steps, after_forward=after):
                pass
            pipe.join()
            if queue is not None:
                queue.finish()
            runs.append(tm.stop())
        ms_e2e = sorted(runs)[1]
        out["e2e"] = {"value": audio_s / (ms_e2e / steps / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps,
                      "ms_per_step_runs": [r / steps for r in runs], "reported": "median of 3 runs",
                      "api": "encodec_b200.pipeline.HostPipeline(model, depth=2).run(pinned host batches)",
                      "single_stream_ms_per_step": ms_serial / steps}

        # ---- variant: fp32-accurate decoder (the default decoder runs one TF32 pass) -----------------------------
        out["variants"] = {}
        if rank == 0 and spec.norm == "weight_norm" and wl["model"] == "24k":
            model.decoder.tf32 = False
            step(0, gather=False)
            tm.torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i, gather=False)
            e1.record()
            torch.cuda.synchronize(dev)
            model.decoder.tf32 = None
            v_ms = e0.elapsed_time(e1) / steps
            out["variants"]["decoder_fp32_accurate"] = {
                "value": batch * wl["seconds"] / (v_ms / 1e3), "unit": "audio-s/s (this rank)", "ms_per_step": v_ms,
                "note": "SEANetDecoder.tf32=False: decoder convs with split operands (3xTF32) like the encoder; the default is one "
                        "TF32 pass (audio within 1e-4 max-abs / 2.5e-5 RMS of the reference on the golden cases and on real "
                        "speech at three loudness levels; bar 1e-3 / 1e-4)"}

        # ---- per-kernel-class timing (CUDA events on the launching stream, inside the library) -----------
        if rank == 0:
            torch.cuda.synchronize(dev)
            nat.profile_begin()
            for i in range(steps):
                step(i, gather=False)
            torch.cuda.synchronize(dev)
            out["prof"] = nat.profile_end()
    return out, model, spec, sd


def rooflines_of(prof, steps, spec, wl, n_items, seg_len, peaks):
    """Per kernel class: live ms, section 8d bytes / flops -> achieved GB/s or TFLOP/s against the measured peaks."""
    from encodec_b200 import synth  # noqa: F401
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    elems = algorithmic_elems(spec, seg_len)
    classes = {}
    for name, v in prof.items():
        ms = v["ms"] / steps
        alg_bytes = 4.0 * elems.get(name, 0.0) * n_items
        launch_bytes = v["bytes"] / steps
        flops = v["flops"] / steps
        entry = {"launches_per_step": v["launches"] / steps, "ms_per_step": ms, "share": v["ms"] / total_ms,
                 "algorithmic_gb_per_step": alg_bytes / 1e9, "launch_counted_gb_per_step": launch_bytes / 1e9,
                 "gbs": alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0,
                 "tflops": flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}
        entry["launch_traffic_ratio"] = launch_bytes / alg_bytes if alg_bytes > 0 else None
        entry["hbm_frac"] = entry["gbs"] / peaks["hbm_gbs"]
        entry["tensor_frac"] = entry["tflops"] / peaks["tflops"]
        entry["bound"] = "tensor" if name in TENSOR_CLASSES else "hbm"
        classes[name] = entry
    return classes, total_ms


def run_cfg5(name, wl, dev, rank, world, steps, warmup, total_clips):
    """Strong scaling: `total_clips` clips in total, sharded by clip, micro-batches of wl['batch'] clips per forward, every
    micro-batch followed by a gather of codes+audio to rank 0 on a side stream. value = total audio seconds / time (max over ranks)."""
    import torch
    from encodec_b200 import _native as nat, dist as ebdist
    spec = make_spec(wl["model"])
    model, _ = build_model(spec, wl, dev)
    counts = ebdist.shard_counts(total_clips, world)
    length = int(wl["seconds"] * spec.sample_rate)
    gen = torch.Generator(device=dev).manual_seed(777 + rank)
    x = (0.3 * torch.randn(counts[rank], spec.channels, length, generator=gen, device=dev)).clamp_(-1, 1)
    tm = Timer(dev, world)
    mb = wl["batch"]
    with torch.no_grad():
        for _ in range(max(warmup, 1)):   # warm-up on one micro-batch per rank (same kernels, same shapes)
            ebdist.forward_shard(model, x[:mb], [min(mb, c) for c in counts], dst=0, micro_batch=mb)
        launches0 = nat.launch_count()
        tm.start()
        for _ in range(steps):
            audio, codes = ebdist.forward_shard(model, x, counts, dst=0, micro_batch=mb)
        ms = tm.stop()
    if rank == 0:
        assert codes.shape[0] == total_clips and audio.shape[0] == total_clips
    audio_s = total_clips * wl["seconds"]
    return {"value": audio_s / (ms / steps / 1e3), "unit": "audio-s/s", "ms_per_step": ms / steps, "total_clips": total_clips,
            "clips_per_rank": counts[0], "micro_batch": mb, "scaling": "strong", "gpu_launches": nat.launch_count() - launches0,
            "gathered_on_rank0": {"codes": list(codes.shape), "audio": list(audio.shape)} if rank == 0 else None}


def run_cfg4(wl, dev, rank, world, steps, warmup, check):
    """RVQ only through ResidualVectorQuantizer.encode (the tensor-core quantiser): frames/s; on rank 0 the codes of a
    65 536-frame sample are compared with the reference quantiser (hard mismatches / near-ties per SURVEY.md section 8c)."""
    import torch
    import encodec_b200 as eb
    from encodec_b200 import _native as nat, synth
    n, n_q, bins, dim = wl["frames"], wl["n_q"], wl["bins"], wl["dim"]
    cbs = synth.hash_normal(4, "cfg4-codebooks", (n_q, bins, dim))
    q = eb.ResidualVectorQuantizer(dimension=dim, n_q=n_q, bins=bins, codebook_dim=dim, share_codebook=False)
    for i, layer in enumerate(q.vq.layers):
        layer._codebook.embed.copy_(torch.from_numpy(cbs[i]))
        layer._codebook.inited.fill_(1)
    q = q.to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    t_f = 1000
    xs = [torch.randn(n // t_f, dim, t_f, generator=gen, device=dev) for _ in range(2)]   # [B, D, T]: 1000 x 1000 frames
    tm = Timer(dev, world)
    with torch.no_grad():
        for i in range(max(warmup, 1)):
            q.encode(xs[i % 2], 75, None)
        launches0 = nat.launch_count()
        # every step is timed by itself (barrier + device events, max over ranks) and the MEDIAN step is reported: a step is one
        # 36 ms kernel behind a few host calls, so a single host hiccup between two steps would otherwise show up as throughput
        per_step = []
        for i in range(steps):
            tm.start()
            codes = q.encode(xs[i % 2], 75, None)
            per_step.append(tm.stop())
        ms = sorted(per_step)[len(per_step) // 2] * steps
        launches = nat.launch_count() - launches0
        nat.profile_begin()
        q.encode(xs[0], 75, None)
        torch.cuda.synchronize(dev)
        prof = nat.profile_end()
    assert tuple(codes.shape) == (n_q, n // t_f, t_f)
    frames_s = n * world / (ms / steps / 1e3)
    flops = 2.0 * n * n_q * bins * dim
    res = {"value": frames_s, "unit": "frames/s", "ms_per_step": ms / steps, "frames_per_gpu": n, "n_q": n_q,
           "tflops_algorithmic": flops * world / (ms / steps / 1e3) / 1e12, "gpu_launches": launches,
           "ms_per_step_runs": per_step, "reported": "median step",
           "kernels": {k: {"ms": v["ms"], "launches": v["launches"]} for k, v in prof.items()}}
    if check and rank == 0:
        from oracle import encodec_oracle as orc
        m = 65536
        frames = synth.hash_normal(4, "cfg4-frames", (m, dim))
        xt = torch.from_numpy(np.ascontiguousarray(frames.reshape(64, m // 64, dim).transpose(0, 2, 1))).to(dev)
        with torch.no_grad():
            got = q.encode(xt, 75, None).cpu().numpy().reshape(n_q, m)
        rate, spent, kind, want = cpu_rvq_throughput(frames, cbs)
        score = orc.score_codes(frames, cbs, want, got)
        res["parity_sample"] = {"frames": m, "decisions": int(m * n_q), "score": score, "checker": kind,
                                "note": "teacher-forced per layer on the checker's residuals; near_tie = the two candidates' "
                                        "fp64 distances differ by < 1e-5 relative (north_star), anything else is `hard`"}
        res["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": kind,
                               "sample": f"{m} frames x {n_q} layers, {spent:.1f} s of CPU work"}
    return res


def run_ecdc_lm(wl, dev, rank, world, steps, warmup, cpu_leg):
    """codes [32, 750] -> entropy-coded bytes (batched LM pass + host coder) -> codes (device decoding loop).
    A step is the round trip of one frame; `value` with the codes resident on the device, `e2e` from / to pinned host memory."""
    import torch
    from encodec_b200 import _native as nat, synth
    from encodec_b200.lm import LMModel
    K, T = wl["n_q"], wl["frames"]
    spec = synth.LMSpec(n_q=K, card=1024, past_context=262)
    lm = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads,
                 past_context=spec.past_context)
    lm.load_state_dict({k: torch.from_numpy(v) for k, v in _LM_SD.setdefault("sd", synth.make_lm_state_dict(spec, 3)).items()})
    lm = lm.to(dev).eval()
    host = [torch.from_numpy(lm_codes(wl, T, 10 + 3 * rank + i)).pin_memory() for i in range(3)]
    resident = [h.to(dev) for h in host]
    tm = Timer(dev, world)
    detail = {}

    def round_trip(codes_dev):
        t0 = time.perf_counter()
        data = lm.encode_frames(codes_dev[None])[0]
        t1 = time.perf_counter()
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
        got, end = lm.decode_frame(buf, 0, K, T)
        t2 = time.perf_counter()
        assert end == len(data)
        return got, len(data), t1 - t0, t2 - t1

    for i in range(max(warmup, 3)):
        round_trip(resident[i % 3])
    launches0 = nat.launch_count()
    tm.start()
    tc = td = 0.0
    for i in range(steps):
        got, nbytes, a, b = round_trip(resident[i % 3])
        tc, td = tc + a, td + b
    ms = tm.stop()
    launches = nat.launch_count() - launches0
    assert torch.equal(got, resident[(steps - 1) % 3]), "entropy-coded round trip changed the codes"
    tm.start()
    for i in range(steps):
        got, nbytes, _, _ = round_trip(host[i % 3].to(dev, non_blocking=True))
        back = got.cpu()
    ms_e2e = tm.stop()
    assert torch.equal(back, host[(steps - 1) % 3])
    # per-kernel-class device time of one decoding loop (launched from the host while profiling) and one batched pass
    nat.profile_begin()
    lm.coder_ranges(resident[0][None])
    torch.cuda.synchronize(dev)
    prof_c = nat.profile_end()
    data = lm.encode_frames(resident[0][None])[0]
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(dev)
    nat.profile_begin()
    lm.decode_frame(buf, 0, K, T)
    torch.cuda.synchronize(dev)
    prof_d = nat.profile_end()
    detail = {"compress_ms": 1e3 * tc / steps, "decompress_ms": 1e3 * td / steps, "decode_us_per_latent_frame": 1e6 * td / steps / T,
              "bytes_per_frame": nbytes, "bits_per_symbol": 8.0 * nbytes / (K * T),
              "batched_pass_kernels": {k: {"ms": v["ms"], "launches": v["launches"]} for k, v in prof_c.items()},
              "decoding_loop_kernels_eager": {k: {"us_per_latent_frame": 1e3 * v["ms"] / T, "launches_per_latent_frame": v["launches"] / T}
                                              for k, v in prof_d.items()},
              "note": "the decoding loop replays ONE captured step per latent frame (4 launches: the transformer as one 8-CTA cluster "
                      "kernel -- profiled under lm_linear --, the heads' linear, softmax + cdf, the decoder pull); the per-class times "
                      "above are from a host-launched loop with CUDA events around every launch (event overhead included)"}
    # the chain is latency-bound: the dominant class by time is lm_linear (the cluster kernel + the heads' linear); its
    # algorithmic bytes per latent frame are the weights read once (the activations of one row are negligible)
    w_bytes = 4.0 * (spec.num_layers * (4 * spec.dim * spec.dim + 2 * spec.dim * spec.hidden) + K * spec.card * spec.dim)
    lin_us = detail["decoding_loop_kernels_eager"].get("lm_linear", {}).get("us_per_latent_frame")
    res = {"value": world * T / (ms / steps / 1e3), "ms_per_step": ms / steps, "gpu_launches": launches,
           "e2e": {"value": world * T / (ms_e2e / steps / 1e3), "unit": "frames/s", "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step": K * T * 8 + nbytes, "d2h_bytes_per_step": K * T * 8 + K * T * 8},
           "detail": detail, "weights_bytes_per_latent_frame": w_bytes, "lm_linear_us_per_latent_frame": lin_us}
    if cpu_leg and rank == 0:
        cpu_lm_throughput(wl, 2)
        v, spent, kind, how = cpu_lm_throughput(wl, wl["cpu_frames"])
        res["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": kind,
                               "sample": f"a frame of {wl['cpu_frames']} latent frames x {K} codebooks, {spent:.1f} s of CPU work", "how": how}
    return res


_REAL_STDOUT = None


def init_dist(dev, world):
    """NCCL over the GPUs of this node. NCCL_DEBUG stays whatever the caller set; since NCCL logs to stdout, file descriptor 1
    is pointed at stderr for the rest of the run and the one JSON line is written to the saved real stdout at the end."""
    global _REAL_STDOUT
    if world == 1:
        return
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()


def emit(line):
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU and forward (default: the workload's)")
    ap.add_argument("--total-clips", type=int, default=None, help="cfg5: clips in total over all GPUs (default 8192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="default workload only: skip the cfg5 / cfg3 / cfg4 / cfg1 sub-results")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    name = args.workload
    wl = dict(WORKLOADS[name])
    if args.batch:
        wl["batch"] = args.batch
    if args.total_clips:
        wl["total_clips"] = args.total_clips

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, name, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0 or world == 1:
        entry.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    init_dist(dev, world)
    from encodec_b200 import _native as nat  # noqa: F401
    peaks = read_peaks()
    line = None

    if name == "cfg4":
        res = run_cfg4(wl, dev, rank, world, args.steps, max(args.warmup, 3), check=not args.no_cpu_baseline)
        if rank == 0:
            tf = res["tflops_algorithmic"] / world
            line = {"metric": "RVQ frames/sec (n_q=32)", "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(name, wl, world),
                    "gpu_launches": res["gpu_launches"],
                    "roofline": {"kernel": "rvq_tc_kernel", "bound": "tensor", "achieved": tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                 "frac": tf / peaks["tflops"], "traffic": None,
                                 "peak_source": peaks["source"] + ", dense bf16 sustained; algorithmic FLOPs 2*N*n_q*1024*128, executed as "
                                                "3xTF32 split operands"},
                    "cpu_baseline": res.get("cpu_baseline"), "parity_sample": res.get("parity_sample"), "kernels": res["kernels"],
                    "e2e": None}
    elif name == "ecdc_lm":
        res = run_ecdc_lm(wl, dev, rank, world, args.steps, args.warmup, cpu_leg=not args.no_cpu_baseline)
        if rank == 0:
            lin_us = res["lm_linear_us_per_latent_frame"]
            gbs = res["weights_bytes_per_latent_frame"] / (lin_us * 1e-6) / 1e9 if lin_us else None
            line = {"metric": LM_METRIC, "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                    "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32 (LM), u64 / f64 (coder)", "data": "synthetic", "config": config_of(name, wl, world),
                    "gpu_launches": res["gpu_launches"], "e2e": res["e2e"], "detail": res["detail"],
                    "roofline": {"kernel": "lm_linear_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": gbs / peaks["hbm_gbs"] if gbs else None, "traffic": None, "peak_source": peaks["source"],
                                 "note": "decoding loop, per latent frame: the transformer (one cluster kernel) and the heads' linear read every "
                                         "weight once (35.8 MB, L2-resident) in the measured time; the step is a chain of dependent phases "
                                         "on one row, bound by memory + barrier latency, not by bandwidth or the tensor pipe"},
                    "cpu_baseline": res.get("cpu_baseline")}
    elif name == "cfg5":
        res = run_cfg5(name, wl, dev, rank, world, max(1, min(args.steps, 2)), args.warmup, wl["total_clips"])
        if rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": "audio-s/s", "n_gpus": world, "steps": max(1, min(args.steps, 2)),
                    "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(name, wl, world),
                    "gpu_launches": res["gpu_launches"], "detail": res, "roofline": None, "cpu_baseline": None, "e2e": None}
    else:
        out, model, spec, sd = run_codec(name, wl, dev, rank, world, args.steps, args.warmup, full=True)
        seg_len = spec.segment_length or int(wl["seconds"] * spec.sample_rate)
        n_seg = 1 if spec.segment_length is None else len(model._segments(int(wl["seconds"] * spec.sample_rate))[0])
        also = {}
        if name == "cfg2" and not args.no_extras:
            # the metric's other configurations, measured in the same launch (short runs; every rank takes part)
            del model
            torch.cuda.empty_cache()
            w5 = dict(WORKLOADS["cfg5"])
            r5 = run_cfg5("cfg5", w5, dev, rank, world, 1, 1, w5["total_clips"])
            also["cfg5_24k_6kbps_8192clips_strong"] = dict(r5, config=config_of("cfg5", w5, world))
            torch.cuda.empty_cache()
            w3 = dict(WORKLOADS["cfg3"])
            o3, m3, _, _ = run_codec("cfg3", w3, dev, rank, world, 3, 2, full=False)
            also["cfg3_48k_24kbps_weak"] = {"value": o3["value"], "unit": "audio-s/s", "ms_per_step": o3["ms_per_step"],
                                            "scaling": "weak", "gpu_launches": o3["gpu_launches"], "config": config_of("cfg3", w3, world)}
            del m3
            torch.cuda.empty_cache()
            w4 = dict(WORKLOADS["cfg4"])
            r4 = run_cfg4(w4, dev, rank, world, 3, 3, check=(world == 1 and not args.no_cpu_baseline))
            also["cfg4_rvq_1m_frames"] = dict(r4, config=config_of("cfg4", w4, world))
            w1 = dict(WORKLOADS["cfg1"])
            o1, m1, _, _ = run_codec("cfg1", w1, dev, rank, world, 50, 10, full=False)
            also["cfg1_24k_6kbps_1x1s_latency"] = {"ms_per_call": o1["ms_per_step"], "value": o1["value"], "unit": "audio-s/s",
                                                   "gpu_launches_per_call": o1["gpu_launches"] / 50}
            del m1
            torch.cuda.empty_cache()
        if rank == 0:
            prof = out.get("prof", {})
            n_items = wl["batch"] * n_seg
            # which recurrence kernel a launch of this size runs (codec.cu tc_lstm): two-layer wavefront up to 128 items, 16
            # units per CTA from 384 (both lstm_tcw_kernel), lstm_tc_kernel in between
            KERNEL_OF_CLASS["lstm_recurrent"] = "lstm_tcw_kernel" if (n_items <= 128 or n_items >= 384) else "lstm_tc_kernel"
            classes, total_ms = rooflines_of(prof, args.steps, spec, wl, n_items, seg_len, peaks) if prof else ({}, 1.0)
            dram, dram_file = parse_step_dram(os.path.join(ROOT, "profiles", "r02_step_dram_*.csv"))
            # the dominant KERNEL by time over the step; tc_conv_kernel is one template profiled as two classes
            by_kernel = {}
            for cname, e in classes.items():
                k = KERNEL_OF_CLASS.get(cname, cname)
                d = by_kernel.setdefault(k, {"ms": 0.0, "alg_gb": 0.0, "launch_gb": 0.0, "tflops_ms": 0.0, "classes": []})
                d["ms"] += e["ms_per_step"]
                d["alg_gb"] += e["algorithmic_gb_per_step"]
                d["launch_gb"] += e["launch_counted_gb_per_step"]
                d["tflops_ms"] += e["tflops"] * e["ms_per_step"]
                d["classes"].append(cname)
            roofline = None
            if by_kernel:
                top = max(by_kernel, key=lambda k: by_kernel[k]["ms"])
                d = by_kernel[top]
                gbs = d["alg_gb"] / (d["ms"] * 1e-3)
                tfl = d["tflops_ms"] / d["ms"]
                tensor_bound = top in ("lstm_tc_kernel", "lstm_tcw_kernel", "rvq_tc_kernel")
                traffic = dram.get(top) if (dram and name == "cfg2") else None
                roofline = {"kernel": top, "classes": d["classes"], "share_of_step": d["ms"] / (total_ms / args.steps),
                            "bound": "tensor" if tensor_bound else "hbm",
                            "achieved": tfl if tensor_bound else gbs, "peak": peaks["tflops"] if tensor_bound else peaks["hbm_gbs"],
                            "unit": "TFLOP/s" if tensor_bound else "GB/s",
                            "frac": (tfl / peaks["tflops"]) if tensor_bound else (gbs / peaks["hbm_gbs"]),
                            "traffic": traffic, "traffic_source": dram_file and os.path.relpath(dram_file, ROOT),
                            "algorithmic_bytes_per_step": d["alg_gb"] * 1e9, "ms_per_step": d["ms"],
                            "launch_counted_bytes_per_step": d["launch_gb"] * 1e9,
                            "launch_traffic_ratio": d["launch_gb"] / d["alg_gb"] if d["alg_gb"] else None,
                            "also": {"bound": "hbm" if tensor_bound else "tensor", "achieved": gbs if tensor_bound else tfl,
                                     "peak": peaks["hbm_gbs"] if tensor_bound else peaks["tflops"],
                                     "unit": "GB/s" if tensor_bound else "TFLOP/s",
                                     "frac": (gbs / peaks["hbm_gbs"]) if tensor_bound else (tfl / peaks["tflops"])},
                            "peak_source": peaks["source"],
                            "note": "achieved = SURVEY.md section 8d bytes of the blocks this kernel runs (one fused kernel per SConv / "
                                    "ResBlock / LSTM: input read once, output written once; 4 B x elements x items) / its live launch time "
                                    "(CUDA events on the launching stream); `traffic` = DRAM bytes of its launches in one step from the "
                                    "committed ncu capture; launch_traffic_ratio = bytes counted per launch (every launch's reads + "
                                    "writes) / algorithmic bytes; FLOPs are algorithmic (2 M N K), executed as 3xTF32 split operands in "
                                    "the encoder / quantiser and as one TF32 pass in the decoder"}
            whole = {"algorithmic_gb_per_step": sum(e["algorithmic_gb_per_step"] for e in classes.values()),
                     "launch_counted_gb_per_step": sum(e["launch_counted_gb_per_step"] for e in classes.values()),
                     "hbm_frac_of_step": sum(e["algorithmic_gb_per_step"] for e in classes.values()) / (out["ms_per_step"] * 1e-3) / peaks["hbm_gbs"],
                     "tensor_frac_of_step": sum(e["tflops"] * e["ms_per_step"] for e in classes.values()) / out["ms_per_step"] / peaks["tflops"],
                     "dram_bytes_per_step_ncu": sum(dram.values()) if (dram and name == "cfg2") else None}
            tf32_peak = measure_tf32_peak(dev)
            cpu_baseline = None
            if world == 1 and not args.no_cpu_baseline:
                clips, seconds = cpu_sample(wl)
                cpu_throughput(spec, sd, wl["bandwidth"], 1, 1.0 if wl["model"] != "fork10hz" else 1000.0)  # warm-up
                v, spent, kind, how = cpu_throughput(spec, sd, wl["bandwidth"], clips, seconds)
                cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count() or 1, "kind": kind,
                                "sample": f"{clips} clip(s) x {seconds:g} s of the same workload, {spent:.1f} s of CPU work", "how": how}
            length = int(wl["seconds"] * spec.sample_rate)
            cfg = config_of(name, wl, world)
            cfg["cache"] = (f"inputs rotate over 3 buffers; per-layer activations ({wl['batch'] * length * 32 * 4 / 1e9:.2f} GB) far "
                            "exceed the 126 MB L2")
            line = {
                "metric": METRIC, "value": out["value"], "unit": "audio-s/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": out["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "clocks": out["clocks"], "e2e": out["e2e"], "gpu_launches": out["gpu_launches"], "roofline": roofline,
                "kernels": classes, "whole_step": whole, "cpu_baseline": cpu_baseline, "variants": out.get("variants", {}),
                "tf32_gemm_tflops_measured": tf32_peak, "also": also,
                # operand tiles in which the fp16 pair conversion of the fp32-accurate convs clipped a value over this whole run
                # (0: every activation stayed inside the fp16 range; ecb_f16_saturation_count)
                "f16_pair_saturated_tiles": nat.f16_saturation_count(),
            }
    if rank == 0 and line is not None:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
