#!/usr/bin/env python
"""Headline benchmark: audio-seconds encoded+decoded per wall-second (EncodecModel.forward).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3]

N > 1 is launched by torchrun (one rank per GPU); the batch is sharded by clip with no data-path
collective, and the only exchange is an NCCL gather of codes and audio to rank 0 (north_star), inside the
timed region. Rank 0 prints ONE JSON line (see the task contract): `value` is measured with inputs
resident in HBM, `e2e` through the public API from pinned host buffers with H2D/D2H copies inside the
timed region, `roofline` describes the dominant kernel class (timed with CUDA events on its stream by the
library's own profiler hooks), `cpu_baseline` is the ATen port of the reference timed on the host cores.
`--impl reference` times the CPU port alone (the reference is pure Python/PyTorch and is not present on the GPU box;
oracle/torch_port.py restates its forward on the same ATen CPU kernels and is pinned, through the numpy oracle, to the
reference's own outputs in tests/golden).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "cfg2": dict(model="24k", bandwidth=24.0, batch=64, seconds=10.0,
                 desc="EnCodec 24 kHz causal mono, 24 kbps (n_q=32), batch 64 x 10 s per GPU, random-init weights"),
    "cfg1": dict(model="24k", bandwidth=6.0, batch=1, seconds=1.0,
                 desc="EnCodec 24 kHz causal mono, 6 kbps (n_q=8), 1 x 1 s"),
    "cfg3": dict(model="48k", bandwidth=24.0, batch=32, seconds=30.0,
                 desc="EnCodec 48 kHz stereo, 24 kbps (n_q=16), 1 s segments 1% overlap, batch 32 x 30 s per GPU"),
    # the fork's own training configuration (params/091224_l1.yaml): 32 x 4 hours of a 10 Hz signal
    "fork10hz": dict(model="fork10hz", bandwidth=0.08, batch=32, seconds=14400.0,
                     desc="the fork's 10 Hz model (layer_norm, ratios 6,5,5,2,1, dimension 256, 1024-wide LSTM), 0.08 kbps (n_q=8), "
                          "batch 32 x 144000 samples (4 h each)"),
    "cfg5": dict(model="24k", bandwidth=6.0, batch=512, seconds=10.0,
                 desc="EnCodec 24 kHz causal mono, 6 kbps (n_q=8), 512 x 10 s clips per GPU and step (long-form shard, "
                      "micro-batch sized to HBM)"),
}


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


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=p["hbm_gbs"], tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


def measure_tf32_peak(dev, seconds=0.4):
    """Dense TF32 GEMM rate of this GPU (cuBLAS through torch, 8192^3, run for ~`seconds`): the denominator for the
    3xTF32 kernels -- MEASURED_PEAKS.json holds the bf16 figure only ('TF32 peak ... measure it')."""
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


_PORT_PARAMS = {}


def cpu_port_throughput(spec, sd, bandwidth, clips, seconds, repeats=1):
    """The reference's forward restated on the ATen CPU kernels the reference itself uses (oracle/torch_port.py: mkldnn
    conv / RNN, MKL sgemm, all host threads), on a bounded sample; returns (audio-s/s, seconds spent)."""
    import torch
    from encodec_b200 import synth
    from oracle import torch_port as port
    torch.set_num_threads(os.cpu_count() or 1)
    length = int(seconds * spec.sample_rate)
    x = synth.make_audio(999, clips, spec.channels, length)
    params = _PORT_PARAMS.setdefault(id(sd), port.TorchParams(sd, spec.norm))   # weight-norm folded once, outside the timed call
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        port.forward(x, sd, spec, bandwidth, params)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return clips * seconds / best, best


def run_reference_arm(args, wl, rank, world):
    """`--impl reference`: the CPU restatement of the reference on this box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from encodec_b200 import synth
    spec = make_spec(wl["model"])
    sd = synth.make_state_dict(spec, seed=0)
    clips, seconds = cpu_sample(wl)
    for _ in range(args.warmup):
        cpu_port_throughput(spec, sd, wl["bandwidth"], 1, min(seconds, 1.0) if wl["model"] != "fork10hz" else 1000.0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_throughput(spec, sd, wl["bandwidth"], clips, seconds)
    dt = (time.perf_counter() - t0) / args.steps
    value = clips * seconds / dt
    cores = os.cpu_count() or 1
    sample = (f"{clips} clip(s) x {seconds:g} s of the {args.workload} workload per step (oracle/torch_port.py: the reference's "
              f"forward on the ATen CPU kernels it uses, {cores} threads)")
    line = {
        "impl": "reference", "metric": "audio-sec/sec encode+decode", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    if rank == 0 or world == 1:
        entry.build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = os.environ.get("ECB_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    import encodec_b200 as eb
    from encodec_b200 import _native as nat, dist as ebdist, synth

    spec = make_spec(wl["model"])
    sd = synth.make_state_dict(spec, seed=0)
    model = eb.EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                       model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment,
                                       name="unset", ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension,
                                       share_codebook=False)
    model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    model = model.to(dev).eval()
    model.set_target_bandwidth(wl["bandwidth"])

    batch = wl["batch"]
    length = int(wl["seconds"] * spec.sample_rate)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    n_rot = 3  # rotate inputs so that no step re-reads the previous step's input from L2
    xs = [(0.3 * torch.randn(batch, spec.channels, length, generator=gen, device=dev)).clamp_(-1, 1) for _ in range(n_rot)]
    audio_seconds_per_step = batch * wl["seconds"] * world

    def step(i, gather=True):
        audio, codes, _, _ = model(xs[i % n_rot])
        if world > 1 and gather:
            ebdist.gather_results(codes, audio, dst=0)
        return audio, codes

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.no_grad():
        for i in range(max(args.warmup, 1)):
            step(i)
        sync_all()

        # ---- timed region 1: inputs resident in HBM ------------------------------------------------
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = nat.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        sync_all()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        launches = nat.launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        ms_per_step = ms / args.steps
        value = audio_seconds_per_step / (ms_per_step / 1e3)

        # ---- timed region 2: end to end through the public API from pinned host memory --------------------
        x_host = xs[0].cpu().pin_memory()
        x_dev = torch.empty_like(xs[0])
        a0, c0 = step(0, gather=False)
        audio_host = torch.empty(a0.shape, dtype=a0.dtype).pin_memory()
        codes_host = torch.empty(c0.shape, dtype=c0.dtype).pin_memory()
        h2d = x_host.numel() * 4
        d2h = audio_host.numel() * 4 + codes_host.numel() * 8
        # (a) the plain call sequence a user of the reference writes, on one stream: copy in, forward, copy out
        sync_all()
        e0.record()
        for i in range(args.steps):
            x_dev.copy_(x_host, non_blocking=True)
            audio, codes, _, _ = model(x_dev)
            audio_host.copy_(audio, non_blocking=True)
            codes_host.copy_(codes, non_blocking=True)
            if world > 1:
                ebdist.gather_results(codes, audio, dst=0)
        e1.record()
        sync_all()
        ms_e2e_serial = e0.elapsed_time(e1)
        # (b) the package's host pipeline (encodec_b200.pipeline.HostPipeline): the same copies and the same forward per
        # step, on three streams so that the link time of the neighbouring batches hides behind the kernels
        from encodec_b200.pipeline import HostPipeline
        pipe = HostPipeline(model, depth=2)
        gather = (lambda a, c: ebdist.gather_results(c, a, dst=0)) if world > 1 else None
        for _ in pipe.run([x_host] * 2, after_forward=gather):
            pass
        sync_all()
        e0.record()
        for _ in pipe.run([x_host] * args.steps, after_forward=gather):
            pass
        pipe.join()
        e1.record()
        sync_all()
        ms_e2e = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e_value = audio_seconds_per_step / (ms_e2e / args.steps / 1e3)

        # ---- host link check: what the e2e number can be at best on this box (pinned 256 MB each way) -------------
        pcie = None
        if rank == 0:
            hb = torch.empty(64 * 1024 * 1024, dtype=torch.float32).pin_memory()
            db = torch.empty_like(hb, device=dev)
            db.copy_(hb, non_blocking=True)
            torch.cuda.synchronize(dev)
            g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            g0.record()
            db.copy_(hb, non_blocking=True)
            g1.record()
            hb.copy_(db, non_blocking=True)
            g2.record()
            torch.cuda.synchronize(dev)
            nb = hb.numel() * 4 / 1e9
            pcie = {"h2d_gbs": nb / (g0.elapsed_time(g1) * 1e-3), "d2h_gbs": nb / (g1.elapsed_time(g2) * 1e-3)}
            del hb, db

        # ---- opt-in variant: single-pass TF32 decoder (fp32-accurate encoder + quantiser unchanged) ------------------
        variants = {}
        if rank == 0 and world == 1 and spec.norm == "weight_norm" and wl["model"] == "24k":
            model.decoder.tf32 = True
            step(0, gather=False)
            torch.cuda.synchronize(dev)
            e0.record()
            for i in range(args.steps):
                step(i, gather=False)
            e1.record()
            torch.cuda.synchronize(dev)
            model.decoder.tf32 = False
            v_ms = e0.elapsed_time(e1) / args.steps
            variants["decoder_tf32"] = {
                "value": audio_seconds_per_step / (v_ms / 1e3), "unit": "audio-s/s", "ms_per_step": v_ms,
                "note": "SEANetDecoder.tf32=True: decoder convs as one TF32 pass; audio within 1e-4 max-abs / 2.3e-5 RMS of the "
                        "fp32-accurate result on the golden cases (bar 1e-3 / 1e-4); NOT the headline value"}

        # ---- per-kernel-class timing (CUDA events on the launching stream, inside the library) -----------
        prof = {}
        if rank == 0:
            torch.cuda.synchronize(dev)
            nat.profile_begin()
            for i in range(args.steps):
                step(i, gather=False)
            torch.cuda.synchronize(dev)
            prof = nat.profile_end()
    sync_all()

    if rank == 0:
        peaks = read_peaks()
        total_ms = sum(v["ms"] for v in prof.values()) or 1.0
        top_name = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
        roofline = None
        breakdown = {}
        for name, v in prof.items():
            breakdown[name] = {"launches_per_step": v["launches"] / args.steps, "ms_per_step": v["ms"] / args.steps,
                               "share": v["ms"] / total_ms,
                               "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0,
                               "gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0}
        # which roof bounds each kernel class: GEMM-shaped work with >= 128 channels, the RVQ distance GEMM and the LSTM
        # recurrence are tensor / FMA bound; the <= 64-channel convs, the edge convs and the element-wise passes move
        # bytes. The tensor roof is the measured dense bf16 rate (MEASURED_PEAKS.json); the
        # kernels compute in TF32 with split operands (3 products per algorithmic FLOP pair), fp32 accumulate.
        tensor_classes = ("tc_conv_wide", "conv_gemm", "rvq_encode", "lstm_recurrent")

        def roof(name):
            v = prof[name]
            sec = v["ms"] * 1e-3
            if name in tensor_classes:
                achieved = v["flops"] / sec / 1e12
                if name == "lstm_recurrent":
                    # the recurrence is fp32 CUDA-core work by design (W_hh lives in registers; a tensor-core version would
                    # have to re-read it from shared memory every step): its practical roof is the packed-FFMA2 issue rate
                    # measured on this pool with tools/fma_probe.cu (62 TFLOP/s on 148 SMs, register operands)
                    extra = {"also": {"bound": "fp32_fma", "achieved": achieved, "peak": 62.0, "unit": "TFLOP/s",
                                      "frac": achieved / 62.0,
                                      "note": "peak = measured fma.rn.f32x2 rate with register operands (tools/fma_probe.cu, "
                                              "profiles/r01_fma_probe.txt); nominal 2*128 lanes*148 SMs*1.9 GHz = 72 TFLOP/s"}}
                else:
                    extra = {}
                return {**extra, "kernel": name, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                        "frac": achieved / peaks["tflops"], "traffic": None,
                        "peak_source": f"{peaks['source']} bf16 dense (sustained); algorithmic FLOPs, computed as 3xTF32 split "
                                       "operands on tcgen05 (fp32 FFMA for the LSTM recurrence)",
                        "share_of_step": v["ms"] / total_ms}
            achieved = v["bytes"] / sec / 1e9
            return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                    "share_of_step": v["ms"] / total_ms}

        tf32_peak = measure_tf32_peak(dev)

        def tf32_view(alg_tflops):
            return {"bound": "tensor_tf32", "achieved": 3.0 * alg_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": 3.0 * alg_tflops / tf32_peak,
                    "note": "executed TF32 products (3 per algorithmic multiply-add: a_hi*w_hi, a_hi*w_lo, a_lo*w_hi) against "
                            "the dense TF32 GEMM rate measured in this run (cuBLAS 8192^3 through torch)"}

        rooflines = {}
        if top_name:
            rooflines = {name: roof(name) for name in prof if prof[name]["ms"] / total_ms >= 0.02}
            for name in ("tc_conv_wide", "rvq_encode"):
                if name in rooflines:
                    rooflines[name]["tf32"] = tf32_view(rooflines[name]["achieved"])
            # The dominant KERNEL is tc_conv_kernel (one template, profiled as two classes by channel width); its roofline
            # is taken over all of its launches, against the roof it sits closer to.
            tc = [n for n in ("tc_conv_narrow", "tc_conv_wide") if n in prof]
            tc_ms = sum(prof[n]["ms"] for n in tc)
            if tc and tc_ms >= max(v["ms"] for k, v in prof.items() if k not in tc):
                sec = tc_ms * 1e-3
                gbs = sum(prof[n]["bytes"] for n in tc) / sec / 1e9
                tfl = sum(prof[n]["flops"] for n in tc) / sec / 1e12
                hbm = {"kernel": "tc_conv_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": gbs / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                       "share_of_step": tc_ms / total_ms,
                       "also": {"bound": "tensor", "achieved": tfl, "peak": peaks["tflops"], "unit": "TFLOP/s",
                                "frac": tfl / peaks["tflops"],
                                "note": "algorithmic FLOPs (3 TF32 products each) vs measured dense bf16"},
                       "note": "algorithmic bytes: activations and weights read once, outputs written once; ncu DRAM traffic "
                               "of the <=64-channel launches equals the algorithmic bytes (profiles/r01_tc_res32b1_ncu.txt); "
                               "the >=128-channel launches run at 66 % tensor-pipe activity (profiles/r01_tc_down256_ncu.txt)"}
                hbm["traffic_evidence"] = {
                    "note": "ncu --set full DRAM bytes of single launches of this build's kernels (profiles/), per launch",
                    "tc_conv_kernel<32,3> res32.b1 (64 x 240000 rows)": {"dram_bytes": 1.966e9 + 1.928e9, "algorithmic_bytes": 3.93e9,
                                                                          "file": "profiles/r01_tc_res32b1_ncu.txt"},
                    "tc_conv_kernel<128,3> down256 (64 x 6000 rows)": {"dram_bytes": 0.56e9, "algorithmic_bytes": 0.55e9,
                                                                        "file": "profiles/r01_tc_down256_ncu.txt"},
                    "tc_res_kernel (64 x 240000 rows)": {"dram_bytes": 1.966e9 + 1.923e9, "algorithmic_bytes": 3.93e9,
                                                         "file": "profiles/r01_tc_res_fused_ncu.txt"}}
                if "tc_conv_wide" in rooflines:
                    hbm["wide_launches_tf32"] = rooflines["tc_conv_wide"].get("tf32")
                    hbm["narrow_launches_hbm_frac"] = rooflines.get("tc_conv_narrow", {}).get("frac")
                roofline = hbm
            else:
                roofline = roof(top_name)
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            clips, seconds = cpu_sample(wl)
            cpu_port_throughput(spec, sd, wl["bandwidth"], 1, 1.0 if wl["model"] != "fork10hz" else 1000.0)  # warm-up
            v, spent = cpu_port_throughput(spec, sd, wl["bandwidth"], clips, seconds)
            cpu_baseline = {"value": v, "unit": "audio-s/s", "cores": os.cpu_count() or 1, "kind": "port",
                            "sample": f"{clips} clip(s) x {seconds:g} s of the same workload, {spent:.1f} s of CPU work "
                                      "(oracle/torch_port.py: the reference's forward restated on the ATen CPU kernels the "
                                      "reference itself calls -- mkldnn conv / RNN, MKL sgemm -- all host threads)"}
        line = {
            "metric": "audio-sec/sec encode+decode", "value": value, "unit": "audio-s/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "global_batch": batch * world, "clip_seconds": wl["seconds"],
                       "parallelism": f"dp{world} (clips sharded, NCCL gather of codes+audio to rank 0)",
                       "cache": f"inputs rotate over {n_rot} buffers; per-layer activations "
                                f"({batch * length * 32 * 4 / 1e9:.2f} GB) far exceed the 126 MB L2"},
            "clocks": clocks, "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                                      "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                                      "api": "encodec_b200.pipeline.HostPipeline(model, depth=2).run(pinned host batches)",
                                      "single_stream_ms_per_step": ms_e2e_serial / args.steps},
            "gpu_launches": launches, "roofline": roofline, "rooflines": rooflines, "kernels": breakdown,
            "cpu_baseline": cpu_baseline, "host_link": pcie, "variants": variants, "tf32_gemm_tflops_measured": tf32_peak,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
