"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference is imported as-is from /root/reference (never copied). Weights, audio and codebooks
come from ``encodec_b200.synth`` (pure hash functions of a seed), so only the reference's OUTPUTS and
two small calibration vectors are stored. Model constructor calls are the ones SURVEY.md section 8d
verified (``bins=1024`` passed explicitly -- fork delta D5). Codebooks are preset and ``inited`` set
to 1 before the first forward (otherwise the reference runs k-means, core_vq.py:143-153,229).
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from encodec_b200 import synth  # noqa: E402
from oracle import encodec_oracle as orc  # noqa: E402

REF = os.environ.get("ENCODEC_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def build_reference(spec, distinct_codebooks: bool):
    import torch
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from encodec.model import EncodecModel
    import quantization.core_vq as core_vq  # registered at top level by the reference's sys.path hack (D8)

    torch.manual_seed(0)
    m = EncodecModel._get_model(spec.target_bandwidths, spec.sample_rate, spec.channels, causal=spec.causal,
                                model_norm=spec.norm, audio_normalize=spec.normalize, segment=spec.segment,
                                name="unset", ratios=spec.ratios, bins=spec.bins, dimension=spec.dimension).eval()
    if distinct_codebooks:
        # the reference loops self.layers[:n_q] (core_vq.py:397) so independent layers are legal reference behaviour
        n_q = len(m.quantizer.vq.layers)
        m.quantizer.vq.layers = torch.nn.ModuleList([
            core_vq.VectorQuantization(dim=spec.dimension, codebook_size=spec.bins, codebook_dim=spec.dimension,
                                       decay=0.99, kmeans_init=True, kmeans_iters=50, threshold_ema_dead_code=2)
            for _ in range(n_q)])
    return m.eval()


def load_sd(model, sd):
    import torch
    missing = model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    return missing


def calibrate(emb_frames: np.ndarray, n_q: int, bins: int, seed: int):
    """SURVEY 8c recipe made portable: per-layer codeword scale from the current residual's std."""
    d = emb_frames.shape[1]
    mean_vec = emb_frames.mean(axis=0).astype(np.float32)
    residual = emb_frames.astype(np.float32).copy()
    scales = np.zeros(n_q, dtype=np.float32)
    base = synth.hash_normal(seed, "calib-codebook", (n_q, bins, d))
    for i in range(n_q):
        centred = residual - (mean_vec if i == 0 else 0.0)
        scales[i] = np.float32(0.32 * centred.std())
        cb = base[i] * scales[i]
        if i == 0:
            cb = cb + mean_vec[None]
        cb = cb.astype(np.float32)
        ind = orc.codebook_quantize(residual, cb)
        residual = residual - cb[ind]
    return mean_vec, scales


def run_case(name, spec, batch, length, bandwidth, distinct, seed):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    model = build_reference(spec, distinct)
    x = synth.make_audio(seed + 1, batch, spec.channels, length)
    xt = torch.from_numpy(x)
    n_q_max = spec.n_q

    # pass 1: weights only, to calibrate codebooks on the reference encoder's own output
    sd = synth.make_state_dict(spec, seed, shared_codebook=not distinct)
    load_sd(model, sd)
    with torch.no_grad():
        seg = spec.segment_length or length
        x0 = xt[:, :, :seg]
        if spec.normalize:
            mono = x0.mean(dim=1, keepdim=True)
            x0 = x0 / (1e-8 + mono.pow(2).mean(dim=2, keepdim=True).sqrt())
        emb0 = model.encoder(x0).numpy()
    frames0 = np.transpose(emb0, (0, 2, 1)).reshape(-1, spec.dimension)
    mean_vec, scales = calibrate(frames0, 1 if not distinct else n_q_max, spec.bins, seed + 2)
    cbs = synth.calibrated_codebooks(seed + 2, mean_vec, scales, spec.bins)
    sd = synth.make_state_dict(spec, seed, codebooks=cbs, shared_codebook=not distinct)
    load_sd(model, sd)
    model.set_target_bandwidth(bandwidth)

    with torch.no_grad():
        t0 = time.perf_counter()
        audio, codes, commit, cbl = model(xt)
        dt = time.perf_counter() - t0
        frames = model.encode(xt)
    emb_list, q_list, scale_list = [], [], []
    with torch.no_grad():
        stride = spec.segment_stride or length
        seg = spec.segment_length or length
        for off in range(0, length, stride):
            xs = xt[:, :, off:off + seg]
            if spec.normalize:
                mono = xs.mean(dim=1, keepdim=True)
                xs = xs / (1e-8 + mono.pow(2).mean(dim=2, keepdim=True).sqrt())
            emb_list.append(model.encoder(xs).numpy())
    for f in frames:
        q_list.append(f["quantized"].numpy())
        scale_list.append(np.zeros((batch, 1), np.float32) if f["scale"] is None else f["scale"].numpy())
    n_q = spec.n_q_for_bandwidth(bandwidth)
    assert codes.shape[1] == n_q, (codes.shape, n_q)

    # oracle cross-check (pins the restatement against the live reference)
    o_audio, o_codes, o_frames = orc.forward(x, sd, spec, bandwidth, np.float32)
    emb_ref = np.concatenate([np.transpose(e, (0, 2, 1)).reshape(-1, spec.dimension) for e in emb_list])
    ref_c = np.concatenate([np.transpose(f["codes"].numpy(), (1, 0, 2)).reshape(n_q, -1) for f in frames], axis=1)
    got_c = np.concatenate([np.transpose(f["codes"], (1, 0, 2)).reshape(n_q, -1) for f in o_frames], axis=1)
    score = orc.score_codes(emb_ref, orc.codebooks_from_state_dict(sd, n_q), ref_c, got_c)
    adiff = np.abs(o_audio - audio.numpy())
    uniq = [int(np.unique(codes.numpy()[:, i]).size) for i in range(n_q)]
    print(f"[{name}] ref forward {dt*1e3:.1f} ms on {torch.get_num_threads()} threads; codes {tuple(codes.shape)} "
          f"unique/layer {uniq[:4]}..{uniq[-1]}; oracle vs ref: codes {score}, audio max {adiff.max():.3e} "
          f"rms {np.sqrt((adiff**2).mean()):.3e}; commit_loss {tuple(commit.shape)} sum {float(commit.sum())}")
    assert score["hard"] == 0
    assert adiff.max() < 1e-4

    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        meta=np.array([seed, batch, length, n_q, int(distinct)], dtype=np.int64), bandwidth=np.float64(bandwidth),
        mean_vec=mean_vec, scales=scales,
        audio=audio.numpy().astype(np.float32), codes=codes.numpy().astype(np.int16),
        emb=np.stack(emb_list[:1])[0] if len(emb_list) == 1 else np.concatenate(emb_list, axis=-1),
        quantized=np.concatenate(q_list, axis=-1), scale=np.concatenate(scale_list, axis=-1),
        commit_loss_shape=np.array(commit.shape, dtype=np.int64))


def run_rvq_case(name, n_frames, n_q, bins, dim, seed):
    """Config-4-shaped check against core_vq.ResidualVectorQuantization.encode/forward/decode."""
    import torch
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    import encodec.model  # noqa: F401  (installs the top-level 'quantization' package, D8)
    import quantization.core_vq as core_vq
    frames = synth.hash_normal(seed, "rvq-frames", (n_frames, dim))
    cbs = synth.hash_normal(seed, "rvq-codebooks", (n_q, bins, dim))
    rvq = core_vq.ResidualVectorQuantization(num_quantizers=n_q, dim=dim, codebook_size=bins, codebook_dim=dim,
                                             kmeans_init=True).eval()
    rvq.layers = torch.nn.ModuleList([core_vq.VectorQuantization(dim=dim, codebook_size=bins, codebook_dim=dim)
                                      for _ in range(n_q)])
    for i, layer in enumerate(rvq.layers):
        layer._codebook.embed.copy_(torch.from_numpy(cbs[i]))
        layer._codebook.inited.fill_(1)
    rvq.eval()
    x = torch.from_numpy(np.ascontiguousarray(frames.T))[None]  # [1, D, N]
    with torch.no_grad():
        codes = rvq.encode(x)  # [n_q, 1, N]
        quantized, codes_f, losses = rvq(x, n_q=n_q)
        dec = rvq.decode(codes)
    assert torch.equal(codes, codes_f)
    o_q, o_codes, _ = orc.rvq_forward(x.numpy(), cbs, n_q)
    score = orc.score_codes(frames, cbs, codes[:, 0].numpy(), o_codes[:, 0])
    print(f"[{name}] oracle vs core_vq: {score}; quantized max diff {np.abs(o_q - quantized.numpy()).max():.3e}")
    assert score["hard"] == 0
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"),
                        meta=np.array([seed, n_frames, n_q, bins, dim], dtype=np.int64),
                        codes=codes[:, 0].numpy().astype(np.int16),
                        quantized_head=quantized[0, :, :256].numpy().astype(np.float32),
                        decoded_head=dec[0, :, :256].numpy().astype(np.float32),
                        quantized_sum=np.float64(quantized.double().sum().item()),
                        decoded_equals_quantized=np.array(bool(torch.equal(dec, quantized))))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    s24, s48 = synth.spec_24khz(), synth.spec_48khz()
    # config 1 of BASELINE.json: 24 kHz causal mono, 6 kbps (n_q 8), 1 x 1 s; fork-shipped shared codebook (D6)
    run_case("cfg1_24k_6kbps_shared", s24, 1, 24000, 6.0, False, 100)
    # ragged length, batch 2, all 32 distinct codebooks (config 2's bandwidth)
    run_case("24k_24kbps_ragged", s24, 2, 24077, 24.0, True, 200)
    # 48 kHz stereo: 2 full segments + 1 short (14 400 samples), 24 kbps (n_q 16), per-segment scale
    run_case("48k_24kbps_3seg", s48, 2, 2 * 47520 + 14400, 24.0, True, 300)
    # config 4 shape at a CPU-friendly size
    run_rvq_case("rvq_nq32_8k", 8192, 32, 1024, 128, 400)
    # SURVEY 8f row 3 -- the fork's own 10 Hz configurations (params/091224_l1.yaml): ConvLayerNorm, a stride-1 stage,
    # dimension 256, 0.08 kbps (n_q 8). A 4-ratio variant the yaml lists (512-wide LSTM) and the active 5-ratio one (1024).
    run_case("fork10hz_ln_r5541", synth.spec_fork10hz((5, 5, 4, 1)), 2, 4033, 0.08, True, 500)
    run_case("fork10hz_ln_r65521", synth.spec_fork10hz((6, 5, 5, 2, 1)), 2, 12100, 0.08, True, 600)
