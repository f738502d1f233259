"""tests/golden/wav24k_loudness.npz: the UNMODIFIED reference on real speech at three loudness levels.

Run in the build container only:   PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_wav.py

A 2 s excerpt of the reference's own test clip (/root/reference/test_24k.wav, read with scipy as SURVEY.md section 8d
describes) is stored as int16 samples -- test INPUT data, not code -- and run through the reference 24 kHz model at
24 kbps as a batch of three clips: the excerpt scaled by 0.1, 1 and 10 ("softer / louder x10"). The file
keeps the reference's outputs (audio, codes, quantized latents); weights and codebooks are functions of the stored seed,
as in make_golden.py. The GPU tests use it to pin the decoder precision default at a non-synthetic signal scale.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from encodec_b200 import synth  # noqa: E402
from oracle import make_golden as mg  # noqa: E402

GAINS = (0.1, 1.0, 10.0)
SEED = 700


def main():
    import scipy.io.wavfile as wavfile
    import torch
    sr, wav = wavfile.read(os.path.join(mg.REF, "test_24k.wav"))
    assert sr == 24000 and wav.dtype == np.int16
    excerpt = np.ascontiguousarray(wav[12 * sr:14 * sr])          # the loudest 2 s of the clip
    base = excerpt.astype(np.float32) / 32768.0
    x = np.stack([g * base for g in GAINS])[:, None, :].astype(np.float32)   # [3, 1, 48000]
    spec = synth.spec_24khz()
    torch.set_num_threads(os.cpu_count() or 1)
    model = mg.build_reference(spec, True)
    sd = synth.make_state_dict(spec, SEED, shared_codebook=False)
    mg.load_sd(model, sd)
    xt = torch.from_numpy(x)
    with torch.no_grad():
        emb0 = model.encoder(xt).numpy()
    frames0 = np.transpose(emb0, (0, 2, 1)).reshape(-1, spec.dimension)
    mean_vec, scales = mg.calibrate(frames0, spec.n_q, spec.bins, SEED + 2)
    cbs = synth.calibrated_codebooks(SEED + 2, mean_vec, scales, spec.bins)
    sd = synth.make_state_dict(spec, SEED, codebooks=cbs, shared_codebook=False)
    mg.load_sd(model, sd)
    model.set_target_bandwidth(24.0)
    with torch.no_grad():
        audio, codes, _, _ = model(xt)
        frames = model.encode(xt)
    q = frames[0]["quantized"].numpy()
    print("audio scale per item:", np.abs(audio.numpy()).max(axis=(1, 2)), "rms", np.sqrt((audio.numpy() ** 2).mean(axis=(1, 2))))
    np.savez_compressed(os.path.join(mg.OUT, "wav24k_loudness.npz"), excerpt=excerpt, gains=np.array(GAINS, np.float32),
                        seed=np.int64(SEED), mean_vec=mean_vec, scales=scales, audio=audio.numpy().astype(np.float32),
                        codes=codes.numpy().astype(np.int16), quantized=q.astype(np.float32))


if __name__ == "__main__":
    main()
