"""Rebuilds the inputs of the golden cases (tests/golden/*.npz) from seeds -- no reference needed.

The .npz files hold the OUTPUTS of the unmodified reference (made by oracle/make_golden.py in the build
container) plus two small calibration vectors; weights, audio and codebooks are pure functions of the
stored seed (encodec_b200.synth).
"""
import os

import numpy as np

from encodec_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_CASES = ["cfg1_24k_6kbps_shared", "24k_24kbps_ragged", "48k_24kbps_3seg"]
FORK_CASES = ["fork10hz_ln_r5541", "fork10hz_ln_r65521"]   # layer_norm, stride-1 stage, dimension 256
RVQ_CASE = "rvq_nq32_8k"


def spec_for(name):
    if name.startswith("fork10hz"):
        return synth.spec_fork10hz(tuple(int(ch) for ch in name.rsplit("_r", 1)[1]))
    return synth.spec_48khz() if name.startswith("48k") else synth.spec_24khz()


def load_model_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    seed, batch, length, n_q, distinct = (int(v) for v in z["meta"])
    spec = spec_for(name)
    x = synth.make_audio(seed + 1, batch, spec.channels, length)
    cbs = synth.calibrated_codebooks(seed + 2, z["mean_vec"], z["scales"], spec.bins)
    sd = synth.make_state_dict(spec, seed, codebooks=cbs, shared_codebook=not distinct)
    return dict(name=name, spec=spec, x=x, sd=sd, bandwidth=float(z["bandwidth"]), n_q=n_q, distinct=bool(distinct),
                audio=z["audio"], codes=z["codes"].astype(np.int64), emb=z["emb"], quantized=z["quantized"],
                scale=z["scale"], commit_loss_shape=tuple(int(v) for v in z["commit_loss_shape"]))


def load_rvq_case(name=RVQ_CASE):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    seed, n_frames, n_q, bins, dim = (int(v) for v in z["meta"])
    frames = synth.hash_normal(seed, "rvq-frames", (n_frames, dim))
    cbs = synth.hash_normal(seed, "rvq-codebooks", (n_q, bins, dim))
    return dict(frames=frames, codebooks=cbs, n_q=n_q, bins=bins, dim=dim, codes=z["codes"].astype(np.int64),
                quantized_head=z["quantized_head"], decoded_head=z["decoded_head"],
                quantized_sum=float(z["quantized_sum"]))


def frames_of(t_bdt):
    """[B, D, T] -> [B*T, D]"""
    return np.ascontiguousarray(np.transpose(t_bdt, (0, 2, 1))).reshape(-1, t_bdt.shape[1])
