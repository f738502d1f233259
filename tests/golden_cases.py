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


# --- entropy-coded stream: tests/golden/lm_ac.npz, made by oracle/make_golden_lm.py -----------------
# name -> (LMSpec, K codebooks in the frame, T steps, seed). "lm24_k8": the LM EncodecModel.get_lm_model builds for the
# 24 kHz model (model.py:268-269) on a 6 kbps frame; "lmsmall": odd sizes, window (past_context) shorter than the frame.
LM_CASES = {
    "lm24_k8": (synth.LMSpec(n_q=32, card=1024, dim=200, num_layers=5, num_heads=8, past_context=262), 8, 24, 11),
    "lmsmall": (synth.LMSpec(n_q=3, card=96, dim=64, num_layers=2, num_heads=4, past_context=5), 3, 23, 12),
}
# A frame longer than the LM's window: probabilities are stored for LM_LONG_STEPS only (first steps, around the step at which
# the all-zero seed row and then real rows leave the window of past_context = 262, and the last step).
LM_LONG = ("lm24_long", synth.LMSpec(n_q=32, card=1024, dim=200, num_layers=5, num_heads=8, past_context=262), 4, 300, 31)
LM_LONG_STEPS = (0, 1, 150, 260, 261, 262, 263, 264, 299)
# name -> (cardinality, steps, seed): the shape of the reference's own coder test (ac.py:263-285)
AC_CASES = {"ac_card2": (2, 400, 21), "ac_card37": (37, 300, 22), "ac_card1024": (1024, 200, 23), "ac_card3999": (3999, 120, 24)}


def lm_case_codes(spec, K, T, seed):
    """codes [K, T] int64 of an LM case."""
    u = synth.hash_uniform(seed, "lm-codes", K * T)
    return np.minimum((u * spec.card).astype(np.int64), spec.card - 1).reshape(K, T)


def ac_case_pdfs(card, steps, seed):
    """(pdfs [steps, card] float32, symbols [steps] int64); every pdf value is n / 2**20 with integer n >= 1, so the
    float32 tensors are reproduced bit for bit anywhere; symbols are drawn from the pdf with integer arithmetic."""
    total = 1 << 20
    u = synth.hash_uniform(seed, "ac-pdf", steps * card).reshape(steps, card) ** 6
    mass = total - 64                      # just under 1, so the reference's own range check (ac.py:50) holds
    n = np.floor(u / u.sum(axis=1, keepdims=True) * (mass - card)).astype(np.int64) + 1
    n[:, 0] += mass - n.sum(axis=1)
    assert (n >= 1).all() and (n.sum(axis=1) == mass).all()
    pdfs = (n.astype(np.float64) / total).astype(np.float32)
    thr = np.minimum((synth.hash_uniform(seed, "ac-sym", steps) * mass).astype(np.int64), mass - 1)
    symbols = np.array([int(np.searchsorted(np.cumsum(n[i]), thr[i], side="right")) for i in range(steps)], dtype=np.int64)
    return pdfs, symbols


def load_lm_golden():
    return np.load(os.path.join(GOLDEN_DIR, "lm_ac.npz"))
