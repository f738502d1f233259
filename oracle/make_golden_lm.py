"""Generates tests/golden/lm_ac.npz from the UNMODIFIED reference (run in the build container, where /root/reference
is mounted): the reference's LMModel, build_stable_quantized_cdf and ArithmeticCoder driven exactly as
compress.compress_to_file drives them with use_lm=True (compress.py:63-87). Only the reference's OUTPUTS are stored
(probabilities, byte streams, a checksum of the quantised cdfs); weights, codes and pdfs are re-derived from
encodec_b200.synth hashes.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_lm.py
"""
import io
import os
import sys
import zlib

import numpy as np

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.environ.get("ENCODEC_REFERENCE", "/root/reference"))

import torch  # noqa: E402
from encodec.model import LMModel  # noqa: E402  (the reference's own classes)
from encodec.quantization.ac import ArithmeticCoder, ArithmeticDecoder, build_stable_quantized_cdf  # noqa: E402

from encodec_b200 import synth  # noqa: E402
from oracle import lm_oracle  # noqa: E402
from tests.golden_cases import LM_CASES, AC_CASES, LM_LONG, LM_LONG_STEPS, lm_case_codes, ac_case_pdfs  # noqa: E402


def run_lm_case(spec: synth.LMSpec, K: int, T: int, seed: int):
    lm = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads,
                 hidden_scale=spec.hidden_scale, past_context=spec.past_context, max_period=spec.max_period).eval()
    sd = synth.make_lm_state_dict(spec, seed)
    lm.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    codes = lm_case_codes(spec, K, T, seed)
    frame = torch.from_numpy(codes)[None]                                # [1, K, T]
    fo = io.BytesIO()
    coder = ArithmeticCoder(fo)
    states, offset = None, 0
    input_ = torch.zeros(1, K, 1, dtype=torch.long)
    probas_all, cdfs = [], []
    for t in range(T):                                                   # compress.py:72-85
        with torch.no_grad():
            probas, states, offset = lm(input_, states, offset)
        input_ = 1 + frame[:, :, t: t + 1]
        probas_all.append(probas[0, :, :, 0].T.contiguous().numpy().copy())      # [K, card]
        for k, value in enumerate(frame[0, :, t].tolist()):
            q_cdf = build_stable_quantized_cdf(probas[0, :, k, 0], coder.total_range_bits, check=False)
            cdfs.append(q_cdf.numpy().copy())
            coder.push(value, q_cdf)
    coder.flush()
    data = fo.getvalue()
    # the reference decodes its own stream
    fo.seek(0)
    dec = ArithmeticDecoder(fo)
    i = 0
    for t in range(T):
        for k in range(K):
            assert dec.pull(torch.from_numpy(cdfs[i])) == codes[k, t]
            i += 1
    assert fo.tell() == len(data)
    probas_np = np.stack(probas_all)                                     # [T, K, card]
    cdfs_np = np.stack(cdfs).reshape(T, K, -1)
    # restatement checks at generation time
    mine = lm_oracle.build_stable_quantized_cdf(probas_np)
    assert np.array_equal(mine, cdfs_np), "cdf restatement differs from the reference"
    assert lm_oracle.encode_frame(codes, cdfs_np) == data, "coder restatement differs from the reference"
    p64 = lm_oracle.lm_probas(sd, codes, num_layers=spec.num_layers, num_heads=spec.num_heads,
                              past_context=spec.past_context, max_period=spec.max_period)
    err = np.abs(p64 - probas_np).max()
    rel = (np.abs(p64 - probas_np) / np.maximum(probas_np, 1e-6)).max()
    print(f"  LM restatement (float64) vs reference: max abs {err:.3e}, max rel {rel:.3e}; stream {len(data)} bytes "
          f"for {K * T} symbols ({8 * len(data) / (K * T):.2f} bits each)")
    return probas_np.astype(np.float32), np.frombuffer(data, dtype=np.uint8), zlib.crc32(cdfs_np.astype(np.int64).tobytes())


def run_ac_case(card: int, steps: int, seed: int):
    pdfs, symbols = ac_case_pdfs(card, steps, seed)
    fo = io.BytesIO()
    coder = ArithmeticCoder(fo)
    cdfs = []
    for pdf, s in zip(pdfs, symbols):                                    # ac.py test(): :263-285
        q = build_stable_quantized_cdf(torch.from_numpy(pdf), coder.total_range_bits)
        cdfs.append(q.numpy().copy())
        coder.push(int(s), q)
    coder.flush()
    data = fo.getvalue()
    fo.seek(0)
    dec = ArithmeticDecoder(fo)
    for q, s in zip(cdfs, symbols):
        assert dec.pull(torch.from_numpy(q)) == s
    assert fo.tell() == len(data)          # the decoder has read every byte the coder wrote (the next frame follows)
    cdfs_np = np.stack(cdfs)
    assert np.array_equal(lm_oracle.build_stable_quantized_cdf(pdfs), cdfs_np)
    assert lm_oracle.encode_frame(symbols[None, :], cdfs_np[:, None, :]) == data
    print(f"  AC card {card} steps {steps}: {len(data)} bytes")
    return np.frombuffer(data, dtype=np.uint8), zlib.crc32(cdfs_np.astype(np.int64).tobytes())


def main():
    out = {}
    for name, (spec, K, T, seed) in LM_CASES.items():
        print("LM case", name)
        probas, data, crc = run_lm_case(spec, K, T, seed)
        out[f"{name}_probas"] = probas
        out[f"{name}_bytes"] = data
        out[f"{name}_cdf_crc"] = np.int64(crc)
    name, spec, K, T, seed = LM_LONG
    print("LM case", name)
    probas, data, crc = run_lm_case(spec, K, T, seed)
    out[f"{name}_probas"] = probas[list(LM_LONG_STEPS)]
    out[f"{name}_bytes"] = data
    out[f"{name}_cdf_crc"] = np.int64(crc)
    for name, (card, steps, seed) in AC_CASES.items():
        data, crc = run_ac_case(card, steps, seed)
        out[f"{name}_bytes"] = data
        out[f"{name}_cdf_crc"] = np.int64(crc)
    path = os.path.join(ROOT, "tests", "golden", "lm_ac.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
