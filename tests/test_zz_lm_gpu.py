"""Entropy-coded `.ecdc` path on the GPU (csrc/lm.cu through the C ABI).

Parity bar: the LM's probabilities are float32 in the reference, compared within |p - p_ref| <= 1e-6 + 1e-4 * p_ref against
outputs of the unmodified reference (tests/golden/lm_ac.npz); everything downstream of a pdf (quantised cdf, coder bytes,
decoded symbols) is integer work and compared bit for bit; the batched (compression) and the step-wise (decompression)
evaluation of the LM must agree bit for bit, otherwise a stream would not decode.
(The file sorts last on purpose: these are the newest kernels, the older parity tests run before them.)"""
import io
import struct
import time
import zlib

import numpy as np
import pytest

from encodec_b200 import synth
from oracle import lm_oracle as lo
from tests import golden_cases as gc

pytestmark = pytest.mark.gpu
LM_ATOL, LM_RTOL = 1e-6, 1e-4


def build_lm(spec, seed):
    import torch
    from encodec_b200.lm import LMModel
    m = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads,
                hidden_scale=spec.hidden_scale, past_context=spec.past_context, max_period=spec.max_period)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_lm_state_dict(spec, seed).items()}, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", list(gc.LM_CASES))
def test_lm_matches_reference_golden(name):
    import torch
    spec, K, T, seed = gc.LM_CASES[name]
    z = gc.load_lm_golden()
    lm = build_lm(spec, seed)
    codes = gc.lm_case_codes(spec, K, T, seed)
    out = lm.frame_outputs(torch.from_numpy(codes)[None].cuda(), probas=True, cdf=True, sym_ranges=True)
    p = out["probas"][0].cpu().numpy()
    ref = z[f"{name}_probas"]
    err = np.abs(p - ref)
    print(f"{name}: max abs {err.max():.3e}, max rel {(err / np.maximum(ref, 1e-6)).max():.3e}")
    assert np.all(err <= LM_ATOL + LM_RTOL * ref)
    # cdf of the kernel == the reference's builder (restated, pinned) applied to the kernel's own probabilities, bit for bit
    cdf = out["cdf"][0].cpu().numpy().astype(np.int64)
    assert np.array_equal(cdf, lo.build_stable_quantized_cdf(p))
    # the coder's pair per symbol, and the bytes the host coder writes from them == the restated coder on those cdfs
    r = out["sym_ranges"][0].cpu().numpy().astype(np.int64)            # [T, K, 2]
    tt, kk = np.meshgrid(np.arange(T), np.arange(K), indexing="ij")
    s = codes.T
    assert np.array_equal(r[..., 1], cdf[tt, kk, s])
    assert np.array_equal(r[..., 0], np.where(s > 0, cdf[tt, kk, np.maximum(s - 1, 0)], 0))
    data = lm.encode_frames(torch.from_numpy(codes)[None].cuda())[0]
    assert data == lo.encode_frame(codes, cdf)
    same = float((cdf == lo.build_stable_quantized_cdf(ref)).all(axis=-1).mean())
    print(f"{name}: {len(data)} bytes (reference {len(z[f'{name}_bytes'])}); {100 * same:.0f} % of the cdfs identical to the reference's")


def test_lm_matches_reference_beyond_the_window():
    """300 steps with past_context = 262 against the reference's probabilities at selected steps (the batched pass), and the
    device decoding loop (cluster kernel, window sliding) returning the codes."""
    import torch
    name, spec, K, T, seed = gc.LM_LONG
    z = gc.load_lm_golden()
    lm = build_lm(spec, seed)
    codes = torch.from_numpy(gc.lm_case_codes(spec, K, T, seed))[None].cuda()
    p = lm.frame_outputs(codes, probas=True, sym_ranges=False)["probas"][0].cpu().numpy()[list(gc.LM_LONG_STEPS)]
    ref = z[f"{name}_probas"]
    frac = np.abs(p - ref) / (LM_ATOL + LM_RTOL * ref)
    print(f"{name}: worst fraction of the tolerance per stored step {[round(float(f.max()), 3) for f in frac]}")
    assert frac.max() <= 1.0
    data = lm.encode_frames(codes)[0]
    assert abs(len(data) - len(z[f"{name}_bytes"])) <= 2
    got, end = lm.decode_frame(torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda(), 0, K, T)
    assert torch.equal(got, codes[0]) and end == len(data)


def test_quantized_cdf_kernel_is_bit_exact_with_reference():
    """ecb_quantized_cdf on the reference's own probabilities / the coder cases' pdfs == the reference's cdfs (checksums)."""
    import torch
    from encodec_b200 import lm as elm
    z = gc.load_lm_golden()
    for name in gc.LM_CASES:
        cdf = elm.quantized_cdf(torch.from_numpy(z[f"{name}_probas"]).cuda()).cpu().numpy().astype(np.int64)
        assert zlib.crc32(cdf.tobytes()) == int(z[f"{name}_cdf_crc"]), name
    for name, (card, steps, seed) in gc.AC_CASES.items():
        pdfs, _ = gc.ac_case_pdfs(card, steps, seed)
        cdf = elm.quantized_cdf(torch.from_numpy(pdfs).cuda()).cpu().numpy().astype(np.int64)
        assert zlib.crc32(cdf.tobytes()) == int(z[f"{name}_cdf_crc"]), name


def test_device_decoder_is_bit_exact_with_reference():
    """The warp-parallel ArithmeticDecoder of the decoding loop, alone, on the streams the reference's ArithmeticCoder wrote
    with the reference's cdfs: the reference's symbols, every byte consumed."""
    import torch
    from encodec_b200 import lm as elm
    z = gc.load_lm_golden()
    cases = []
    for name, (card, steps, seed) in gc.AC_CASES.items():
        pdfs, symbols = gc.ac_case_pdfs(card, steps, seed)
        cases.append((name, lo.build_stable_quantized_cdf(pdfs), symbols, z[f"{name}_bytes"].tobytes()))
    for name, (spec, K, T, seed) in gc.LM_CASES.items():
        cdfs = lo.build_stable_quantized_cdf(z[f"{name}_probas"]).reshape(T * K, spec.card)
        cases.append((name, cdfs, gc.lm_case_codes(spec, K, T, seed).T.reshape(-1), z[f"{name}_bytes"].tobytes()))
    for name, cdfs, symbols, data in cases:
        buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
        got, used = elm.ac_decode_device(buf, torch.from_numpy(cdfs.astype(np.int32)).cuda())
        assert got.cpu().tolist() == list(symbols), name
        assert used == len(data), name
        with pytest.raises(EOFError, match="ended sooner"):
            elm.ac_decode_device(buf[: max(1, len(data) // 3)], torch.from_numpy(cdfs.astype(np.int32)).cuda())


@pytest.mark.parametrize("name", list(gc.LM_CASES))
def test_streaming_steps_equal_batched_pass_bit_for_bit(name):
    """LMModel.forward fed step by step / in ragged chunks with its state (the reference's streaming API, model.py:65-83)
    returns the very bits of the one-pass evaluation compression uses."""
    import torch
    spec, K, T, seed = gc.LM_CASES[name]
    lm = build_lm(spec, seed)
    codes = torch.from_numpy(gc.lm_case_codes(spec, K, T, seed + 100)).cuda()
    codes2 = torch.stack([codes, codes.flip(1)])                        # two independent streams
    whole = lm.frame_outputs(codes2, probas=True, sym_ranges=False)["probas"]          # [2, T, K, card]
    indices = torch.cat([torch.zeros_like(codes2[:, :, :1]), 1 + codes2[:, :, :-1]], dim=2)
    for chunks in ([1] * T, [3, 1, 7, T - 11]):
        states, offset, got = None, 0, []
        for n in chunks:
            probas, states, offset = lm(indices[:, :, offset: offset + n], states, offset)
            assert probas.shape == (2, spec.card, K, n)
            got.append(probas.permute(0, 3, 2, 1))
        assert offset == T
        assert torch.equal(torch.cat(got, dim=1), whole), chunks[:3]
    with pytest.raises(ValueError):
        lm(indices[:, :, :1], states, 3)


@pytest.mark.parametrize("name", list(gc.LM_CASES))
def test_entropy_coded_frames_round_trip_on_device(name):
    """encode_frames (batched LM + host coder) -> decode_frame (device loop): codes and byte positions, frames back to back
    with foreign bytes between them as in a real stream (the scale field, compress.py:64-65)."""
    import torch
    spec, K, T, seed = gc.LM_CASES[name]
    lm = build_lm(spec, seed)
    T2 = 2 * T + 5
    u = synth.hash_uniform(seed, "rt-codes", 3 * K * T2).reshape(3, K, T2)
    codes = torch.from_numpy(np.minimum((u ** 3 * spec.card).astype(np.int64), spec.card - 1)).cuda()   # skewed symbols
    chunks = lm.encode_frames(codes)
    blob = b"".join(b"\xde\xad\xbe\xef" + c for c in chunks)
    data = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
    import os
    # one captured step replayed / every launch from the host; the step's transformer as one cluster kernel / per-phase launches
    for graph, cl in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
        os.environ["ECB_LM_GRAPH"] = graph
        os.environ["ECB_LM_CLUSTER"] = cl
        try:
            pos = 0
            for i, c in enumerate(chunks):
                got, end = lm.decode_frame(data, pos + 4, K, T2)
                assert torch.equal(got, codes[i]), (name, i, graph, cl)
                assert end == pos + 4 + len(c)
                pos = end
            assert pos == len(blob)
        finally:
            os.environ.pop("ECB_LM_GRAPH", None)
            os.environ.pop("ECB_LM_CLUSTER", None)
    with pytest.raises(EOFError, match="ended sooner"):
        lm.decode_frame(data[: 4 + len(chunks[0]) // 2], 4, K, T2)
    # fewer codebooks than the model has (a lower bandwidth): the first K - 1 embeddings / heads only (model.py:79-82)
    if K > 1:
        c1 = lm.encode_frames(codes[:1, : K - 1])[0]
        got, end = lm.decode_frame(torch.frombuffer(bytearray(c1), dtype=torch.uint8).cuda(), 0, K - 1, T2)
        assert torch.equal(got, codes[0, : K - 1]) and end == len(c1)


@pytest.mark.parametrize("kind", ["24k", "48k"])
def test_compress_decompress_with_lm(kind):
    """compress(..., use_lm=True) -> decompress through the model (compress.py:28-156): header, per-frame scale, the decoded
    audio equals decoding the codes of encode(); the entropy-coded body is what the restated coder writes for the LM's cdfs."""
    import torch
    from encodec_b200 import compress as ec
    from tests import util_gpu as ug
    spec = synth.spec_24khz() if kind == "24k" else synth.spec_48khz()
    m = ug.build_model(spec, synth.make_state_dict(spec, seed=3), 6.0 if kind == "24k" else 12.0, True)
    lm = m.get_lm_model(state_dict={k: torch.from_numpy(v) for k, v in synth.make_lm_state_dict(
        synth.LMSpec(n_q=m.quantizer.n_q, card=m.quantizer.bins, past_context=int(3.5 * m.frame_rate)), 7).items()})
    assert lm.transformer.past_context == (262 if kind == "24k" else 525) and not lm.training
    length = 31000 if kind == "24k" else 2 * 47520 + 9000
    wav = torch.from_numpy(synth.make_audio(8, 1, spec.channels, length)[0]).cuda()
    t0 = time.time()
    blob = ec.compress(m, wav, use_lm=True, lm=lm)
    t1 = time.time()
    fo = io.BytesIO(blob)
    meta = ec.read_ecdc_header(fo)
    n_q = m.quantizer.get_num_quantizers_for_bandwidth(m.frame_rate, m.bandwidth)
    assert meta == {"m": m.name, "al": length, "nc": n_q, "lm": True}
    out, sr = ec.decompress(blob, m, lm=lm)
    t2 = time.time()
    frames = m.encode(wav[None])
    ref = m.decode([(f["codes"], f["scale"]) for f in frames])[0, :, :length]
    assert sr == spec.sample_rate and torch.equal(out, ref)
    body = io.BytesIO()
    for f in frames:
        if f["scale"] is not None:
            body.write(struct.pack("!f", f["scale"].cpu().item()))
        cdf = lm.frame_outputs(f["codes"][:1].contiguous(), cdf=True, sym_ranges=False)["cdf"][0].cpu().numpy()
        body.write(lo.encode_frame(f["codes"][0].cpu().numpy(), cdf))
    assert blob.endswith(body.getvalue()) and len(blob) == fo.tell() + len(body.getvalue())
    plain = ec.compress(m, wav)
    print(f"{kind}: {len(blob)} bytes with the (random-weight) LM, {len(plain)} without; compress {t1 - t0:.3f} s, "
          f"decompress {t2 - t1:.3f} s for {length / spec.sample_rate:.2f} s of audio")
    with pytest.raises(EOFError):
        ec.decompress(blob[: len(blob) - 40], m, lm=lm)


def test_lm_throughput_note():
    """Not a parity test: prints what the two directions cost at the reference's full size (n_q = 32, 10 s at 75 Hz)."""
    import torch
    spec = synth.LMSpec(n_q=32, card=1024, past_context=262)
    lm = build_lm(spec, 3)
    K, T = 32, 750
    u = synth.hash_uniform(4, "tp-codes", K * T).reshape(1, K, T)
    codes = torch.from_numpy(np.minimum((u * spec.card).astype(np.int64), spec.card - 1)).cuda()
    lm.coder_ranges(codes)
    torch.cuda.synchronize()
    t0 = time.time()
    r = lm.coder_ranges(codes)
    torch.cuda.synchronize()
    t1 = time.time()
    data = lm.encode_frames(codes)[0]
    t2 = time.time()
    buf = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    got, end = lm.decode_frame(buf, 0, K, T)
    t3 = time.time()
    got, end = lm.decode_frame(buf, 0, K, T)
    t4 = time.time()
    assert torch.equal(got, codes[0]) and end == len(data) and r.shape == (1, T, K, 2)
    import os
    os.environ["ECB_LM_GRAPH"] = "0"
    try:
        lm.decode_frame(buf, 0, K, T)
        t5 = time.time()
        got2, _ = lm.decode_frame(buf, 0, K, T)
        t6 = time.time()
    finally:
        os.environ.pop("ECB_LM_GRAPH", None)
    assert torch.equal(got2, codes[0])
    os.environ["ECB_LM_CLUSTER"] = "0"
    try:
        lm.decode_frame(buf, 0, K, T)
        t7 = time.time()
        got3, _ = lm.decode_frame(buf, 0, K, T)
        t8 = time.time()
    finally:
        os.environ.pop("ECB_LM_CLUSTER", None)
    assert torch.equal(got3, codes[0])
    print(f"per-phase launches instead of the cluster kernel: {1e6 * (t8 - t7) / T:.0f} us per step")
    print(f"LM 32 x 750: batched pass {1e3 * (t1 - t0):.2f} ms, pass + host coder {1e3 * (t2 - t1):.2f} ms, "
          f"device decoding loop {1e3 * (t4 - t3):.1f} ms ({1e6 * (t4 - t3) / T:.0f} us per step; "
          f"{1e6 * (t6 - t5) / T:.0f} us per step without the step graph), {len(data)} bytes")


def test_heads_in_row_chunks_are_bit_identical():
    """Frames with more rows than fit the logits scratch buffer are processed in row chunks (1024 rows; lowered here through
    ECB_LM_HEAD_ROWS): probabilities, cdfs and coder ranges must not depend on the chunking."""
    import os
    import torch
    spec, K, T, seed = gc.LM_CASES["lm24_k8"]
    lm = build_lm(spec, seed)
    u = synth.hash_uniform(seed, "chunk-codes", 3 * K * 37).reshape(3, K, 37)
    codes = torch.from_numpy(np.minimum((u * spec.card).astype(np.int64), spec.card - 1)).cuda()     # 111 rows
    whole = lm.frame_outputs(codes, probas=True, cdf=True, sym_ranges=True)
    os.environ["ECB_LM_HEAD_ROWS"] = "16"                                                          # 7 chunks, the last one ragged
    try:
        parts = lm.frame_outputs(codes, probas=True, cdf=True, sym_ranges=True)
    finally:
        os.environ.pop("ECB_LM_HEAD_ROWS", None)
    for key in ("probas", "cdf", "sym_ranges"):
        assert torch.equal(whole[key], parts[key]), key
