"""Entropy-coded `.ecdc` path, CPU side: the restatement (oracle/lm_oracle.py) and the host
arithmetic coder of the C-ABI library against outputs of the UNMODIFIED reference (tests/golden/lm_ac.npz, made by
oracle/make_golden_lm.py from the reference's LMModel, build_stable_quantized_cdf, ArithmeticCoder and ArithmeticDecoder).
No GPU is touched: ecb_ac_encode / ecb_ac_decode are host functions (the decoder is the very code the device loop runs)."""
import ctypes
import zlib

import numpy as np
import pytest

from encodec_b200 import synth
from oracle import lm_oracle as lo
from tests import golden_cases as gc

# LM probabilities are float32 in the reference; its restatements agree with it to a few float32 roundings of a 5-layer
# transformer. Tolerance of every LM comparison in this repository: |p - p_ref| <= 1e-6 + 1e-4 * p_ref.
LM_ATOL, LM_RTOL = 1e-6, 1e-4


def _crc(cdfs):
    return zlib.crc32(np.ascontiguousarray(cdfs, dtype=np.int64).tobytes())


@pytest.mark.parametrize("name", list(gc.LM_CASES))
def test_lm_oracle_matches_reference(name):
    spec, K, T, seed = gc.LM_CASES[name]
    z = gc.load_lm_golden()
    sd = synth.make_lm_state_dict(spec, seed)
    codes = gc.lm_case_codes(spec, K, T, seed)
    p = lo.lm_probas(sd, codes, num_layers=spec.num_layers, num_heads=spec.num_heads, past_context=spec.past_context,
                     max_period=spec.max_period)
    ref = z[f"{name}_probas"]
    assert p.shape == ref.shape == (T, K, spec.card)
    assert np.all(np.abs(p - ref) <= LM_ATOL + LM_RTOL * ref), float(np.abs(p - ref).max())
    # downstream of the reference's own probabilities everything is exact: cdfs (checksum) and the coder's bytes
    cdfs = lo.build_stable_quantized_cdf(ref)
    assert _crc(cdfs) == int(z[f"{name}_cdf_crc"])
    assert lo.encode_frame(codes, cdfs) == z[f"{name}_bytes"].tobytes()


def test_lm_oracle_matches_reference_beyond_the_window():
    """300 steps with past_context = 262: the seed row and then real rows drop out of the window. The reference's probabilities
    at the stored steps; the stream the restated coder writes from the restatement's own cdfs has the reference stream's length
    (the bytes themselves depend on the last bits of the probabilities)."""
    name, spec, K, T, seed = gc.LM_LONG
    z = gc.load_lm_golden()
    sd = synth.make_lm_state_dict(spec, seed)
    codes = gc.lm_case_codes(spec, K, T, seed)
    p = lo.lm_probas(sd, codes, num_layers=spec.num_layers, num_heads=spec.num_heads, past_context=spec.past_context)
    ref = z[f"{name}_probas"]
    got = p[list(gc.LM_LONG_STEPS)]
    assert np.all(np.abs(got - ref) <= LM_ATOL + LM_RTOL * ref), float((np.abs(got - ref) / (LM_ATOL + LM_RTOL * ref)).max())
    mine = lo.encode_frame(codes, lo.build_stable_quantized_cdf(p.astype(np.float32)))
    assert abs(len(mine) - len(z[f"{name}_bytes"])) <= 2          # same model of the data: the streams have the same length


@pytest.mark.parametrize("name", list(gc.AC_CASES))
def test_coder_oracle_matches_reference(name):
    card, steps, seed = gc.AC_CASES[name]
    z = gc.load_lm_golden()
    pdfs, symbols = gc.ac_case_pdfs(card, steps, seed)
    cdfs = lo.build_stable_quantized_cdf(pdfs)
    assert _crc(cdfs) == int(z[f"{name}_cdf_crc"])
    data = z[f"{name}_bytes"].tobytes()
    assert lo.encode_frame(symbols[None, :], cdfs[:, None, :]) == data
    dec = lo.ArithmeticDecoder(data)
    assert [dec.pull(c) for c in cdfs] == symbols.tolist()
    assert dec.bytes_consumed == len(data)


def _ranges(cdfs, symbols):
    """(cdf[s - 1] or 0, cdf[s]) per symbol: what ArithmeticCoder.push reads (ac.py:143-144)."""
    flat = cdfs.reshape(-1, cdfs.shape[-1])
    s = np.asarray(symbols).reshape(-1)
    lo_ = np.where(s > 0, flat[np.arange(len(s)), np.maximum(s - 1, 0)], 0)
    return np.stack([lo_, flat[np.arange(len(s)), s]], axis=1)


def _all_coder_cases():
    z = gc.load_lm_golden()
    for name, (card, steps, seed) in gc.AC_CASES.items():
        pdfs, symbols = gc.ac_case_pdfs(card, steps, seed)
        yield name, lo.build_stable_quantized_cdf(pdfs), symbols, z[f"{name}_bytes"].tobytes()
    for name, (spec, K, T, seed) in gc.LM_CASES.items():
        codes = gc.lm_case_codes(spec, K, T, seed)
        cdfs = lo.build_stable_quantized_cdf(z[f"{name}_probas"]).reshape(T * K, spec.card)
        yield name, cdfs, codes.T.reshape(-1), z[f"{name}_bytes"].tobytes()      # time-major, codebooks inside (compress.py:79)


def test_host_coder_is_bit_exact_with_reference():
    """ecb_ac_encode / ecb_ac_decode through the C ABI: the reference's bytes, the reference's symbols, every byte consumed."""
    from encodec_b200 import lm
    for name, cdfs, symbols, data in _all_coder_cases():
        assert lm.ac_encode(_ranges(cdfs, symbols)) == data, name
        got, used = lm.ac_decode(data, cdfs)
        assert got.tolist() == list(symbols), name
        assert used == len(data), name


def test_host_coder_errors_and_random_round_trips():
    from encodec_b200 import lm, _native as nat
    rng = np.random.default_rng(5)
    for card in (2, 3, 200, 1024):
        n = 500
        pdf = rng.random((n, card)).astype(np.float32) ** 8
        pdf /= pdf.sum(axis=1, keepdims=True) * 1.0001
        cdfs = lo.build_stable_quantized_cdf(pdf)
        symbols = rng.integers(0, card, n)
        data = lm.ac_encode(_ranges(cdfs, symbols))
        assert data == lo.encode_frame(symbols[None, :], cdfs[:, None, :])
        got, used = lm.ac_decode(data, cdfs)
        assert got.tolist() == symbols.tolist() and used == len(data)
        with pytest.raises(EOFError, match="ended sooner"):          # compress.py:143-144
            lm.ac_decode(data[: len(data) // 2], cdfs)
    # 28 range bits: beyond the exact-integer window of ac_core.h, the coder runs the reference's double arithmetic
    pdf = rng.random((300, 64)).astype(np.float32) ** 6
    pdf /= pdf.sum(axis=1, keepdims=True) * 1.0001
    cdfs28 = lo.build_stable_quantized_cdf(pdf, total_range_bits=28)
    sym28 = rng.integers(0, 64, 300)
    coder = lo.ArithmeticCoder(28)
    for sy, c in zip(sym28, cdfs28):
        coder.push(int(sy), c)
    want = coder.finish()
    assert lm.ac_encode(_ranges(cdfs28, sym28), 28) == want
    got, used = lm.ac_decode(want, cdfs28, 28)
    assert got.tolist() == sym28.tolist() and used == len(want)
    # a buffer that is too small is an error, not an overrun
    r = _ranges(cdfs, symbols).astype(np.int32)
    out = np.zeros(8, dtype=np.uint8)
    n_out = ctypes.c_int64(0)
    assert nat.lib.ecb_ac_encode(r.ctypes.data, len(r), 24, out.ctypes.data, out.size, ctypes.byref(n_out)) != 0
    assert "too small" in nat.last_error()
    bad = np.array([[5, 5]], dtype=np.int32)                          # empty range
    assert nat.lib.ecb_ac_encode(bad.ctypes.data, 1, 24, out.ctypes.data, out.size, ctypes.byref(n_out)) != 0


def test_lm_handle_conventions_without_gpu():
    from encodec_b200 import _native as nat
    h = ctypes.c_void_p()
    spec = nat.EcbLmSpec(32, 1024, 200, 5, 8, 800, 262, 10000.0)
    assert nat.lib.ecb_lm_create(ctypes.byref(spec), ctypes.byref(h)) == 0
    assert nat.lib.ecb_lm_cache_bytes(h, 1, 750) == 5 * 751 * 400 * 4
    assert nat.lib.ecb_lm_workspace_bytes(h, 750, 8) > 750 * (4 * 200 + 800) * 4
    assert nat.lib.ecb_lm_finalize(h, None, None) != 0 and "was not loaded" in nat.last_error()
    nat.lib.ecb_lm_destroy(h)
    bad = nat.EcbLmSpec(32, 1024, 200, 5, 3, 800, 262, 10000.0)      # 200 % 3 != 0
    assert nat.lib.ecb_lm_create(ctypes.byref(bad), ctypes.byref(h)) != 0 and "head" in nat.last_error()


def test_lm_module_state_dict_layout():
    """LMModel has the reference's parameter names and shapes (model.py:59-63, transformer.py:83-97)."""
    from encodec_b200.lm import LMModel
    spec = synth.LMSpec(n_q=3, card=96, dim=64, num_layers=2, num_heads=4, past_context=5)
    m = LMModel(spec.n_q, spec.card, dim=spec.dim, num_layers=spec.num_layers, num_heads=spec.num_heads,
                past_context=spec.past_context)
    sd = synth.make_lm_state_dict(spec, 1)
    own = m.state_dict()
    assert set(own) == set(sd)
    assert all(tuple(own[k].shape) == sd[k].shape for k in sd)
    with pytest.raises(NotImplementedError):
        LMModel(3, 96, dim=64, gelu=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        import torch
        m(torch.zeros(1, 3, 1, dtype=torch.long))
