"""`.ecdc` code stream without entropy coding: the oracle and the GPU packer against the bytes
the reference's own binary.BitPacker wrote (tests/golden/ecdc_bitpack.npz, oracle/make_golden_ecdc.py), and the
compress -> decompress round trip through the model (the reference's own test, compress.py:188-207, checks only shapes)."""
import io
import os

import numpy as np
import pytest

from oracle import ecdc_oracle as eo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ecdc_bitpack.npz")


def _cases():
    z = np.load(GOLDEN)
    return [(z[f"codes_{i}"], z[f"bytes_{i}"].tobytes(), int(z[f"bits_{i}"])) for i in range(int(z["n_cases"]))]


def test_oracle_packer_matches_reference_bytes():
    for codes, ref_bytes, bits in _cases():
        assert eo.pack_frame(codes, bits) == ref_bytes
        np.testing.assert_array_equal(eo.unpack_frame(ref_bytes, codes.shape[0], codes.shape[1], bits), codes)


@pytest.mark.gpu
def test_gpu_packer_is_byte_exact_with_reference():
    import torch
    from encodec_b200 import compress as ec
    for codes, ref_bytes, bits in _cases():
        t = torch.from_numpy(codes).cuda()
        got = ec.pack_codes(t, bits).cpu().numpy().tobytes()
        assert got == ref_bytes, (codes.shape, bits)
        # non-contiguous view, as model.encode returns it ([B, K, T] is a transposed view of [K, B, T])
        tv = torch.from_numpy(np.ascontiguousarray(codes.T)).cuda().t()
        assert ec.pack_codes(tv, bits).cpu().numpy().tobytes() == ref_bytes
        back = ec.unpack_codes(torch.frombuffer(bytearray(ref_bytes), dtype=torch.uint8).cuda(), codes.shape[0], codes.shape[1],
                               bits)
        np.testing.assert_array_equal(back.cpu().numpy(), codes)
    with pytest.raises(RuntimeError):
        ec.unpack_codes(torch.zeros(3, dtype=torch.uint8).cuda(), 8, 75, 10)   # stream too short


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["24k", "48k"])
def test_compress_decompress_round_trip(kind):
    import torch
    from encodec_b200 import compress as ec, synth
    from tests import util_gpu as ug
    spec = synth.spec_24khz() if kind == "24k" else synth.spec_48khz()
    sd = synth.make_state_dict(spec, seed=3)
    m = ug.build_model(spec, sd, 6.0 if kind == "24k" else 24.0, True)
    length = 31000 if kind == "24k" else 2 * 47520 + 9000     # ragged; the 48 kHz clip spans 3 segments
    wav = torch.from_numpy(synth.make_audio(8, 1, spec.channels, length)[0]).cuda()
    blob = ec.compress(m, wav)
    meta = ec.read_ecdc_header(io.BytesIO(blob))
    assert meta == {"m": m.name, "al": length, "nc": m.quantizer.get_num_quantizers_for_bandwidth(m.frame_rate, m.bandwidth),
                    "lm": False}
    out, sr = ec.decompress(blob, m)
    assert sr == spec.sample_rate and out.shape == wav.shape
    # the stream carries exactly the codes (and scales) of encode(): decoding it equals decoding those codes
    frames = m.encode(wav[None])
    ref = m.decode([(f["codes"], f["scale"]) for f in frames])[0, :, :length]
    assert torch.equal(out, ref)
    # ... and the bytes are what the oracle's packer writes for those codes
    body = io.BytesIO()
    for f in frames:
        if f["scale"] is not None:
            import struct
            body.write(struct.pack("!f", f["scale"].cpu().item()))
        body.write(eo.pack_frame(f["codes"][0].cpu().numpy(), m.bits_per_codebook))
    assert blob.endswith(body.getvalue())
    with pytest.raises(RuntimeError, match="No LM pre-trained"):      # reference model.py:275-278: unknown model name
        ec.compress(m, wav, use_lm=True)
