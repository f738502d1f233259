"""The numpy oracle against the reference's own outputs (golden fixtures) -- runs on CPU everywhere."""
import numpy as np
import pytest

from oracle import encodec_oracle as orc
from tests import golden_cases as gc


@pytest.mark.parametrize("name", gc.MODEL_CASES + gc.FORK_CASES)
def test_oracle_forward_matches_reference(name):
    case = gc.load_model_case(name)
    spec = case["spec"]
    audio, codes, frames = orc.forward(case["x"], case["sd"], spec, case["bandwidth"], np.float32)
    assert codes.shape == case["codes"].shape
    n_q = case["n_q"]
    # golden emb/codes are concatenated over segments along time: [B, D, sum T_f] / [B, n_q, sum T_f]
    emb_ref = gc.frames_of(case["emb"])
    ref_c = np.transpose(case["codes"], (1, 0, 2)).reshape(n_q, -1)
    got_c = np.transpose(codes, (1, 0, 2)).reshape(n_q, -1)
    score = orc.score_codes(emb_ref, orc.codebooks_from_state_dict(case["sd"], n_q), ref_c, got_c)
    assert score["hard"] == 0, score
    assert score["mismatched"] <= 1e-3 * score["compared"], score
    diff = np.abs(audio - case["audio"])
    assert diff.max() < 2e-5, diff.max()
    assert np.sqrt((diff ** 2).mean()) < 2e-6
    emb = np.concatenate([f["emb"] for f in frames], axis=-1)
    assert np.abs(emb - case["emb"]).max() < 1e-5
    if spec.normalize:
        scale = np.concatenate([f["scale"] for f in frames], axis=-1)
        np.testing.assert_allclose(scale, case["scale"], rtol=1e-6)


def test_oracle_rvq_matches_core_vq():
    case = gc.load_rvq_case()
    x = np.ascontiguousarray(case["frames"].T)[None]
    q, codes, _ = orc.rvq_forward(x, case["codebooks"], case["n_q"])
    score = orc.score_codes(case["frames"], case["codebooks"], case["codes"], codes[:, 0])
    assert score["hard"] == 0 and score["mismatched"] <= 2, score
    if score["mismatched"] == 0:
        np.testing.assert_array_equal(q[0][:, :256], case["quantized_head"])
    dec = orc.rvq_decode(case["codes"][:, None, :], case["codebooks"])
    np.testing.assert_array_equal(dec[0][:, :256], case["decoded_head"])


def test_overlap_add_weights_single_frame_identity():
    # a sample covered by one frame: (w*x)/w must stay within 1 ulp of x (utils.py:52-56)
    x = np.linspace(-1, 1, 1000, dtype=np.float32)[None, None]
    y = orc.linear_overlap_add([x], 990)
    np.testing.assert_allclose(y, x, rtol=3e-7, atol=1e-9)


def test_padding_rules():
    # SConv1d output length is ceil(L / stride) (conv.py:55-62), checked on ragged lengths
    for length in (1, 7, 320, 321, 24077):
        for k, s in ((4, 2), (8, 4), (10, 5), (16, 8), (7, 1)):
            extra = orc.get_extra_padding_for_conv1d(length, k, s, k - s)
            t_out = (length + (k - s) + extra - k) // s + 1
            assert t_out == -(-length // s)
