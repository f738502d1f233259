"""The ATen restatement used for the CPU baseline (oracle/torch_port.py) against the numpy oracle and the
reference's golden outputs (tests/golden): same codes, same audio."""
import numpy as np
import pytest

from oracle import encodec_oracle as orc
from oracle import torch_port as tp
from tests import golden_cases as gc


@pytest.mark.parametrize("name", gc.MODEL_CASES + gc.FORK_CASES)
def test_torch_port_matches_reference_golden(name):
    case = gc.load_model_case(name)
    spec = case["spec"]
    audio, codes = tp.forward(case["x"], case["sd"], spec, case["bandwidth"])
    n_q = case["codes"].shape[1]
    score = orc.score_codes(gc.frames_of(case["emb"]), orc.codebooks_from_state_dict(case["sd"], n_q),
                            np.transpose(case["codes"], (1, 0, 2)).reshape(n_q, -1),
                            np.transpose(codes, (1, 0, 2)).reshape(n_q, -1))
    assert score["hard"] == 0, score
    if score["mismatched"] == 0:
        assert np.abs(audio - case["audio"]).max() < 1e-5
