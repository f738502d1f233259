"""GPU unit tests of the tcgen05 implicit-GEMM convolution (csrc/tc_conv.cu) through the C ABI's diagnostic entry.

Each case restates the reference arithmetic in float64 numpy -- SConv1d's reflect-padded strided conv
(/root/reference/encodec/modules/conv.py:202-221), SConvTranspose1d as a 2-tap GEMM over frames with zero padding
(conv.py:241-263), the fused 1x1 shortcut of SEANetResnetBlock (modules/seanet.py:63-64) and the LSTM input
projection (modules/lstm.py:24) -- and compares the kernel's raw / ELU outputs and reflected halo rows with it.
Tolerance: split-operand (3xTF32) mode must be fp32-accurate: its error against float64 may be at most 4x the
error of a float32 numpy matmul of the same operands plus 1e-6 of the output scale; single-pass TF32 mode <= 2e-3.
"""
import numpy as np
import pytest
import torch

from encodec_b200 import _native as nat

pytestmark = pytest.mark.gpu

HALO = 16


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def _run_case(T, C0, taps, stride, pad_left, N, items, zero_pad=False, C1=0, split=3, out_halo=0, seed=0,
              both=False, round_out=0, x_scale=1.0, check=True):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((items, T, C0)) * x_scale).astype(np.float32)
    M = -(-T // stride) if not zero_pad else T
    ktot = taps * C0 + C1
    w = (rng.standard_normal((ktot, N)) / np.sqrt(ktot)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    need_right = (M - 1) * stride + taps - pad_left - T
    if zero_pad:
        xp = np.pad(x, ((0, 0), (pad_left, max(need_right, 0)), (0, 0)))
        buf = x.copy()
        a0_first, a0_rows = 0, T
    else:
        assert pad_left <= HALO and need_right <= HALO
        buf = np.pad(x, ((0, 0), (HALO, HALO), (0, 0)), mode="reflect")
        xp = buf[:, HALO - pad_left:, :]
        a0_first, a0_rows = -HALO, T + 2 * HALO
    # float64 reference
    A = np.zeros((items, M, ktot), dtype=np.float64)
    for j in range(taps):
        A[:, :, j * C0:(j + 1) * C0] = xp[:, j:j + (M - 1) * stride + 1:stride, :]
    x1 = None
    if C1:
        x1 = rng.standard_normal((items, M, C1)).astype(np.float32)
        A[:, :, taps * C0:] = x1
    ref = A @ w.astype(np.float64) + bias.astype(np.float64)
    ref32 = (A.astype(np.float32) @ w + bias).astype(np.float64)   # what an fp32 BLAS gives
    err32 = np.abs(ref32 - ref).max() / np.abs(ref).max()

    dev = torch.device("cuda")
    t_buf = torch.from_numpy(buf).to(dev)
    t_w = torch.from_numpy(w).to(dev)
    t_b = torch.from_numpy(bias).to(dev)
    t_x1 = torch.from_numpy(x1).to(dev) if C1 else None
    rows_out = M + 2 * out_halo
    outs = {}
    for name in (["raw", "elu"] if both else ["raw"]):
        outs[name] = torch.full((items, rows_out, N), float("nan"), dtype=torch.float32, device=dev)
    a0_ptr = t_buf.data_ptr()  # (item 0, sample a0_first)
    def optr(t):
        return t.data_ptr() + out_halo * N * 4
    rc = nat.lib.ecb_debug_tc_conv(
        a0_ptr, buf.shape[1] * C0, C0, a0_first, a0_rows, taps, stride, pad_left,
        t_x1.data_ptr() if C1 else None, M * C1, C1, M,
        t_w.data_ptr(), t_b.data_ptr(), N, M, items,
        optr(outs["raw"]), optr(outs["elu"]) if both else None, rows_out * N, out_halo, round_out, split,
        nat.stream_ptr(dev))
    nat.check(rc)
    torch.cuda.synchronize()
    if not check:
        return {}
    scale = np.abs(ref).max()
    tol = (4 * err32 + 1e-6) if split in (2, 3) else 2e-3
    res = {"fp32_blas": err32}
    for name, t in outs.items():
        got = t.cpu().numpy().astype(np.float64)
        want = ref if name == "raw" else _elu(ref)
        body = got[:, out_halo:out_halo + M, :]
        assert np.isfinite(body).all(), (name, "non-finite values in the output body")
        err = np.abs(body - want).max() / scale
        res[name] = err
        assert err < tol, (name, err)
        if out_halo:
            wantp = np.pad(want, ((0, 0), (out_halo, out_halo), (0, 0)), mode="reflect")
            assert np.isfinite(got).all(), (name, "halo rows were not all written")
            herr = np.abs(got - wantp).max() / scale
            assert herr < tol, (name, "halo", herr)
    return res


# fp32-accurate operand schemes: 3 = split TF32 (a, a_lo; w_hi, w_lo), 2 = fp16 pair (a1 + 2^-11 a2, w1 + 2^-11 w2)
ACCURATE = [3, 2]


@pytest.mark.parametrize("split", ACCURATE)
def test_k3_conv_narrow_dual_output_with_halo(split):
    # SEANetResnetBlock first conv at 32 channels (hidden padded to 32), ragged length, both outputs + halos
    r = _run_case(T=1000, C0=32, taps=3, stride=1, pad_left=2, N=32, items=3, out_halo=HALO, both=True, split=split)
    print("k3 narrow", split, r)


@pytest.mark.parametrize("split", ACCURATE)
def test_fused_shortcut_two_sources(split):
    r = _run_case(T=777, C0=32, taps=1, stride=1, pad_left=0, N=64, items=2, C1=64, both=True, out_halo=HALO, split=split)
    print("fused shortcut", split, r)


@pytest.mark.parametrize("split", ACCURATE)
def test_strided_down_conv_causal_extra_padding(split):
    # k = 2s = 8, causal pad_left = 4, T not a multiple of the stride -> reflected "extra" padding on the right
    r = _run_case(T=4001, C0=64, taps=8, stride=4, pad_left=4, N=128, items=2, out_halo=HALO, split=split)
    print("down conv", split, r)


@pytest.mark.parametrize("split", ACCURATE)
def test_strided_down_conv_noncausal(split):
    # non-causal: padding_total = 5 split 3 left / 2 right (conv.py:215-219)
    r = _run_case(T=3003, C0=128, taps=10, stride=5, pad_left=3, N=256, items=2, split=split)
    print("down conv s5", split, r)


@pytest.mark.parametrize("split", ACCURATE)
def test_transposed_conv_as_two_tap_gemm_zero_padded(split):
    r = _run_case(T=600, C0=256, taps=2, stride=1, pad_left=1, N=5 * 128, items=3, zero_pad=True, split=split)
    print("convtr", split, r)


@pytest.mark.parametrize("split", ACCURATE)
def test_k7_conv_wide(split):
    r = _run_case(T=300, C0=512, taps=7, stride=1, pad_left=6, N=128, items=2, split=split)
    print("k7 wide", split, r)


def test_fp16_pair_range():
    """fp16 pair operands over the range the scheme is specified for: activations scaled to ~1e-4 (fp16 subnormals in a1, the
    remainder keeps the absolute error) and to ~1e4 (close to the saturation point 65504) stay fp32-accurate relative to the
    output scale."""
    for scale in (1e-4, 1.0, 1e4):
        r = _run_case(T=640, C0=128, taps=3, stride=1, pad_left=1, N=128, items=2, split=2, seed=5, x_scale=scale)
        print("fp16 pair, activation scale", scale, r)


def test_lstm_projection_shape_split3_and_split1():
    r2 = _run_case(T=750, C0=512, taps=1, stride=1, pad_left=0, N=2048, items=2, zero_pad=True, split=2)
    print("lstm proj fp16 pair", r2)
    r3 = _run_case(T=750, C0=512, taps=1, stride=1, pad_left=0, N=2048, items=2, zero_pad=True)
    r1 = _run_case(T=750, C0=512, taps=1, stride=1, pad_left=0, N=2048, items=2, zero_pad=True, split=1)
    print("lstm proj", r3, r1)


def test_single_pass_rounded_output():
    r = _run_case(T=512, C0=64, taps=3, stride=1, pad_left=2, N=64, items=2, split=1, round_out=1, out_halo=HALO,
                  both=True)
    print("split1", r)


def test_many_tiles_persistent_schedule():
    # more tiles than SMs so every CTA loops (accumulator double buffering, stage phase wrap-around)
    r = _run_case(T=128 * 40 + 5, C0=32, taps=3, stride=1, pad_left=2, N=32, items=9, out_halo=HALO, both=True)
    print("many tiles", r)


def test_fp16_pair_saturation_is_counted():
    """Activations beyond the fp16 range are clipped by the pair conversion -- and counted, so that a caller can tell
    (ecb_f16_saturation_count): none on ordinary inputs, some when the input is scaled past 65504."""
    nat.f16_saturation_count(reset=True)
    _run_case(T=640, C0=64, taps=3, stride=1, pad_left=1, N=64, items=2, split=2, seed=6)
    assert nat.f16_saturation_count() == 0
    _run_case(T=640, C0=64, taps=3, stride=1, pad_left=1, N=64, items=2, split=2, seed=6, x_scale=1e5, check=False)
    assert nat.f16_saturation_count(reset=True) > 0
    assert nat.f16_saturation_count() == 0
