"""CPU: the oracle (oracle/qoracle.c) against the reference-generated golden vectors (tests/golden/)."""
import numpy as np
import pytest

import oracle
from conftest import conv_fixture_names, load_conv_fixture, load_pack_kat

NP_DT = {"float32": np.float32, "float64": np.float64, "float16": np.float16, "int8": np.int8, "int32": np.int32}


@pytest.mark.parametrize("case", load_pack_kat(), ids=lambda c: f"b{c['n_bits']}{'s' if c['sign'] else 'u'}-{c['dtype']}-{len(c['values'])}")
def test_pack_kat(case):
    x = np.array(case["values"], dtype=NP_DT[case["dtype"]]).reshape(case["shape"])
    packed, des = oracle.tpack(x, case["n_bits"], case["sign"])
    assert packed.tobytes().hex() == case["packed_hex"]
    assert des.tolist() == case["des"]
    back = oracle.tunpack(packed, des)
    assert str(back.dtype) == case["unpacked_dtype"]
    assert np.array_equal(back.astype(np.int64), x.astype(np.int64))


def test_pack_errors():
    with pytest.raises(RuntimeError, match="out of range"):
        oracle.tpack(np.array([8.0]), 4, True)
    with pytest.raises(RuntimeError, match="out of range"):
        oracle.tpack(np.array([-1.0]), 4, False)
    with pytest.raises(RuntimeError, match=r"\(0, 8\]"):
        oracle.tpack(np.array([0.0]), 9, True)
    with pytest.raises(RuntimeError, match="too short"):
        oracle.tunpack(np.zeros(1, np.uint8), np.array([4, 1]))


@pytest.mark.parametrize("name", conv_fixture_names())
def test_conv_fixture(name):
    f = load_conv_fixture(name)
    # activation quantizer: bit-exact integers vs the reference's Quantizer (quantizer.py:215)
    qa = oracle.act_quantize(f["x"], float(f["a_scale"][0]), float(f["a_zero"][0]), float(f["qmin"][0]), float(f["qmax"][0]))
    assert np.array_equal(qa.astype(np.uint8), f["q_x"])
    # weights: tunpack of the packed stream == what the reference module unpacked on load
    qw = oracle.tunpack(f["w_packed"], f["w_des"])
    assert np.array_equal(qw, f["q_w"])
    assert not f["w_zero"].any()
    # integer form == the reference's packed forward and fake-quant forward (fp32 conv) within 1e-3
    acc, wsum = oracle.conv_acc(qa.astype(np.uint8), qw, f["stride"], f["pad"])
    out = oracle.dequant(acc, wsum, float(f["a_scale"][0]), float(f["a_zero"][0]), f["w_scale"], f["bias"])
    for ref in (f["out_packed"], f["out_fake"]):
        tol = 1e-3 * np.abs(ref) + 1e-4 * np.abs(ref).max()
        assert np.all(np.abs(out - ref) <= tol), float(np.abs(out - ref).max())


def test_weightonly_matches_float_conv():
    """The op's weight-only semantic (quantconv2d_float_input.cu:83-119) vs a float64 conv of the same operands."""
    rng = np.random.default_rng(0)
    N, C, H, W, K, R, S, stride, pad = 2, 5, 7, 6, 4, 3, 3, 2, 1
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    for n_bits, sign, per_tensor in ((4, True, False), (8, False, True), (5, True, True)):
        lo, hi = (-(1 << (n_bits - 1)), (1 << (n_bits - 1)) - 1) if sign else (0, (1 << n_bits) - 1)
        qw = rng.integers(lo, hi + 1, size=(K, C, R, S))
        packed, des = oracle.tpack(qw, n_bits, sign)
        n_s = 1 if per_tensor else K
        scale = (rng.random(n_s) * 0.1 + 0.01).astype(np.float32)
        zero = rng.integers(-3, 4, size=n_s).astype(np.float32)
        bias = rng.standard_normal(K).astype(np.float32)
        out = oracle.quantconv2d_float_input(x, packed, des, scale, zero, bias, stride, pad)
        wf = (qw.astype(np.float64) - zero.reshape(-1, 1, 1, 1)) * scale.reshape(-1, 1, 1, 1).astype(np.float64)
        xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
        P, Q = oracle.conv_out_hw(H, W, R, S, stride, pad)
        ref = np.zeros((N, K, P, Q))
        for p in range(P):
            for q in range(Q):
                patch = xp[:, :, p * stride:p * stride + R, q * stride:q * stride + S]
                ref[:, :, p, q] = np.einsum("ncrs,kcrs->nk", patch, wf) + bias
        assert np.allclose(out, ref, rtol=1e-4, atol=1e-4)


def test_linear_oracle_against_float64():
    """quantlinear_float_input restatement (weight-only): sequential fp32 FMA sum vs an exact float64 product."""
    rng = np.random.default_rng(3)
    B, in_f, out_f = 5, 70, 9
    x = rng.standard_normal((B, in_f)).astype(np.float32)
    qw = rng.integers(-8, 8, size=(out_f, in_f)).astype(np.int64)
    packed, des = oracle.tpack(qw, 4, True)
    w_scale = (rng.random(out_f) * 0.02 + 0.001).astype(np.float32)
    w_zero = rng.integers(-2, 3, size=out_f).astype(np.float32)
    bias = rng.standard_normal(out_f).astype(np.float32)
    got = oracle.quantlinear_float_input(x, packed, des, w_scale, w_zero, bias)
    wf = ((qw.astype(np.float32) - w_zero[:, None]) * w_scale[:, None]).astype(np.float64)
    want = x.astype(np.float64) @ wf.T + bias
    assert got.dtype == np.float32 and got.shape == (B, out_f)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5)
