"""GPU: the `quant_engine` op surface (what the reference's QuantConv2dOp2 calls, quantconv2dop.py:44-66)."""
import numpy as np
import pytest
import torch

import oracle
from gpu_util import random_conv_case
from test_conv_gpu import assert_close_1e3, oracle_case

pytestmark = pytest.mark.gpu


def _tensors(c, dev="cuda"):
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    K = c["shape"].K
    return dict(x=t(c["x"]), packed=t(c["packed"]), des=t(c["des"]), w_scale=t(c["w_scale"]).reshape(-1, 1, 1, 1),
                w_zero=torch.zeros(c["w_scale"].size, 1, 1, 1, device=dev), bias=t(c["bias"]),
                a_scale=torch.tensor(c["a_scale"], device=dev).reshape(1, 1, 1, 1),
                a_zero=torch.tensor(c["a_zero"], device=dev).reshape(1, 1, 1, 1),
                qmin=torch.tensor(int(c["qmin"]), device=dev), qmax=torch.tensor(int(c["qmax"]), device=dev))


def test_fused_op_keyword_extension(engine):
    c = random_conv_case(5, 2, 64, 14, 14, 96, 3, 1, 1)
    T = _tensors(c)
    out = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 1,
                                         input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=T["qmin"],
                                         input_qmax=T["qmax"])
    _, _, ref = oracle_case(c)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == ref.shape
    assert_close_1e3(out.cpu().numpy(), ref)
    # second call hits the prepared-weight cache and python-number qmin/qmax work too
    out2 = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 1,
                                          input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=0, input_qmax=255)
    assert torch.equal(out, out2)
    # in-place weight update invalidates the cache (version bump)
    T["packed"].zero_()
    out3 = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], None, 1, 1,
                                          input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=0, input_qmax=255)
    assert not torch.equal(out, out3)


def test_reference_signature_weight_only(engine):
    """The unchanged 8-positional-argument call of QuantConv2dOp2.forward (quantconv2dop.py:59-60)."""
    c = random_conv_case(6, 2, 16, 10, 10, 24, 3, 2, 1, w_bits=4)
    T = _tensors(c)
    out = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 2, 1)
    want = oracle.quantconv2d_float_input(c["x"], c["packed"], c["des"], c["w_scale"], np.zeros_like(c["w_scale"]),
                                          c["bias"], 2, 1)
    assert np.array_equal(out.cpu().numpy(), want)


def test_fake_quantized_input_gives_same_result(engine):
    """Passing the module's dequantized activations (q+z)*s (quantconv2d.py:207-209) or the raw activations is the same."""
    c = random_conv_case(7, 2, 32, 9, 9, 32, 3, 1, 1)
    T = _tensors(c)
    kw = dict(input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=T["qmin"], input_qmax=T["qmax"])
    raw = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 1, **kw)
    q = torch.clamp(torch.round(T["x"] / T["a_scale"] - T["a_zero"]), 0, 255)
    xdq = ((q + T["a_zero"]) * T["a_scale"]).contiguous()
    fq = engine.quantconv2d_float_input(xdq, T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 1, **kw)
    assert torch.equal(raw, fq)


def test_asymmetric_weight_zero_point(engine):
    c = random_conv_case(8, 1, 16, 8, 8, 16, 3, 1, 1)
    T = _tensors(c)
    wz = torch.full_like(T["w_zero"], 2.0)
    out = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], wz, T["bias"], 1, 1,
                                         input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=0, input_qmax=255)
    qa = oracle.act_quantize(c["x"], c["a_scale"], c["a_zero"], 0, 255)
    xdq = ((qa + np.float32(c["a_zero"])) * np.float32(c["a_scale"])).astype(np.float32)
    want = oracle.quantconv2d_float_input(xdq, c["packed"], c["des"], c["w_scale"], np.full_like(c["w_scale"], 2.0),
                                          c["bias"], 1, 1)
    assert_close_1e3(out.cpu().numpy(), want)


def test_argument_errors_match_reference(engine):
    c = random_conv_case(9, 1, 16, 8, 8, 16, 3, 1, 1)
    T = _tensors(c)
    with pytest.raises(RuntimeError, match="input must be a float tensor"):
        engine.quantconv2d_float_input(T["x"].half(), T["packed"], T["des"], T["w_scale"], T["w_zero"], None, 1, 1)
    with pytest.raises(RuntimeError, match="input must be contiguous"):
        engine.quantconv2d_float_input(T["x"].transpose(2, 3), T["packed"], T["des"], T["w_scale"], T["w_zero"], None, 1, 1)
    with pytest.raises(RuntimeError, match="weight must be a CUDA tensor"):
        engine.quantconv2d_float_input(T["x"], T["packed"].cpu(), T["des"], T["w_scale"], T["w_zero"], None, 1, 1)
    with pytest.raises(RuntimeError, match="outside the quantized-operator path"):   # the float ops are not part of this engine
        engine.linear(T["x"].reshape(2, -1), T["x"].reshape(2, -1))
    with pytest.raises(RuntimeError, match="must be packed uint8 tensors"):
        engine.quantlinear(T["x"], T["des"], T["w_scale"], T["w_zero"], T["packed"], T["des"], T["w_scale"], T["w_zero"], None)


def test_non_default_stream_and_launch_counter(engine):
    c = random_conv_case(10, 2, 64, 14, 14, 64, 1, 1, 0, relu=True)
    T = _tensors(c)
    kw = dict(input_scale=T["a_scale"], input_zero=T["a_zero"], input_qmin=0, input_qmax=255)
    ref = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 0, **kw)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    engine._launch_count_reset()
    with torch.cuda.stream(s):
        out = engine.quantconv2d_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], T["bias"], 1, 0, **kw)
    s.synchronize()
    assert torch.equal(out, ref)
    assert engine._launch_count() == 2          # act-quantize + conv (weights already prepared)


@pytest.mark.parametrize("shape,k,s,p", [((3, 5, 112, 112), 3, 2, 1), ((2, 3, 17, 23), 3, 2, 1), ((1, 2, 9, 9), 2, 2, 0),
                                         ((2, 4, 31, 7), 3, 1, 1), ((1, 1, 5, 300), 5, 3, 2), ((2, 2, 224, 224), 3, 2, 1),
                                         ((2, 3, 6, 66), 3, 2, 1), ((1, 2, 4, 4), 3, 2, 1), ((2, 2, 10, 64), 3, 2, 1)])
def test_max_pool2d_matches_torch(engine, shape, k, s, p):
    """engine.max_pool2d (the stem's pooling in the packed ResNet forward) == torch.nn.functional.max_pool2d, bit for bit"""
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).cuda()
    got = engine.max_pool2d(x, k, s, p)
    want = torch.nn.functional.max_pool2d(x, k, s, p)
    assert got.shape == want.shape and torch.equal(got, want)
    x[0, 0, 2, 3] = float("nan")
    got, want = engine.max_pool2d(x, k, s, p), torch.nn.functional.max_pool2d(x, k, s, p)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.equal(got.nan_to_num(0.0), want.nan_to_num(0.0))


# ---------------------------------------------------------------------------------------------------
# quantlinear_float_input
# ---------------------------------------------------------------------------------------------------
def _linear_case(seed, B, in_f, out_f, w_bits=8, sign=True, per_tensor=False, bias=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, in_f)).astype(np.float32)
    lo, hi = (-(1 << (w_bits - 1)), (1 << (w_bits - 1)) - 1) if sign else (0, (1 << w_bits) - 1)
    qw = rng.integers(lo, hi + 1, size=(out_f, in_f)).astype(np.int64)
    packed, des = oracle.tpack(qw, w_bits, sign)
    n = 1 if per_tensor else out_f
    w_scale = (rng.random(n) * 0.02 + 0.001).astype(np.float32)
    w_zero = rng.integers(-3, 4, size=n).astype(np.float32)
    b = rng.standard_normal(out_f).astype(np.float32) if bias else None
    return dict(x=x, qw=qw, packed=packed, des=des, w_scale=w_scale, w_zero=w_zero, bias=b)


@pytest.mark.parametrize("cfg", [(5, 64, 40, 8, True, False, True), (33, 2048, 1000, 8, True, False, True),
                                 (7, 96, 17, 4, True, True, False), (3, 128, 64, 5, False, False, True),
                                 (4, 50, 9, 8, True, False, True)],
                         ids=lambda c: "B{}in{}out{}w{}{}{}{}".format(c[0], c[1], c[2], c[3], "s" if c[4] else "u",
                                                                      "t" if c[5] else "", "b" if c[6] else ""))
def test_quantlinear_weight_only_matches_reference_and_oracle(engine, cfg):
    """The reference's 6-argument call: bit-identical to the oracle restatement, and to the reference's own kernel
    (compiled unmodified into oracle/_ref) whenever in_features % 32 == 0 — the reference multiplies stale shared-memory
    entries of its last 32-wide tile otherwise (quantlinear_float_input.cu:66-68 vs :94)."""
    B, in_f, out_f, wb, sign, pt, has_b = cfg
    c = _linear_case(sum(cfg[:4]), B, in_f, out_f, wb, sign, pt, has_b)
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    args = (t(c["x"]), t(c["packed"]), t(c["des"]), t(c["w_scale"]), t(c["w_zero"]), t(c["bias"]))
    out = engine.quantlinear_float_input(*args)
    assert tuple(out.shape) == (B, out_f) and out.dtype == torch.float32
    want = oracle.quantlinear_float_input(c["x"], c["packed"], c["des"], c["w_scale"], c["w_zero"], c["bias"])
    assert np.array_equal(out.cpu().numpy(), want)
    from oracle import build_ref
    if build_ref.available() and in_f % 32 == 0:
        ref = build_ref.load().quantlinear_float_input(*args)
        assert torch.equal(out, ref)


def test_quantlinear_fused_is_the_1x1_convolution(engine):
    """With the activation quantizer's parameters the linear layer runs the integer path: same floats as the conv op on
    the [B, in, 1, 1] view, within 1e-3 of the oracle."""
    B, in_f, out_f = 300, 256, 1000
    c = _linear_case(11, B, in_f, out_f)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x = t(c["x"])
    a_scale = torch.tensor([(float(x.max()) - float(x.min())) / 255.0]).cuda()
    a_zero = torch.tensor([float(x.min())]).cuda() / a_scale
    kw = dict(input_scale=a_scale, input_zero=a_zero, input_qmin=0, input_qmax=255)
    zero = torch.zeros(out_f).cuda()
    out = engine.quantlinear_float_input(x, t(c["packed"]), t(c["des"]), t(c["w_scale"]), zero, t(c["bias"]), **kw)
    des6 = torch.tensor([8, 1, out_f, in_f, 1, 1], dtype=torch.int32).cuda()
    conv = engine.quantconv2d_float_input(x.view(B, in_f, 1, 1), t(c["packed"]), des6, t(c["w_scale"]), zero, t(c["bias"]), 1, 0, **kw)
    assert torch.equal(out, conv.view(B, out_f))
    want, _ = oracle.quantconv2d_fused(c["x"].reshape(B, in_f, 1, 1), c["packed"], np.array([8, 1, out_f, in_f, 1, 1]),
                                       c["w_scale"], c["bias"], 1, 0, float(a_scale), float(a_zero), 0.0, 255.0)
    assert_close_1e3(out.cpu().numpy(), want.reshape(B, out_f))


def test_quantlinear_argument_errors(engine):
    c = _linear_case(1, 2, 32, 8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    with pytest.raises(RuntimeError, match="2D"):
        engine.quantlinear_float_input(t(c["x"]).view(2, 32, 1), t(c["packed"]), t(c["des"]), t(c["w_scale"]), t(c["w_zero"]), None)
    with pytest.raises(RuntimeError, match="features"):
        engine.quantlinear_float_input(torch.zeros(2, 31).cuda(), t(c["packed"]), t(c["des"]), t(c["w_scale"]), t(c["w_zero"]), None)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        engine.quantlinear_float_input(torch.zeros(2, 32), t(c["packed"]), t(c["des"]), t(c["w_scale"]), t(c["w_zero"]), None)


@pytest.mark.parametrize("rng_", [(0, 255), (0, 15), (-128, 127), (-7, 7), (0, 65535)], ids=lambda r: "q{}_{}".format(*r))
@pytest.mark.parametrize("n", [1, 1000003, 4096])
def test_fake_quantize_matches_the_torch_ops(engine, rng_, n):
    """Quantizer.simulate (quantizer.py:194, :215-218) as one kernel, bit-identical to the five
    torch kernels of the reference — every range, rounding ties, values far outside the range, inf / NaN."""
    qmin, qmax = rng_
    g = torch.Generator().manual_seed(n + qmax)
    x = (torch.randn(n, generator=g) * 3).cuda()
    scale = torch.tensor([0.0371], device="cuda")
    zero = torch.tensor([-40.3 if qmin == 0 else 0.0], device="cuda")
    if n > 16:
        ks = torch.arange(qmin - 3, qmin + 9, device="cuda", dtype=torch.float32)
        x[:12] = (ks + 0.5 + zero) * scale                      # rounding ties (to even)
        x[12:16] = torch.tensor([1e30, -1e30, float("inf"), float("nan")], device="cuda")
    got = engine.fake_quantize(x, scale, zero, qmin, qmax)
    want = (torch.clamp(torch.round(x / scale - zero), qmin, qmax) + zero) * scale
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(got.nan_to_num(0.0), want.nan_to_num(0.0))
    # a 4-D, non-16-byte-aligned view takes the scalar path
    y = x[1:].contiguous()[: (n - 1) // 2 * 2].view(-1, 2) if n > 4 else x.view(1, 1)
    assert torch.equal(engine.fake_quantize(y, scale, zero, qmin, qmax).nan_to_num(0.0),
                       ((torch.clamp(torch.round(y / scale - zero), qmin, qmax) + zero) * scale).nan_to_num(0.0))


@pytest.mark.parametrize("shape", [(3, 2048, 7, 7), (2, 5, 1, 1), (1, 7, 33, 9), (4, 64, 14, 14)])
def test_avg_pool_global_matches_torch(engine, shape):
    """engine.avg_pool_global == adaptive_avg_pool2d(x, (1, 1)) within fp32 summation-order rounding; float64 mean as the
    referee (the engine's fixed-order sum must be at least as close to it as 2 ulps of the plane's absolute mean)."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).cuda()
    got = engine.avg_pool_global(x)
    want = torch.nn.functional.adaptive_avg_pool2d(x, (1, 1))
    exact = x.double().mean(dim=(2, 3), keepdim=True)
    assert got.shape == want.shape and got.dtype == torch.float32
    tol = 4 * torch.finfo(torch.float32).eps * x.abs().double().mean(dim=(2, 3), keepdim=True) + 1e-12
    assert ((got.double() - exact).abs() <= tol).all()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
