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
    with pytest.raises(RuntimeError, match="outside the hot path"):
        engine.quantlinear_float_input(T["x"], T["packed"], T["des"], T["w_scale"], T["w_zero"], None)


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
                                         ((2, 4, 31, 7), 3, 1, 1), ((1, 1, 5, 300), 5, 3, 2), ((2, 2, 224, 224), 3, 2, 1)])
def test_max_pool2d_matches_torch(engine, shape, k, s, p):
    """engine.max_pool2d (the stem's pooling in the packed ResNet forward) == torch.nn.functional.max_pool2d, bit for bit"""
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(sum(shape))).cuda()
    got = engine.max_pool2d(x, k, s, p)
    want = torch.nn.functional.max_pool2d(x, k, s, p)
    assert got.shape == want.shape and torch.equal(got, want)
    x[0, 0, 2, 3] = float("nan")
    got, want = engine.max_pool2d(x, k, s, p), torch.nn.functional.max_pool2d(x, k, s, p)
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.equal(got.nan_to_num(0.0), want.nan_to_num(0.0))
