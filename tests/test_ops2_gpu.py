"""GPU: the SURVEY 8(f) "next" rows built in round 2 — quantconv2d (packed activations), quantlinear, calibration
reductions — and the boundary fixes (signed activation ranges, per-device state, inference mode, packed checkpoints)."""
import ctypes
import threading

import numpy as np
import pytest
import torch

import oracle
from gpu_util import stream
from quantize_b200 import capi
from test_conv_gpu import assert_close_1e3

pytestmark = pytest.mark.gpu


def _ref_engine():
    from oracle import build_ref
    if not build_ref.available():
        return None
    m = build_ref.load()
    return m if hasattr(m, "quantconv2d") else None


def t(a, dev="cuda"):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def packed_conv_case(seed, N, C, H, W, K, R, stride, pad, in_bits=8, in_sign=False, w_bits=8, per_channel_in=False, w_zero=0.0):
    rng = np.random.default_rng(seed)
    lo, hi = (-(1 << (in_bits - 1)), (1 << (in_bits - 1)) - 1) if in_sign else (0, (1 << in_bits) - 1)
    qa = rng.integers(lo, hi + 1, size=(N, C, H, W))
    qw = rng.integers(-(1 << (w_bits - 1)) + 1, 1 << (w_bits - 1), size=(K, C, R, R))
    ip, ides = oracle.tpack(qa, in_bits, in_sign)
    wp, wdes = oracle.tpack(qw, w_bits, True)
    n_is = C if per_channel_in else 1
    return dict(qa=qa, qw=qw, ip=ip, ides=ides, wp=wp, wdes=wdes,
                in_scale=(rng.random(n_is) * 0.05 + 0.01).astype(np.float32),
                in_zero=(rng.standard_normal(n_is) * 3).astype(np.float32),
                w_scale=(rng.random(K) * 0.02 + 0.001).astype(np.float32), w_zero=np.full(K, w_zero, np.float32),
                bias=rng.standard_normal(K).astype(np.float32), stride=stride, pad=pad,
                shape=capi.conv_shape(N, C, H, W, K, C, R, R, stride, pad, w_bits, 1), in_bits=in_bits, in_sign=in_sign)


def run_op1(engine, c, bias=True):
    return engine.quantconv2d(t(c["ip"]), t(c["ides"]), t(c["in_scale"]), t(c["in_zero"]), t(c["wp"]), t(c["wdes"]),
                              t(c["w_scale"]), t(c["w_zero"]), t(c["bias"]) if bias else None, c["stride"], c["pad"])


OP1_CASES = [
    # N, C, H, W, K, R, stride, pad, in_bits, in_sign, w_bits
    (2, 64, 14, 14, 96, 3, 1, 1, 8, False, 8),
    (3, 32, 9, 11, 40, 1, 1, 0, 8, True, 8),
    (2, 48, 15, 15, 64, 3, 2, 1, 4, False, 4),
    (1, 16, 8, 8, 16, 5, 1, 2, 5, True, 6),
    (4, 3, 20, 20, 32, 7, 2, 3, 8, False, 8),
    (2, 128, 7, 7, 256, 1, 1, 0, 7, True, 8),
]


@pytest.mark.parametrize("case", OP1_CASES)
def test_quantconv2d_integer_path(engine, case):
    """per-tensor input quantizer + symmetric weights: tensor-core integer GEMM.  int32 accumulators bit-exact vs the oracle
    (through the C-ABI); fp32 within 1e-3 of the reference's per-MAC fp32 arithmetic (oracle restatement and, when the
    compiled reference is present, its own kernel)."""
    N, C, H, W, K, R, stride, pad, ib, isg, wb = case
    c = packed_conv_case(hash(case) % 1000, N, C, H, W, K, R, stride, pad, ib, isg, wb)
    L = capi.lib()
    prepared = torch.empty(L.qb200_conv_prepared_bytes(ctypes.byref(c["shape"])), dtype=torch.uint8, device="cuda")
    capi.check(L.qb200_conv_prepare_weights(ctypes.byref(c["shape"]), t(c["wp"]).data_ptr(), prepared.data_ptr(), stream()), "prepare")
    ws = torch.empty(L.qb200_quantconv2d_packed_workspace_bytes(ctypes.byref(c["shape"])), dtype=torch.uint8, device="cuda")
    P, Q = capi.conv_out_hw(c["shape"])
    acc = torch.empty(N, K, P, Q, dtype=torch.int32, device="cuda")
    ip, isc, izr, wsc = t(c["ip"]), t(c["in_scale"]), t(c["in_zero"]), t(c["w_scale"])
    capi.check(L.qb200_quantconv2d_packed(ctypes.byref(c["shape"]), ip.data_ptr(), ib, int(isg), isc.data_ptr(), izr.data_ptr(),
                                          prepared.data_ptr(), wsc.data_ptr(), K, None, ws.data_ptr(), acc.data_ptr(),
                                          capi.OUT_ACC, stream()), "quantconv2d_packed")
    torch.cuda.synchronize()
    want_acc = oracle.quantconv2d_acc(c["ip"], c["ides"], c["wp"], c["wdes"], stride, pad)
    assert np.array_equal(acc.cpu().numpy(), want_acc)
    out = run_op1(engine, c)
    want = oracle.quantconv2d(c["ip"], c["ides"], c["in_scale"], c["in_zero"], c["wp"], c["wdes"], c["w_scale"], c["w_zero"],
                              c["bias"], stride, pad)
    assert out.dtype == torch.float32 and tuple(out.shape) == want.shape
    assert_close_1e3(out.cpu().numpy(), want)
    ref = _ref_engine()
    if ref is not None:
        r = ref.quantconv2d(t(c["ip"]), t(c["ides"]).to(torch.int32), t(c["in_scale"]), t(c["in_zero"]), t(c["wp"]),
                            t(c["wdes"]).to(torch.int32), t(c["w_scale"]), t(c["w_zero"]), t(c["bias"]), stride, pad)
        assert_close_1e3(out.cpu().numpy(), r.cpu().numpy())
        assert np.array_equal(r.cpu().numpy(), want), "oracle restatement of quantconv2d.cu differs from the compiled reference"


@pytest.mark.parametrize("per_channel_in,w_zero", [(True, 0.0), (False, 1.5), (True, -2.0)])
def test_quantconv2d_fp32_path_bit_identical(engine, per_channel_in, w_zero):
    """per-input-channel input scales / asymmetric weights: the reference's fp32 arithmetic in its order — bit-identical."""
    c = packed_conv_case(11, 2, 24, 10, 10, 20, 3, 1, 1, 6, True, 5, per_channel_in=per_channel_in, w_zero=w_zero)
    out = run_op1(engine, c, bias=per_channel_in)
    want = oracle.quantconv2d(c["ip"], c["ides"], c["in_scale"], c["in_zero"], c["wp"], c["wdes"], c["w_scale"], c["w_zero"],
                              c["bias"] if per_channel_in else None, 1, 1)
    assert np.array_equal(out.cpu().numpy(), want)
    ref = _ref_engine()
    if ref is not None:
        r = ref.quantconv2d(t(c["ip"]), t(c["ides"]).to(torch.int32), t(c["in_scale"]), t(c["in_zero"]), t(c["wp"]),
                            t(c["wdes"]).to(torch.int32), t(c["w_scale"]), t(c["w_zero"]),
                            t(c["bias"]) if per_channel_in else None, 1, 1)
        assert torch.equal(out, r)


def test_quantconv2d_through_reference_dispatch(engine):
    """quantconv2dop.py:88-91: uint8 input + uint8 weight -> QuantConv2dOp1 -> quantconv2d; the activation stream is what
    Quantizer.pack + tpack produce from the module's quantized activations."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 32, 12, 12)).astype(np.float32)
    s_a = np.float32((x.max() - x.min()) / 255.0)
    z_a = np.float32(x.min() / s_a)
    qa = oracle.act_quantize(x, float(s_a), float(z_a), 0, 255)          # q in [0, 255]; module dequant is (q + z) * s
    ip, ides = engine.tpack(t(qa), 8, False)
    qw = rng.integers(-127, 128, size=(48, 32, 3, 3))
    wp, wdes = engine.tpack(t(qw.astype(np.float32)), 8, True)
    sw = (rng.random(48) * 0.02 + 0.001).astype(np.float32)
    # Op1's convention is (q - zero) * scale: the module's zero point enters negated
    out = engine.quantconv2d(ip, ides, t(np.array([s_a])), t(np.array([-z_a])), wp, wdes, t(sw), torch.zeros(48, device="cuda"),
                             None, 1, 1)
    fused = engine.quantconv2d_float_input(t(x), wp, wdes, t(sw).reshape(-1, 1, 1, 1), torch.zeros(48, 1, 1, 1, device="cuda"),
                                           None, 1, 1, input_scale=float(s_a), input_zero=float(z_a), input_qmin=0, input_qmax=255)
    assert torch.equal(out, fused)     # same integers, same epilogue arithmetic


QL_CASES = [(5, 64, 40, 8, True, 8, True), (33, 96, 70, 4, False, 4, True), (7, 50, 33, 6, True, 5, False), (64, 256, 128, 8, False, 8, True)]


@pytest.mark.parametrize("case", QL_CASES)
def test_quantlinear(engine, case):
    B, in_f, out_f, ib, isg, wb, wsg = case
    rng = np.random.default_rng(B * 7 + in_f)
    lo_i, hi_i = (-(1 << (ib - 1)), (1 << (ib - 1)) - 1) if isg else (0, (1 << ib) - 1)
    lo_w, hi_w = (-(1 << (wb - 1)), (1 << (wb - 1)) - 1) if wsg else (0, (1 << wb) - 1)
    qi, qw = rng.integers(lo_i, hi_i + 1, size=(B, in_f)), rng.integers(lo_w, hi_w + 1, size=(out_f, in_f))
    ip, ides = oracle.tpack(qi, ib, isg)
    wp, wdes = oracle.tpack(qw, wb, wsg)
    isc, izr = (rng.random(B) * 0.05 + 0.01).astype(np.float32), (rng.standard_normal(B) * 2).astype(np.float32)
    wsc, wzr = (rng.random(out_f) * 0.02 + 0.001).astype(np.float32), (rng.standard_normal(out_f)).astype(np.float32)
    bias = rng.standard_normal(out_f).astype(np.float32)
    out = engine.quantlinear(t(ip), t(ides), t(isc), t(izr), t(wp), t(wdes), t(wsc), t(wzr), t(bias))
    want = oracle.quantlinear(ip, ides, isc, izr, wp, wdes, wsc, wzr, bias)
    assert np.array_equal(out.cpu().numpy(), want)
    # 0-d scales / zeros are expanded like the reference's wrapper does (quantlinear.cu:275-289); bias None -> zeros
    out0 = engine.quantlinear(t(ip), t(ides), torch.tensor(0.03, device="cuda"), torch.tensor(0.5, device="cuda"), t(wp), t(wdes),
                              torch.tensor(0.01, device="cuda"), torch.tensor(0.0, device="cuda"), None)
    want0 = oracle.quantlinear(ip, ides, np.float32(0.03), np.float32(0.5), wp, wdes, np.float32(0.01), np.float32(0.0), None)
    assert np.array_equal(out0.cpu().numpy(), want0)
    ref = _ref_engine()
    if ref is not None and in_f % 32 == 0:      # the reference's tiles keep stale entries for ragged input sizes
        r = ref.quantlinear(t(ip), t(ides).to(torch.int32), t(isc), t(izr), t(wp), t(wdes).to(torch.int32), t(wsc), t(wzr), t(bias))
        assert torch.equal(out, r)


@pytest.mark.parametrize("shape,gran,flag", [((4, 16, 9, 9), 0, "activation"), ((4, 16, 9, 9), 1, "activation"),
                                              ((32, 16, 3, 3), 1, "weight"), ((3, 5, 7), 1, "activation"),
                                              ((1 << 20) + 3, 0, "weight"), ((64, 100001), 1, "weight")])
@pytest.mark.parametrize("symmetric", [False, True])
def test_minmax_matches_torch(engine, shape, gran, flag, symmetric):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(shape, generator=g, device="cuda") * 3
    fl = 1 if flag == "activation" else 0
    lo, hi = engine.minmax(x, gran, fl, symmetric)
    rows = x.reshape(1, -1) if gran == 0 else (x.transpose(0, 1).flatten(1) if fl else x.flatten(1))
    if symmetric:
        wlo, whi = torch.zeros(rows.shape[0], device="cuda"), rows.abs().max(dim=1)[0]
    else:
        wlo, whi = rows.min(dim=1)[0], rows.max(dim=1)[0]
    if gran == 0:
        wlo, whi = wlo[0], whi[0]
    assert lo.shape == wlo.shape and torch.equal(lo, wlo) and torch.equal(hi, whi)
    olo, ohi = oracle.minmax(x.cpu().numpy(), gran, flag, symmetric)
    assert np.array_equal(lo.cpu().numpy(), olo) and np.array_equal(hi.cpu().numpy(), ohi)
    # fused state updates: running min/max (MinMax.update) and moving average (MAMinMax.update), bit-identical to torch
    y = torch.randn(shape, generator=g, device="cuda") * 2 + 0.5
    rmin, rmax = lo.clone(), hi.clone()
    nlo, nhi = engine.minmax(y, gran, fl, symmetric, 1, 0.0, rmin, rmax)
    ylo, yhi = engine.minmax(y, gran, fl, symmetric)
    assert torch.equal(nlo, torch.min(lo, ylo)) and torch.equal(nhi, torch.max(hi, yhi))
    assert torch.equal(rmin, nlo) and torch.equal(rmax, nhi)
    rmin, rmax = lo.clone(), hi.clone()
    m = 0.1
    nlo, nhi = engine.minmax(y, gran, fl, symmetric, 2, m, rmin, rmax)
    assert torch.equal(nlo, m * ylo + (1 - m) * lo) and torch.equal(nhi, m * yhi + (1 - m) * hi)


def test_minmax_nan_and_kthvalue(engine):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(6, 8, 5, 5, generator=g, device="cuda")
    x[2, 3, 1, 1] = float("nan")
    lo, hi = engine.minmax(x, 1, 1, False)
    assert torch.isnan(lo[3]) and torch.isnan(hi[3]) and not torch.isnan(lo[[0, 1, 2, 4, 5, 6, 7]]).any()
    lo, hi = engine.minmax(x, 0, 0, True)
    assert torch.isnan(hi) and float(lo) == 0.0
    # kthvalue: exact, NaN last
    for shape, gran, fl in [((4, 16, 9, 9), 0, 1), ((4, 16, 9, 9), 1, 1), ((32, 147), 1, 0), ((300000,), 0, 0)]:
        y = torch.randn(shape, generator=g, device="cuda")
        y.view(-1)[::97] = y.view(-1)[5]          # ties
        rows = y.reshape(1, -1) if gran == 0 else (y.transpose(0, 1).flatten(1) if fl else y.flatten(1))
        n = rows.shape[1]
        for k in (1, 2, n // 3, n - 1, n):
            got = engine.kthvalue(y, k, gran, fl, False)
            want = rows.kthvalue(k, dim=1)[0]
            assert torch.equal(got, want[0] if gran == 0 else want), (shape, gran, k)
            got = engine.kthvalue(y, k, gran, fl, True)
            want = rows.abs().kthvalue(k, dim=1)[0]
            assert torch.equal(got, want[0] if gran == 0 else want), (shape, gran, k, "abs")
    z = torch.tensor([3.0, float("nan"), -1.0, 2.0], device="cuda")
    assert float(engine.kthvalue(z, 3, 0, 0, False)) == 3.0 and torch.isnan(engine.kthvalue(z, 4, 0, 0, False))
    with pytest.raises(RuntimeError, match="out of range"):
        engine.kthvalue(z, 5, 0, 0, False)


@pytest.mark.parametrize("cfg", [dict(symmetric=False, granularity="layer", range={"name": "maminmax", "momentum": 0.1}),
                                 dict(symmetric=True, granularity="channel", range={"name": "minmax", "percentile": 0.0}),
                                 dict(symmetric=False, granularity="layer", range={"name": "minmax", "percentile": 0.01}),
                                 dict(symmetric=True, granularity="layer", range={"name": "minmax", "percentile": 0.001})])
def test_host_range_estimators_engine_equals_torch(cfg):
    """host.MinMax / MAMinMax on CUDA tensors: engine reductions == the reference's torch expressions, bit for bit,
    over several calibration batches (state updates included)."""
    from quantize_b200 import host
    g = torch.Generator(device="cuda").manual_seed(4)
    flag = "weight" if cfg["granularity"] == "channel" else "activation"
    res = {}
    for use in (True, False):
        host.MinMax.use_engine = use
        try:
            q = host.Quantizer(n_bits=8, signed=True, flag=flag, n_channels=16, dim=4, **cfg)
            g.manual_seed(4)
            for i in range(3):
                q.calibrate(torch.randn(8, 16, 7, 7, generator=g, device="cuda") * (i + 1))
            res[use] = (q.scale.detach().clone(), q.zero.detach().clone(), q.qmin.clone(), q.qmax.clone())
        finally:
            host.MinMax.use_engine = True
    for a, b in zip(res[True], res[False]):
        assert torch.equal(a.cuda().reshape(-1), b.cuda().reshape(-1))


def test_signed_activation_range_takes_the_fp32_path(engine):
    """a symmetric signed activation quantizer (qmin = -128, minmax.py:124-127) does not fit the unsigned byte
    operand; the op must not wrap negative values — it fake-quantizes and runs the fp32 kernel."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 16, 9, 9)).astype(np.float32)
    qw = rng.integers(-127, 128, size=(24, 16, 3, 3))
    wp, wdes = oracle.tpack(qw, 8, True)
    sw = (rng.random(24) * 0.02 + 0.001).astype(np.float32)
    s_a = np.float32(np.abs(x).max() / 127.0)
    out = engine.quantconv2d_float_input(t(x), t(wp), t(wdes), t(sw).reshape(-1, 1, 1, 1), torch.zeros(24, 1, 1, 1, device="cuda"),
                                         None, 1, 1, input_scale=float(s_a), input_zero=0.0,
                                         input_qmin=torch.tensor(-128, device="cuda"), input_qmax=torch.tensor(127, device="cuda"))
    q = np.clip(np.rint(x / s_a - np.float32(0)), -128, 127).astype(np.float32)
    want = oracle.quantconv2d_float_input((q * s_a).astype(np.float32), wp, wdes, sw, np.zeros(24, np.float32), None, 1, 1)
    assert np.array_equal(out.cpu().numpy(), want)
    # and the chain op refuses such a quantizer instead of wrapping
    layer = (t(wp), t(wdes), t(sw).reshape(-1, 1, 1, 1), torch.zeros(24, 1, 1, 1, device="cuda"), None, 1, 1, float(s_a), 0.0, -128, 127, True)
    with pytest.raises(RuntimeError, match=r"\[0, 255\]"):
        engine.quantconv2d_chain(t(x), [layer])


def test_inference_mode_and_threads(engine):
    """Tensor::_version() throws for inference tensors; the GIL is released around the launches."""
    rng = np.random.default_rng(6)
    x = rng.standard_normal((2, 64, 14, 14)).astype(np.float32)
    qw = rng.integers(-127, 128, size=(64, 64, 1, 1))
    sw = (rng.random(64) * 0.02 + 0.001).astype(np.float32)
    kw = dict(input_scale=0.02, input_zero=-100.0, input_qmin=0, input_qmax=255)
    wp, wdes = engine.tpack(t(qw.astype(np.float32)), 8, True)
    base = engine.quantconv2d_float_input(t(x), wp, wdes, t(sw).reshape(-1, 1, 1, 1), torch.zeros(64, 1, 1, 1, device="cuda"), None, 1, 0, **kw)
    with torch.inference_mode():
        wp2, wdes2 = engine.tpack(t(qw.astype(np.float32)), 8, True)
        out = engine.quantconv2d_float_input(t(x), wp2, wdes2, t(sw).reshape(-1, 1, 1, 1), torch.zeros(64, 1, 1, 1, device="cuda"),
                                             None, 1, 0, **kw)
        assert torch.equal(out, base)
    results, errs = [None] * 4, []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(20):
                    results[i] = engine.quantconv2d_float_input(t(x), wp, wdes, t(sw).reshape(-1, 1, 1, 1),
                                                                torch.zeros(64, 1, 1, 1, device="cuda"), None, 1, 0, **kw)
            s.synchronize()
        except Exception as e:   # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [a.start() for a in th]
    [a.join() for a in th]
    assert not errs, errs
    assert all(torch.equal(r, base) for r in results)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_thread_two_devices(engine):
    """the >48 KB shared-memory attribute is per device: cuda:0 then cuda:1 from one thread."""
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 64, 28, 28)).astype(np.float32)
    qw = rng.integers(-127, 128, size=(64, 64, 3, 3))
    sw = (rng.random(64) * 0.02 + 0.001).astype(np.float32)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        wp, wdes = engine.tpack(t(qw.astype(np.float32), dev), 8, True)
        o = engine.quantconv2d_float_input(t(x, dev), wp, wdes, t(sw, dev).reshape(-1, 1, 1, 1), torch.zeros(64, 1, 1, 1, device=dev),
                                           None, 1, 1, input_scale=0.02, input_zero=-100.0, input_qmin=0, input_qmax=255)
        p = engine.max_pool2d(o, 3, 2, 1)
        torch.cuda.synchronize(dev)
        outs.append((o.cpu(), p.cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[1][1])


def test_packed_state_dict_loads_into_a_fresh_model():
    """reference quantconv2d.py:212-235 / quantlinear.py:170-186: a packed checkpoint (it holds w_des) loads into a freshly
    reconstructed, un-packed model; this mirror keeps the packed byte stream."""
    from quantize_b200 import models
    net = models.build_packed("resnet20", 8, 8, calib_batch=8, device="cuda", seed=0)
    x = models.synthetic_batch("resnet20", 4, 9, "cuda")
    with torch.no_grad():
        want = net(x)
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    fresh = models.build_quantized("resnet20", 8, 8, seed=123).to("cuda")     # different random weights, not calibrated
    missing = fresh.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    convs = [m for m in fresh.modules() if hasattr(m, "w_des")]
    assert convs and all(m.packed and m.weight.dtype == torch.uint8 and m.weight.is_cuda for m in convs)
    from quantize_b200 import host
    host.set_mode(fresh, calibrating=False, quantized=True)
    with torch.no_grad():
        got = fresh(x)
    assert torch.equal(got, want)
