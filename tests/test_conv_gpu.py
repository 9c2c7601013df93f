"""GPU parity: the fused quantized conv (C-ABI) vs the oracle, the reference-generated fixtures and the reference kernel.

Bar (BASELINE.json north_star): integer results bit-exact (quantized activations, int32 accumulators);
dequantized fp32 outputs within 1e-3 relative.
"""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from conftest import conv_fixture_names, load_conv_fixture
from gpu_util import (ActQ, c_act_quantize, c_conv_fused, c_prepare, c_weightonly, random_conv_case)
from quantize_b200 import capi

pytestmark = pytest.mark.gpu
ALGOS = {"direct": capi.ALGO_DIRECT, "umma": capi.ALGO_UMMA_FUSED_QUANT, "umma2k": capi.ALGO_UMMA_TWO_KERNELS}


def assert_close_1e3(got, ref):
    """|got - ref| <= 1e-3*|ref| + 1e-4*max|ref|  (north star: 1e-3 relative; the absolute floor covers results
    that cancel to ~0, where 'relative' has no meaning for a sum of thousands of products)."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    tol = 1e-3 * np.abs(ref) + 1e-4 * np.abs(ref).max()
    bad = np.abs(got - ref) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} outside tolerance, worst {np.abs(got - ref).max():.4e}"


def run_case(c, algo, check_float=True):
    """c: dict from random_conv_case / fixture adapter.  Returns (acc, out)."""
    dev = "cuda"
    x = torch.from_numpy(c["x"]).to(dev)
    aq = ActQ(c["a_scale"], c["a_zero"], c["qmin"], c["qmax"])
    prepared = c_prepare(c["shape"], torch.from_numpy(c["packed"]).to(dev))
    w_scale = torch.from_numpy(c["w_scale"]).to(dev)
    bias = torch.from_numpy(c["bias"]).to(dev) if c["bias"] is not None else None
    acc = c_conv_fused(c["shape"], x, prepared, w_scale, bias, aq, capi.OUT_ACC, algo)
    out = c_conv_fused(c["shape"], x, prepared, w_scale, bias, aq, capi.OUT_F32, algo) if check_float else None
    return acc, out


def oracle_case(c):
    qa = oracle.act_quantize(c["x"], c["a_scale"], c["a_zero"], c["qmin"], c["qmax"]).astype(np.uint8)
    qw = oracle.tunpack(c["packed"], c["des"])
    acc, wsum = oracle.conv_acc(qa, qw, c["stride"], c["pad"])
    out = oracle.dequant(acc, wsum, c["a_scale"], c["a_zero"], c["w_scale"], c["bias"])
    return qa, acc, out


# ---------------------------------------------------------------------------------------------------
# activation quantizer
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 3, 9, 7), (1, 64, 14, 14), (3, 160, 5, 5), (2, 256, 7, 9), (1, 1, 1, 1),
                                   (5, 33, 13, 11)])
def test_act_quantize_bit_exact(shape):
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    if x.size >= 4:
        x.reshape(-1)[:4] = [0.0, -0.0, 1e-30, -1e30]
    for (s, z, lo, hi) in ((0.0371, -107.9, 0, 255), (0.5, 0.0, 0, 15), (1.0 / 3.0, -8.164, 0, 15), (0.013, -129.283, 0, 255)):
        aq = ActQ(s, z, lo, hi)
        q = c_act_quantize(torch.from_numpy(x).cuda(), aq).cpu().numpy()
        want = oracle.act_quantize(x, s, z, lo, hi).astype(np.uint8)
        N, C, H, W = shape
        assert np.array_equal(q[..., :C], want.transpose(0, 2, 3, 1))
        assert not q[..., C:].any()          # padded channels are zero


def test_act_quantize_rounding_boundaries_exact():
    """The kernel replaces the IEEE divide by a reciprocal + two exact-residual corrections; the integers must still
    be those of fl(fl(x/s) - z) rounded half-even.  Probe every rounding boundary k + 0.5 of every quantization
    step with the inputs on it and a few ulps around it, for awkward scales / zero points, plus special values."""
    rng = np.random.default_rng(2024)
    cases = [(0.0371, -107.9, 0, 255), (1.0 / 3.0, -8.164, 0, 15), (0.013, -129.283, 0, 255), (1.0, 0.0, 0, 255),
             (np.float32(1.9999999), 0.5, 0, 255), (np.float32(1.1754944e-3), -0.25, 0, 127), (3.0e-5, 17.3, 3, 200)]
    cases += [(float(np.float32(rng.uniform(1e-4, 10))), float(np.float32(rng.uniform(-200, 50))), 0, 255) for _ in range(6)]
    for (s, z, lo, hi) in cases:
        s32, z32 = np.float32(s), np.float32(z)
        k = np.arange(lo - 3, hi + 4, dtype=np.float64)
        xs = []
        for frac in (0.5, 0.0, 0.25, 0.499999, 0.500001):
            base = ((k + frac + np.float64(z32)) * np.float64(s32)).astype(np.float32)
            for ulps in range(-4, 5):
                v = base.copy()
                for _ in range(abs(ulps)):
                    v = np.nextafter(v, np.float32(np.inf if ulps > 0 else -np.inf))
                xs.append(v)
        specials = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 3.4e38, -3.4e38, 1e-40, -1e-40, 1e-30, 123456.7],
                            dtype=np.float32)
        x = np.concatenate(xs + [specials, (rng.standard_normal(1 << 16) * 50).astype(np.float32)])
        pad = (-x.size) % 64
        x = np.concatenate([x, np.zeros(pad, np.float32)]).reshape(1, 16, -1, 4)      # H*W % 4 == 0: vector kernel
        aq = ActQ(s, z, lo, hi)
        want = oracle.act_quantize(x, float(s32), float(z32), lo, hi)
        finite = ~np.isnan(x)                                  # NaN has no defined image (the kernel maps it to qmin)
        want = np.where(finite, want, lo).astype(np.uint8)
        got = c_act_quantize(torch.from_numpy(x).cuda(), aq).cpu().numpy()[..., :16].transpose(0, 3, 1, 2)
        bad = got != want
        assert not bad.any(), (s, z, x[bad][:5], got[bad][:5], want[bad][:5])
        # the generic (non-vector) kernel through an odd spatial size
        xo = x.reshape(-1)[: 16 * 49 * 3].reshape(3, 16, 7, 7).copy()
        wo = oracle.act_quantize(xo, float(s32), float(z32), lo, hi)
        wo = np.where(~np.isnan(xo), wo, lo).astype(np.uint8)
        go = c_act_quantize(torch.from_numpy(xo).cuda(), aq).cpu().numpy()[..., :16].transpose(0, 3, 1, 2)
        assert np.array_equal(go, wo)


def test_act_quantize_ties_round_half_even():
    # x/s - z lands exactly on .5: torch.round and rintf both round half to even (quantizer.py:31)
    x = (np.arange(0, 64, dtype=np.float32) + 0.5).reshape(1, 1, 8, 8)
    aq = ActQ(1.0, 0.0, 0, 255)
    q = c_act_quantize(torch.from_numpy(x).cuda(), aq).cpu().numpy()[..., 0].reshape(-1)
    assert np.array_equal(q, np.round(x.reshape(-1)).astype(np.uint8))
    assert q[0] == 0 and q[1] == 2 and q[2] == 2


# ---------------------------------------------------------------------------------------------------
# fixtures generated by the reference's own modules (tests/golden/make_golden.py)
# ---------------------------------------------------------------------------------------------------
def fixture_case(name):
    f = load_conv_fixture(name)
    des = [int(v) for v in f["w_des"]]
    N, C, H, W = f["x"].shape
    shape = capi.conv_shape(N, C, H, W, des[2], des[3], des[4], des[5], f["stride"], f["pad"], des[0], des[1])
    return dict(x=f["x"], packed=f["w_packed"], des=f["w_des"], a_scale=float(f["a_scale"][0]), a_zero=float(f["a_zero"][0]),
                qmin=float(f["qmin"][0]), qmax=float(f["qmax"][0]), w_scale=f["w_scale"], bias=f["bias"], shape=shape,
                stride=f["stride"], pad=f["pad"], groups=f["groups"]), f


@pytest.mark.parametrize("algo", ["direct", "umma", "umma2k"])
@pytest.mark.parametrize("name", conv_fixture_names())
def test_reference_fixture(name, algo):
    c, f = fixture_case(name)
    if algo != "direct" and c["groups"] != 1:
        pytest.skip("tensor-core kernel is groups == 1; depthwise runs on the CUDA-core kernel")
    acc, out = run_case(c, ALGOS[algo])
    qa, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(qa, f["q_x"])
    assert np.array_equal(acc.cpu().numpy(), acc_ref)                 # int32 accumulators: bit-exact
    assert_close_1e3(out.cpu().numpy(), out_ref)
    assert_close_1e3(out.cpu().numpy(), f["out_packed"])              # the reference module's packed forward
    assert_close_1e3(out.cpu().numpy(), f["out_fake"])                # and its fake-quant forward


# ---------------------------------------------------------------------------------------------------
# seeded synthetic layers vs the oracle (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------
SYNTH = [
    # N, C, H, W, K, R, stride, pad, groups, w_bits, a_bits, w_sign, relu, per_tensor
    (2, 64, 14, 14, 64, 3, 1, 1, 1, 8, 8, True, False, False),     # ResNet 3x3 (KC=64)
    (1, 64, 56, 56, 64, 1, 1, 0, 1, 8, 8, True, True, False),      # ResNet-50 layer1 1x1
    (2, 128, 28, 28, 128, 3, 2, 1, 1, 8, 8, True, False, False),   # strided 3x3 (KC=128)
    (2, 256, 14, 14, 512, 1, 2, 0, 1, 8, 8, True, True, False),    # 1x1 stride-2 downsample, two N tiles
    (3, 3, 33, 33, 64, 7, 2, 3, 1, 8, 8, True, False, False),      # stem: C=3 padded to 32 (KC=32)
    (2, 96, 9, 9, 24, 1, 1, 0, 1, 4, 8, True, False, False),       # MobileNetV2 pointwise, W4, K<64, KC=32
    (1, 512, 7, 7, 512, 3, 1, 1, 1, 8, 8, True, True, False),      # deepest 3x3 (Kgemm 4608), M=49 (< one tile)
    (2, 32, 10, 12, 40, 3, 1, 1, 1, 4, 4, True, False, True),      # W4A4, per-tensor weight scale, ragged K
    (2, 32, 8, 8, 32, 3, 1, 1, 32, 8, 8, True, False, False),      # depthwise
    (2, 32, 8, 8, 64, 3, 2, 1, 4, 8, 8, True, False, False),       # grouped
    (1, 64, 12, 12, 64, 3, 1, 1, 1, 8, 8, False, False, False),    # unsigned weights (u8 B operand)
    (2, 64, 9, 9, 128, 5, 1, 2, 1, 5, 7, True, False, False),      # 5x5, odd bit widths
    (1, 16, 6, 6, 8, 3, 1, 0, 1, 8, 8, True, False, False),        # no padding
    (4, 2048, 7, 7, 64, 1, 1, 0, 1, 8, 8, True, True, False),      # long K (16 k-blocks of 128)
    (3, 256, 14, 14, 64, 1, 1, 0, 1, 8, 8, True, False, False),    # 1x1 single-kernel path, KC=128, z_a != 0, ragged M
    (2, 512, 28, 28, 128, 1, 1, 0, 1, 8, 8, True, True, False),    # 1x1 single-kernel path, 4 k-blocks, both groups
    (5, 64, 6, 6, 96, 1, 1, 0, 1, 4, 4, True, False, True),        # 1x1 single-kernel path, KC=64, A4, ragged K
    (2, 192, 8, 8, 32, 1, 1, 0, 1, 8, 8, True, False, False),      # 1x1, C % 128 == 64 (KC=64, 3 k-blocks)
]


@pytest.mark.parametrize("algo", ["direct", "umma", "umma2k"])
@pytest.mark.parametrize("cfg", SYNTH, ids=lambda c: "N{}C{}H{}W{}K{}R{}s{}p{}g{}w{}a{}{}{}{}".format(*c[:11], "s" if c[11] else "u", "r" if c[12] else "", "t" if c[13] else ""))
def test_synthetic_vs_oracle(cfg, algo):
    N, C, H, W, K, R, stride, pad, groups, wb, ab, wsign, relu, pt = cfg
    if algo != "direct" and groups != 1:
        pytest.skip("tensor-core kernel is groups == 1")
    c = random_conv_case(hash(cfg) % (2 ** 31), N, C, H, W, K, R, stride, pad, groups, wb, ab, wsign, relu, pt)
    acc, out = run_case(c, ALGOS[algo])
    if not wsign:
        # the oracle's integer conv takes int8 weights; unsigned weights are checked through an exact float64 conv
        qa = oracle.act_quantize(c["x"], c["a_scale"], c["a_zero"], c["qmin"], c["qmax"])
        ref = torch.nn.functional.conv2d(torch.from_numpy(qa).double(), torch.from_numpy(c["qw"]).double(), None, stride, pad)
        assert np.array_equal(acc.cpu().numpy(), ref.numpy().astype(np.int32))
        return
    _, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(acc.cpu().numpy(), acc_ref)
    assert_close_1e3(out.cpu().numpy(), out_ref)


def test_auto_dispatch_matches():
    c = random_conv_case(3, 2, 64, 14, 14, 64, 3, 1, 1)
    a0, o0 = run_case(c, capi.ALGO_AUTO)
    a1, o1 = run_case(c, capi.ALGO_DIRECT)
    assert torch.equal(a0, a1) and torch.equal(o0, o1)        # same epilogue arithmetic -> identical floats


DEPTHWISE = [
    (2, 32, 17, 19, 32, 3, 1, 1, 32, 8, 8),     # MobileNetV2-style 3x3, ragged plane
    (3, 24, 15, 15, 24, 3, 2, 1, 24, 4, 8),     # stride 2, W4A8
    (1, 16, 9, 9, 32, 3, 1, 1, 16, 8, 8),       # channel multiplier 2
    (2, 8, 11, 11, 8, 5, 1, 2, 8, 8, 4),        # 5x5 (run-time loops), A4
    (2, 96, 112, 112, 96, 3, 2, 1, 96, 4, 8),   # MobileNetV2 block 2 depthwise at its real spatial size
    (2, 960, 7, 7, 960, 3, 1, 1, 960, 4, 8),    # the 7x7 tail
    # streaming kernel (3x3, pad 1, one filter per channel): lane-group widths 8 / 16 / 32, several column groups, odd sizes
    (2, 5, 14, 14, 5, 3, 1, 1, 5, 8, 8),        # stride 1, 16-lane groups, two planes per warp, odd plane count
    (2, 6, 7, 7, 6, 3, 2, 1, 6, 8, 8),          # stride 2 on odd 7x7 planes: W odd -> the band kernel
    (3, 7, 14, 14, 7, 3, 2, 1, 7, 4, 8),        # stride 2, 8-lane groups (Q = 7), four planes per warp
    (1, 3, 56, 70, 3, 3, 1, 1, 3, 8, 8),        # stride 1, three column groups of 30 (70 columns)
    (2, 4, 33, 64, 4, 3, 2, 1, 4, 8, 8),        # stride 2, odd H (last output row ends on the last input row), Q = 32: two groups
    (1, 2, 112, 112, 2, 3, 1, 1, 2, 8, 4),      # stride 1 at 112x112, A4
    (2, 9, 28, 28, 9, 3, 2, 1, 9, 8, 8),        # stride 2, Q = 14: 16-lane groups
]


@pytest.mark.parametrize("cfg", DEPTHWISE, ids=lambda c: "N{}C{}H{}W{}K{}R{}s{}p{}g{}w{}a{}".format(*c))
def test_depthwise_fused_kernel(cfg):
    """groups == C layers run as ONE fused kernel (quantize + stencil + dequant, conv_dw.cu) by default: accumulators
    bit-exact vs the oracle, fp32 within 1e-3, and identical floats to the two-kernel CUDA-core path."""
    N, C, H, W, K, R, stride, pad, groups, wb, ab = cfg
    c = random_conv_case(sum(cfg), N, C, H, W, K, R, stride, pad, groups, wb, ab)
    assert capi.lib().qb200_conv_is_single_kernel(ctypes.byref(c["shape"]), None) == 1
    acc, out = run_case(c, capi.ALGO_AUTO)
    _, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(acc.cpu().numpy(), acc_ref)
    assert_close_1e3(out.cpu().numpy(), out_ref)
    acc_d, out_d = run_case(c, capi.ALGO_DIRECT)
    assert torch.equal(acc, acc_d) and torch.equal(out, out_d)


# ---------------------------------------------------------------------------------------------------
# full BASELINE sizes: exact check against an independent float64 convolution on the GPU + shard property
# ---------------------------------------------------------------------------------------------------
FULL = [
    (256, 64, 56, 56, 64, 3, 1, 1),      # ResNet-50 layer1 conv2, batch 256
    (256, 256, 56, 56, 128, 1, 1, 0),    # layer2.0 conv1
    (128, 3, 224, 224, 64, 7, 2, 3),     # ResNet-18 stem, batch 128
    (256, 512, 7, 7, 2048, 1, 1, 0),     # layer4 conv3
    (256, 1024, 14, 14, 2048, 1, 2, 0),  # layer4 downsample
]


@pytest.mark.parametrize("cfg", FULL, ids=lambda c: "N{}C{}H{}K{}R{}s{}".format(c[0], c[1], c[2], c[4], c[5], c[6]))
def test_full_size_exact_and_shardable(cfg):
    N, C, H, W, K, R, stride, pad = cfg
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(N, C, H, W, generator=g, device="cuda")
    qw = torch.randint(-127, 128, (K, C, R, R), generator=g, device="cuda", dtype=torch.int32)
    import oracle as _o
    packed, des = _o.tpack(qw.cpu().numpy(), 8, True)
    shape = capi.conv_shape(N, C, H, W, K, C, R, R, stride, pad, 8, 1)
    xmin, xmax = float(x.min()), float(x.max())
    s_a = np.float32((xmax - xmin) / 255.0)
    aq = ActQ(s_a, np.float32(np.float32(xmin) / s_a), 0, 255)
    prepared = c_prepare(shape, torch.from_numpy(packed).cuda())
    w_scale = torch.full((K,), 0.01, device="cuda")
    acc = c_conv_fused(shape, x, prepared, w_scale, None, aq, capi.OUT_ACC)
    # independent exact reference: quantize with torch ops (same formula), convolve a few images in float64
    qa = torch.clamp(torch.round(x / aq.t[0] - aq.t[1]), 0, 255)
    for n0 in (0, N // 2, N - 1):
        ref = torch.nn.functional.conv2d(qa[n0:n0 + 1].double(), qw.double(), None, stride, pad)
        assert torch.equal(acc[n0:n0 + 1].double(), ref)
    # checksum over the whole batch: sum_k,p,q acc == conv with the k-summed filter (linearity), in float64
    wsum = qw.sum(0, keepdim=True).double()
    ref_sum = torch.nn.functional.conv2d(qa.double(), wsum, None, stride, pad)
    assert torch.equal(acc.sum(1, keepdim=True, dtype=torch.int64).double(), ref_sum)
    # batch sharding: a shard's result equals the matching rows of the full-batch result
    sh = N // 8
    shape_s = capi.conv_shape(sh, C, H, W, K, C, R, R, stride, pad, 8, 1)
    acc_s = c_conv_fused(shape_s, x[3 * sh:4 * sh].contiguous(), prepared, w_scale, None, aq, capi.OUT_ACC)
    assert torch.equal(acc_s, acc[3 * sh:4 * sh])


# ---------------------------------------------------------------------------------------------------
# weight-only semantic of the 8-argument op
# ---------------------------------------------------------------------------------------------------
def test_weightonly_vs_oracle_and_reference_kernel():
    from oracle import build_ref
    ref = build_ref.load() if build_ref.available() else None
    rng = np.random.default_rng(21)
    for (N, C, H, W, K, R, stride, pad, nb, sign, pt) in ((2, 16, 9, 9, 12, 3, 1, 1, 8, True, False),
                                                          (1, 8, 11, 7, 6, 3, 2, 1, 4, True, True),
                                                          (2, 6, 8, 8, 10, 1, 1, 0, 5, False, False),
                                                          (1, 3, 20, 20, 8, 7, 2, 3, 8, True, False)):
        x = rng.standard_normal((N, C, H, W)).astype(np.float32)
        lo, hi = (-(1 << (nb - 1)), (1 << (nb - 1)) - 1) if sign else (0, (1 << nb) - 1)
        qw = rng.integers(lo, hi + 1, size=(K, C, R, R))
        packed, des = oracle.tpack(qw, nb, sign)
        ns = 1 if pt else K
        scale = (rng.random(ns) * 0.1 + 0.01).astype(np.float32)
        zero = rng.integers(-3, 4, size=ns).astype(np.float32)   # weight_zero != 0: the kernel's (q - zero) convention
        bias = rng.standard_normal(K).astype(np.float32)
        shape = capi.conv_shape(N, C, H, W, K, C, R, R, stride, pad, nb, sign)
        t = lambda a: torch.from_numpy(a).cuda()
        got = c_weightonly(shape, t(x), t(packed), t(scale), t(zero), t(bias)).cpu().numpy()
        want = oracle.quantconv2d_float_input(x, packed, des, scale, zero, bias, stride, pad)
        assert np.array_equal(got, want), float(np.abs(got - want).max())     # same order, same FMA: bit-exact
        if ref is not None:
            r = ref.quantconv2d_float_input(t(x), t(packed), t(des), t(scale), t(zero), t(bias), stride, pad)
            assert np.array_equal(got, r.cpu().numpy())


def test_fused_vs_reference_kernel_on_dequantized_input():
    """Op-level oracle on the GPU: the unmodified reference kernel fed the fake-quantized activations
    (what the reference's intended wiring passes it, quantconv2d.py:202-210)."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    ref = build_ref.load()
    c = random_conv_case(99, 2, 32, 12, 12, 48, 3, 1, 1)
    acc, out = run_case(c, capi.ALGO_AUTO)
    t = lambda a: torch.from_numpy(a).cuda()
    x = t(c["x"])
    s, z = torch.tensor(c["a_scale"]).cuda(), torch.tensor(c["a_zero"]).cuda()
    q = torch.clamp(torch.round(x / s - z), c["qmin"], c["qmax"])
    xdq = ((q + z) * s).contiguous()
    K = c["shape"].K
    r = ref.quantconv2d_float_input(xdq, t(c["packed"]), t(c["des"]), t(c["w_scale"]), torch.zeros(K).cuda(),
                                    t(c["bias"]), c["stride"], c["pad"])
    assert_close_1e3(out.cpu().numpy(), r.cpu().numpy())
    # integer accumulators through the reference kernel: scale 1, zero 0, integer inputs (exact in fp32 below 2^24)
    r_acc = ref.quantconv2d_float_input(q.contiguous(), t(c["packed"]), t(c["des"]), torch.ones(1).cuda(),
                                        torch.zeros(1).cuda(), None, c["stride"], c["pad"])
    assert float(r_acc.abs().max()) < 2 ** 24
    assert torch.equal(acc.float(), r_acc)


def test_act_quantize_random_quantizers_match_the_ieee_sequence():
    """Seeded random (scale, zero, qmin, qmax) — 4 decades of scale, fractional zero points, ranges that do not start at
    0 — with inputs planted on every rounding boundary and up to 3 ulps either side: the division-free quantizer
    (tight clamp, packed fp32 pipe, byte-permute packing) equals clamp(round(x / s - z)) computed with IEEE fp32 ops."""
    import random
    L = capi.lib()
    rnd = random.Random(5)
    N, C, H, W = 2, 64, 16, 32
    for it in range(60):
        qmax = rnd.choice([255, 255, 15, 127, 3, 200])
        qmin = rnd.choice([0, 0, 0, 1, 5]) if qmax > 10 else 0
        s = 10 ** rnd.uniform(-4, 1.5) * rnd.uniform(1, 2)
        z = rnd.choice([0.0, -rnd.uniform(0, 300), rnd.uniform(0, 50), -127.5, -0.5])
        scale = torch.tensor([s], dtype=torch.float32, device="cuda")
        zero = torch.tensor([z], dtype=torch.float32, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(it)
        x = torch.randn(N, C, H, W, device="cuda", generator=g) * (s * (qmax - qmin) * rnd.choice([0.1, 0.5, 2.0]))
        x = x + (zero + (qmin + qmax) / 2) * scale
        ks = torch.arange(qmin - 2, min(qmax + 3, qmin + 400), device="cuda", dtype=torch.float32)
        flat = x.view(-1)
        for d in range(-3, 4):
            v = (ks + 0.5 + zero) * scale
            for _ in range(abs(d)):
                v = torch.nextafter(v, torch.full_like(v, float("inf") if d > 0 else float("-inf")))
            flat[(d + 3) * ks.numel():(d + 4) * ks.numel()] = v
        qmn, qmx = torch.tensor([float(qmin)], device="cuda"), torch.tensor([float(qmax)], device="cuda")
        aq = capi.ActQuant(scale.data_ptr(), zero.data_ptr(), qmn.data_ptr(), qmx.data_ptr())
        out = torch.empty(N, H, W, C, dtype=torch.uint8, device="cuda")
        capi.check(L.qb200_act_quantize_nhwc(x.data_ptr(), N, C, H, W, ctypes.byref(aq), out.data_ptr(), None), "act_quantize")
        want = torch.clamp(torch.round(x / scale - zero), qmin, qmax).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
        assert torch.equal(out, want), (it, s, z, qmin, qmax)


# ---------------------------------------------------------------------------------------------------
# band quantizer (round 2): planes with H*W % 4 != 0 and sub-sampled inputs of strided 1x1 layers
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(3, 512, 7, 7), (2, 2048, 7, 7), (2, 96, 7, 7), (5, 3, 7, 7), (2, 40, 9, 9), (1, 64, 15, 15),
                                   (2, 24, 27, 27), (3, 32, 5, 3), (1, 7, 1, 3), (2, 64, 19, 21)])
def test_act_quantize_band_kernel_planes(shape):
    rng = np.random.default_rng(sum(shape) + 7)
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    for (s, z, lo, hi) in ((0.0371, -107.9, 0, 255), (1.0 / 3.0, -8.164, 0, 15), (0.02, -3.3, 2, 200)):
        aq = ActQ(s, z, lo, hi)
        want = oracle.act_quantize(x, s, z, lo, hi).astype(np.uint8)
        N, C, H, W = shape
        for off in (0, 1):       # off = 1: a base pointer that is only 4-byte aligned (4-byte cp.async path)
            buf = torch.zeros(x.size + 4, dtype=torch.float32, device="cuda")
            xt = buf[off:off + x.size].view(shape)
            xt.copy_(torch.from_numpy(x))
            q = c_act_quantize(xt, aq).cpu().numpy()
            assert np.array_equal(q[..., :C], want.transpose(0, 2, 3, 1)), (shape, off)
            assert not q[..., C:].any()


@pytest.mark.parametrize("cfg", [(2, 256, 14, 14, 64, 2), (3, 96, 28, 28, 48, 2), (2, 64, 7, 7, 32, 2), (2, 40, 15, 13, 24, 2),
                                 (1, 128, 56, 56, 64, 2), (2, 32, 12, 12, 16, 3), (2, 1024, 14, 14, 64, 2)])
def test_strided_1x1_through_the_band_quantizer(cfg):
    """1x1 / stride s / pad 0: the quantizer writes only the sampled pixels; int32 accumulators bit-exact vs the oracle."""
    N, C, H, W, K, stride = cfg
    c = random_conv_case(sum(cfg), N, C, H, W, K, 1, stride, 0)
    acc, out = run_case(c, ALGOS["umma"])
    qa, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(acc.cpu().numpy(), acc_ref)
    assert_close_1e3(out.cpu().numpy(), out_ref)


# ---------------------------------------------------------------------------------------------------
# CTA-pair variant (tcgen05.mma.cta_group::2, round 2): deep reductions
# ---------------------------------------------------------------------------------------------------
PAIR_CASES = [
    # N, C, H, W, K, R, stride, pad
    (2, 256, 14, 14, 256, 3, 1, 1),      # 3x3 @14: 4 pixel tiles, one channel tile
    (3, 128, 28, 28, 128, 3, 1, 1),      # 19 pixel tiles: odd tail pair, channel tile 128
    (5, 1024, 7, 7, 512, 1, 1, 0),       # 1x1 deep, two channel tiles, ragged last pixel tile
    (2, 512, 7, 7, 512, 3, 1, 1),        # 3x3 @7 (the tensor-bound layer of ResNet-50)
    (2, 128, 28, 28, 128, 3, 2, 1),      # strided 3x3
    (2, 1024, 14, 14, 2048, 1, 2, 0),    # strided 1x1 over the sub-sampled buffer, 8 channel tiles
    (9, 256, 9, 9, 768, 3, 1, 1),        # odd everything, 3 channel tiles
]


@pytest.mark.parametrize("cfg", PAIR_CASES)
def test_cta_pair_deep_reductions(cfg):
    """pair variant == one-CTA-per-tile variant == oracle: int32 accumulators bit-exact, fp32 bit-identical between the two
    kernels (same epilogue arithmetic)."""
    N, C, H, W, K, R, stride, pad = cfg
    c = random_conv_case(sum(cfg) + 11, N, C, H, W, K, R, stride, pad)
    acc_p, out_p = run_case(c, capi.ALGO_UMMA_PAIR)   # the CTA-pair kernel wherever it is supported
    acc_1, out_1 = run_case(c, ALGOS["umma2k"])       # one CTA per tile
    _, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(acc_p.cpu().numpy(), acc_ref)
    assert torch.equal(acc_p, acc_1) and torch.equal(out_p, out_1)
    assert_close_1e3(out_p.cpu().numpy(), out_ref)


# ---------------------------------------------------------------------------------------------------
# fused stem (round 2): few-channel 7x7-type layers — the grouped im2col rows are built in shared memory from the fp32
# input inside the tensor-core kernel instead of being written to and read back from HBM
# ---------------------------------------------------------------------------------------------------
STEM_CASES = [
    # N, C, H, W, K, R, stride, pad, relu_input
    (3, 3, 224, 224, 64, 7, 2, 3, False),    # the ImageNet ResNet stem: 98 tiles per image, tiles span 2-3 output rows
    (2, 3, 60, 60, 64, 7, 2, 3, False),      # 30x30 outputs: a tile spans 5-6 output rows, ragged last tile (900 = 7*128 + 4)
    (2, 1, 64, 48, 32, 7, 1, 3, True),       # one channel (64-byte rows, one k-block), stride 1, zero point 0
    (1, 2, 36, 40, 128, 7, 2, 3, False),     # two channels (128-byte rows), 128 output channels
    (5, 3, 32, 32, 256, 7, 1, 2, False),     # pad 2 (margin 4 > pad), asymmetric borders, K = 256
    (1, 3, 20, 16, 64, 7, 2, 3, False),      # a single ragged tile per image (10 x 8 outputs)
]


@pytest.mark.parametrize("cfg", STEM_CASES)
def test_fused_stem_rows_built_in_shared_memory(cfg):
    N, C, H, W, K, R, stride, pad, relu = cfg
    c = random_conv_case(sum(cfg[:8]) + 11, N, C, H, W, K, R, stride, pad, relu=relu)
    L = capi.lib()
    x = torch.from_numpy(c["x"]).cuda()
    assert L.qb200_conv_is_single_kernel(ctypes.byref(c["shape"]), x.data_ptr()) == 1
    acc_s, out_s = run_case(c, capi.ALGO_AUTO)            # the product's choice: fused stem
    acc_2, out_2 = run_case(c, ALGOS["umma2k"])           # quantizer writes the rows, the conv reads them back
    _, acc_ref, out_ref = oracle_case(c)
    assert np.array_equal(acc_s.cpu().numpy(), acc_ref)
    assert torch.equal(acc_s, acc_2) and torch.equal(out_s, out_2)
    assert_close_1e3(out_s.cpu().numpy(), out_ref)


# ---------------------------------------------------------------------------------------------------
# A-stationary fused quantize (round 2): 1x1 layers with several channel tiles, pixels quantized once per pixel tile
# ---------------------------------------------------------------------------------------------------
ASTAT_CASES = [
    # N, C, H, W, K
    (3, 128, 28, 28, 512),      # ResNet-50 layer2 expand: 2 k-blocks, 2 channel tiles, 7 pixel tiles per image (ragged last)
    (5, 256, 14, 14, 1024),     # layer3 expand: 4 k-blocks, 4 channel tiles, 128 + 68 pixel tiles
    (2, 64, 8, 8, 320),         # one k-block, ragged last channel tile (320 = 256 + 64), half-empty pixel tile
    (2, 512, 6, 6, 768),        # 8 k-blocks: the A ring holds one tile plus lookahead, 3 channel tiles
    (40, 128, 14, 14, 512),     # more pixel tiles than SMs: CTAs run several pixel tiles (ring wrap, phases)
    (2, 768, 4, 4, 512),        # 12 k-blocks: the ring is exactly one pixel tile
]


@pytest.mark.parametrize("cfg", ASTAT_CASES, ids=lambda c: "N{}C{}H{}W{}K{}".format(*c))
def test_a_stationary_fused_quantize(cfg):
    """int32 accumulators bit-exact vs the oracle and vs the two-kernel path; fp32 within 1e-3 and identical to the
    two-kernel path's (same epilogue arithmetic)."""
    N, C, H, W, K = cfg
    c = random_conv_case(sum(cfg), N, C, H, W, K, 1, 1, 0)
    acc, out = run_case(c, ALGOS["umma"])          # fused quantize forced: K > 256 -> A-stationary
    acc_2, out_2 = run_case(c, ALGOS["umma2k"])
    assert torch.equal(acc, acc_2) and torch.equal(out, out_2)
    if N * H * W * K * C <= 4e9:
        _, acc_ref, out_ref = oracle_case(c)
        assert np.array_equal(acc.cpu().numpy(), acc_ref)
        assert_close_1e3(out.cpu().numpy(), out_ref)
