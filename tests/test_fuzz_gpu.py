"""GPU: seeded random layer shapes and chains.  Every kernel variant the dispatcher can pick (tensor-core plain / halo /
im2col-rows / fused-quantize / ragged, the fused depthwise kernel, the CUDA-core kernels, the hand-off chain) must give
the SAME bits: int32 accumulators and fp32 outputs of the default dispatch == the CUDA-core reference kernel == the
forced tensor-core variants; chains == layer-by-layer calls.  (A longer run of this fuzz found a shared-memory sizing
bug of the depthwise kernel for 7x7 / stride-2 shapes.)"""
import random

import numpy as np
import pytest
import torch

from quantize_b200 import capi
from gpu_util import random_conv_case
from test_conv_gpu import run_case

pytestmark = pytest.mark.gpu


def _random_layer(rnd):
    R = rnd.choice([1, 1, 3, 3, 5, 7])
    stride = rnd.choice([1, 1, 2])
    pad = rnd.choice([0, R // 2]) if R > 1 else rnd.choice([0, 0, 1])
    H = rnd.randint(max(1, R - 2 * pad), 40)
    W = rnd.randint(max(1, R - 2 * pad), 40)
    dw = rnd.random() < 0.2
    C = rnd.choice([1, 3, 4, 7, 16, 24, 32, 33, 64, 96, 128, 144, 200, 256, 320])
    K = C if dw else rnd.choice([1, 8, 16, 24, 40, 64, 96, 128, 130, 256, 300, 512])
    return (rnd.randint(1, 4), C, H, W, K, R, stride, pad, C if dw else 1, rnd.choice([8, 8, 4, 5]), rnd.choice([8, 8, 4]))


@pytest.mark.parametrize("seed", [1, 7, 123])
def test_random_layers_all_variants_bit_identical(seed):
    rnd = random.Random(seed)
    for it in range(40):
        cfg = _random_layer(rnd)
        N, C, H, W, K, R, stride, pad, groups, wb, ab = cfg
        c = random_conv_case(1000 * seed + it, N, C, H, W, K, R, stride, pad, groups, wb, ab, True, rnd.random() < 0.5, False)
        a0, o0 = run_case(c, capi.ALGO_AUTO)
        a1, o1 = run_case(c, capi.ALGO_DIRECT)
        assert torch.equal(a0, a1) and torch.equal(o0, o1), cfg
        if groups == 1:
            a4, o4 = run_case(c, capi.ALGO_UMMA_FUSED_QUANT)
            assert torch.equal(a0, a4) and torch.equal(o0, o4), cfg


def test_random_chains_equal_layer_by_layer(engine):
    rnd = random.Random(5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for it in range(25):
        N, C, H = rnd.randint(1, 3), rnd.choice([16, 24, 64, 96, 128, 256]), rnd.choice([5, 8, 14, 28, 30])
        x = torch.randn(N, C, H, H, generator=torch.Generator().manual_seed(it)).cuda()
        tuples, want, prev, Cc, Hc = [], x, x, C, H
        for li in range(rnd.randint(2, 3)):
            R, stride = rnd.choice([1, 3]), rnd.choice([1, 1, 2])
            K = rnd.choice([16, 24, 64, 96, 128, 200, 256, 512])
            c = random_conv_case(100 * it + li, N, Cc, Hc, Hc, K, R, stride, R // 2, 1)
            lo, hi = float(want.min()), float(want.max())
            a_scale = torch.tensor([(hi - lo) / 255.0 + 1e-9], dtype=torch.float32).cuda()
            a_zero = torch.tensor([lo], dtype=torch.float32).cuda() / a_scale
            tup = (t(c["packed"]), t(c["des"]), t(c["w_scale"]), torch.zeros(K).cuda(), t(c["bias"]), stride, R // 2,
                   a_scale, a_zero, 0, 255, rnd.random() < 0.6)
            tuples.append(tup)
            prev = want
            want = engine.quantconv2d_float_input(want, *tup[:7], input_scale=a_scale, input_zero=a_zero, input_qmin=0,
                                                  input_qmax=255, fuse_relu=tup[11])
            Cc, Hc = K, want.shape[2]
        assert torch.equal(engine.quantconv2d_chain(x, tuples), want), it
        res, last = torch.randn_like(want), tuples[-1]
        want_r = engine.quantconv2d_float_input(prev, *last[:7], input_scale=last[7], input_zero=last[8], input_qmin=0,
                                                input_qmax=255, residual=res, fuse_relu=last[11])
        assert torch.equal(engine.quantconv2d_chain(x, tuples, residual=res), want_r), it
