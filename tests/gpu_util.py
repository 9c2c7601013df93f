"""Helpers for the -m gpu tests: drive the C-ABI (libqb200.so via ctypes) with torch CUDA tensors."""
import ctypes

import numpy as np
import torch

from quantize_b200 import capi

TORCH_CODE = {torch.uint8: capi.U8, torch.int8: capi.I8, torch.int16: capi.I16, torch.int32: capi.I32,
              torch.int64: capi.I64, torch.float16: capi.F16, torch.float32: capi.F32, torch.float64: capi.F64,
              torch.bfloat16: capi.BF16}


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def c_tpack(x, n_bits, sign, check_range=True):
    L = capi.lib()
    n = x.numel()
    out = torch.full((max(int(L.qb200_packed_bytes(n, n_bits)), 0),), 0xAA, dtype=torch.uint8, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    capi.check(L.qb200_tpack(x.data_ptr(), TORCH_CODE[x.dtype], n, n_bits, int(sign), out.data_ptr(), flag.data_ptr(),
                             stream()), "tpack")
    return out, int(flag.item())


def c_tunpack(packed, n, n_bits, sign):
    L = capi.lib()
    out = torch.empty(n, dtype=torch.int8 if sign else torch.uint8, device=packed.device)
    capi.check(L.qb200_tunpack(packed.data_ptr(), n, n_bits, int(sign), out.data_ptr(), stream()), "tunpack")
    return out


class ActQ:
    """device-resident activation quantizer parameters"""

    def __init__(self, scale, zero, qmin, qmax, device="cuda"):
        self.t = [torch.tensor([float(v)], dtype=torch.float32, device=device) for v in (scale, zero, qmin, qmax)]
        self.c = capi.ActQuant(*[t.data_ptr() for t in self.t])
        self.scale, self.zero, self.qmin, self.qmax = float(scale), float(zero), float(qmin), float(qmax)


def c_act_quantize(x, aq):
    L = capi.lib()
    N, C, H, W = x.shape
    Cp = L.qb200_padded_channels(C)
    q = torch.full((N, H, W, Cp), 0xEE, dtype=torch.uint8, device=x.device)
    capi.check(L.qb200_act_quantize_nhwc(x.data_ptr(), N, C, H, W, ctypes.byref(aq.c), q.data_ptr(), stream()),
               "act_quantize")
    return q


def c_prepare(shape, w_packed):
    L = capi.lib()
    buf = torch.empty(L.qb200_conv_prepared_bytes(ctypes.byref(shape)), dtype=torch.uint8, device=w_packed.device)
    capi.check(L.qb200_conv_prepare_weights(ctypes.byref(shape), w_packed.data_ptr(), buf.data_ptr(), stream()),
               "prepare")
    return buf


def c_conv_fused(shape, x, prepared, w_scale, bias, aq, out_kind=capi.OUT_F32, algo=capi.ALGO_AUTO):
    L = capi.lib()
    P, Q = capi.conv_out_hw(shape)
    ws = torch.empty(L.qb200_conv_workspace_bytes(ctypes.byref(shape)), dtype=torch.uint8, device=x.device)
    out = torch.full((shape.N, shape.K, P, Q), -12345, dtype=torch.int32 if out_kind == capi.OUT_ACC else torch.float32,
                     device=x.device)
    L.qb200_set_conv_algo(algo)
    try:
        capi.check(L.qb200_quantconv2d_fused(ctypes.byref(shape), x.data_ptr(), prepared.data_ptr(), w_scale.data_ptr(),
                                             w_scale.numel(), bias.data_ptr() if bias is not None else None,
                                             ctypes.byref(aq.c), ws.data_ptr(), out.data_ptr(), out_kind, stream()),
                   "quantconv2d_fused")
    finally:
        L.qb200_set_conv_algo(capi.ALGO_AUTO)
    torch.cuda.synchronize()
    return out


def c_weightonly(shape, x, w_packed, w_scale, w_zero, bias):
    L = capi.lib()
    P, Q = capi.conv_out_hw(shape)
    out = torch.empty((shape.N, shape.K, P, Q), dtype=torch.float32, device=x.device)
    capi.check(L.qb200_quantconv2d_weightonly(ctypes.byref(shape), x.data_ptr(), w_packed.data_ptr(), w_scale.data_ptr(),
                                              w_zero.data_ptr(), w_scale.numel(),
                                              bias.data_ptr() if bias is not None else None, out.data_ptr(), stream()),
               "weightonly")
    torch.cuda.synchronize()
    return out


def random_conv_case(seed, N, C, H, W, K, R, stride, pad, groups=1, w_bits=8, a_bits=8, w_sign=True, relu=False,
                     per_tensor=False, device="cuda"):
    """Seeded synthetic layer: fp32 activations, integer weights packed with the ORACLE's tpack (reference format)."""
    import oracle
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    if relu:
        x = np.maximum(x, 0)
    Cg = C // groups
    lo, hi = (-(1 << (w_bits - 1)) + 1, (1 << (w_bits - 1)) - 1) if w_sign else (0, (1 << w_bits) - 1)
    qw = rng.integers(lo, hi + 1, size=(K, Cg, R, R)).astype(np.int64)
    packed, des = oracle.tpack(qw, w_bits, w_sign)
    qmax = float((1 << a_bits) - 1)
    xmin, xmax = float(x.min()), float(x.max())
    a_scale = np.float32((xmax - xmin) / qmax)
    a_zero = np.float32(np.float32(xmin) / a_scale)     # range/minmax.py:136-143
    w_scale = (rng.random(1 if per_tensor else K) * 0.02 + 0.001).astype(np.float32)
    bias = rng.standard_normal(K).astype(np.float32)
    shape = capi.conv_shape(N, C, H, W, K, Cg, R, R, stride, pad, w_bits, w_sign)
    return dict(x=x, qw=qw, packed=packed, des=des, a_scale=float(a_scale), a_zero=float(a_zero), qmin=0.0, qmax=qmax,
                w_scale=w_scale, bias=bias, shape=shape, stride=stride, pad=pad, groups=groups)
