"""oracle — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg.  It is the checker; nothing under quantize_b200/ imports it.  See oracle/qoracle.c (C, exact integer
work) and oracle/fakequant.py (torch CPU restatement of the reference's module-level fake-quant path).
Parity status: pinned (tests/test_oracle_golden.py).
"""
import ctypes
import os

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = build_oracle.SO
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(build_oracle.SRC):
            build_oracle.build()
        _lib = ctypes.CDLL(so)
        _lib.qo_tpack_f32.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def tpack(x, n_bits, sign):
    """tpack.cu:203-255.  x: array of integer-valued numbers (any dtype/shape).  Returns (packed uint8 1-D,
    des int32 [n_bits, sign, *shape]); raises RuntimeError like the reference on range / n_bits errors."""
    if not (0 < n_bits <= 8):
        raise RuntimeError("n_bits must be in the range (0, 8]")
    x = np.ascontiguousarray(x)
    xf = x.astype(np.float32).reshape(-1)
    out = np.empty((xf.size * n_bits + 7) // 8, dtype=np.uint8)
    bad = lib().qo_tpack_f32(_p(xf), ctypes.c_int64(xf.size), n_bits, int(bool(sign)), _p(out))
    if bad:
        raise RuntimeError("The input tensor is out of range.")
    des = np.array([n_bits, int(bool(sign))] + list(x.shape), dtype=np.int32)
    return out, des


def tunpack(packed, des):
    """tpack.cu:429-476.  Returns int8 (signed) / uint8 array shaped des[2:]."""
    des = np.asarray(des).astype(np.int64)
    if des.size < 3:
        raise RuntimeError("The description is too short, which should be at least 3.")
    n_bits, sign = int(des[0]), int(des[1])
    if not (0 < n_bits <= 8):
        raise RuntimeError("n_bits must be in the range (0, 8]")
    shape = tuple(int(v) for v in des[2:])
    n = int(np.prod(shape))
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    out = np.empty(n, dtype=np.uint8)
    lib().qo_tunpack(_p(packed), ctypes.c_int64(n), n_bits, sign, _p(out))
    return (out.view(np.int8) if sign else out).reshape(shape)


def act_quantize(x, scale, zero, qmin, qmax):
    """quantizer.py:215.  fp32 array -> fp32 array of integers (same shape)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.empty_like(x)
    lib().qo_act_quantize(_p(x), ctypes.c_int64(x.size), ctypes.c_float(scale), ctypes.c_float(zero),
                          ctypes.c_float(qmin), ctypes.c_float(qmax), _p(q))
    return q


def conv_out_hw(H, W, R, S, stride, pad):
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


def conv_acc(qa, qw, stride, pad):
    """Exact integer conv.  qa uint8 [N,C,H,W]; qw int8 [K,Cg,R,S].  Returns (acc int32 [N,K,P,Q], wsum int32 [K,P,Q])."""
    qa = np.ascontiguousarray(qa, dtype=np.uint8)
    qw = np.ascontiguousarray(qw, dtype=np.int8)
    N, C, H, W = qa.shape
    K, Cg, R, S = qw.shape
    P, Q = conv_out_hw(H, W, R, S, stride, pad)
    acc = np.empty((N, K, P, Q), dtype=np.int32)
    wsum = np.empty((K, P, Q), dtype=np.int32)
    lib().qo_conv_acc(_p(qa), _p(qw), N, C, H, W, K, Cg, R, S, stride, pad, _p(acc), _p(wsum))
    return acc, wsum


def dequant(acc, wsum, s_a, z_a, s_w, bias):
    acc = np.ascontiguousarray(acc, dtype=np.int32)
    wsum = np.ascontiguousarray(wsum, dtype=np.int32)
    N, K, P, Q = acc.shape
    s_w = np.ascontiguousarray(np.asarray(s_w, dtype=np.float32).reshape(-1))
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    out = np.empty((N, K, P, Q), dtype=np.float32)
    lib().qo_dequant(_p(acc), _p(wsum), N, K, P, Q, ctypes.c_float(s_a), ctypes.c_float(z_a), _p(s_w),
                     int(s_w.size), None if b is None else _p(b), _p(out))
    return out


def quantconv2d_fused(x, w_packed, w_des, w_scale, bias, stride, pad, s_a, z_a, qmin, qmax):
    """The fused op's expected result: (out fp32, acc int32).  x fp32 NCHW."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    qa = act_quantize(x, s_a, z_a, qmin, qmax).astype(np.uint8)
    qw = tunpack(w_packed, w_des)
    if qw.dtype != np.int8:
        qw = qw.astype(np.int16).astype(np.int8) if qw.max(initial=0) < 128 else None
        if qw is None:
            raise ValueError("unsigned weights above 127 are not representable in the s8 B operand")
    acc, wsum = conv_acc(qa, qw, stride, pad)
    return dequant(acc, wsum, s_a, z_a, w_scale, bias), acc


def quantconv2d_float_input(x, w_packed, w_des, w_scale, w_zero, bias, stride, pad):
    """quantconv2d_float_input.cu:45-121 weight-only semantic (sequential fp32 accumulate)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, C, H, W = x.shape
    des = [int(v) for v in np.asarray(w_des)]
    n_bits, sign, K, Cw, R, S = des
    assert Cw == C, "the reference op indexes the weight with input.size(1) (SURVEY fact 6)"
    P, Q = conv_out_hw(H, W, R, S, stride, pad)
    w_scale = np.ascontiguousarray(np.asarray(w_scale, dtype=np.float32).reshape(-1))
    w_zero = np.ascontiguousarray(np.asarray(w_zero, dtype=np.float32).reshape(-1))
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    w_packed = np.ascontiguousarray(w_packed, dtype=np.uint8)
    out = np.empty((N, K, P, Q), dtype=np.float32)
    lib().qo_quantconv2d_float_input(_p(x), _p(w_packed), _p(w_scale), _p(w_zero), int(w_scale.size == 1),
                                     n_bits, sign, None if b is None else _p(b), _p(out),
                                     N, C, H, W, K, R, S, stride, pad)
    return out


def quantlinear_float_input(x, w_packed, w_des, w_scale, w_zero, bias):
    """quantlinear_float_input.cu:36-106 weight-only semantic (sequential fp32 FMA accumulate, bias last)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, in_f = x.shape
    n_bits, sign, out_f, in_w = [int(v) for v in np.asarray(w_des)[:4]]
    assert in_w == in_f
    w_scale = np.ascontiguousarray(np.asarray(w_scale, dtype=np.float32).reshape(-1))
    w_zero = np.ascontiguousarray(np.asarray(w_zero, dtype=np.float32).reshape(-1))
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    w_packed = np.ascontiguousarray(w_packed, dtype=np.uint8)
    out = np.empty((B, out_f), dtype=np.float32)
    lib().qo_quantlinear_float_input(_p(x), _p(w_packed), _p(w_scale), _p(w_zero), int(w_scale.size == 1), n_bits, sign,
                                     None if b is None else _p(b), _p(out), B, in_f, out_f)
    return out


def quantconv2d(in_packed, in_des, in_scale, in_zero, w_packed, w_des, w_scale, w_zero, bias, stride, pad):
    """quantconv2d.cu:49-264 (packed activations x packed weights, (q - zero) * scale on both, sequential fp32)."""
    n_bits, sign, N, C, H, W = [int(v) for v in np.asarray(in_des)[:6]]
    wb, ws, K, Cw, R, S = [int(v) for v in np.asarray(w_des)[:6]]
    assert Cw == C
    P, Q = conv_out_hw(H, W, R, S, stride, pad)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
    in_scale, in_zero, w_scale, w_zero = f(in_scale), f(in_zero), f(w_scale), f(w_zero)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    in_packed = np.ascontiguousarray(in_packed, dtype=np.uint8)
    w_packed = np.ascontiguousarray(w_packed, dtype=np.uint8)
    out = np.empty((N, K, P, Q), dtype=np.float32)
    lib().qo_quantconv2d(_p(in_packed), n_bits, sign, _p(in_scale), _p(in_zero), int(in_scale.size == 1), _p(w_packed), wb, ws,
                         _p(w_scale), _p(w_zero), int(w_scale.size == 1), None if b is None else _p(b), _p(out),
                         N, C, H, W, K, R, S, stride, pad)
    return out


def quantconv2d_acc(in_packed, in_des, w_packed, w_des, stride, pad):
    """exact integer accumulators sum(u * qw) over in-bounds taps, u = the stored (offset-binary) activation values."""
    n_bits, _, N, C, H, W = [int(v) for v in np.asarray(in_des)[:6]]
    wb, ws, K, Cw, R, S = [int(v) for v in np.asarray(w_des)[:6]]
    P, Q = conv_out_hw(H, W, R, S, stride, pad)
    acc = np.empty((N, K, P, Q), dtype=np.int32)
    lib().qo_quantconv2d_acc(_p(np.ascontiguousarray(in_packed, dtype=np.uint8)), n_bits,
                             _p(np.ascontiguousarray(w_packed, dtype=np.uint8)), wb, ws, _p(acc), N, C, H, W, K, R, S, stride, pad)
    return acc


def quantlinear(in_packed, in_des, in_scale, in_zero, w_packed, w_des, w_scale, w_zero, bias):
    """quantlinear.cu:39-133, :231-297 ((q + zero) convention, per-row input scale, sequential fp32)."""
    ib, isg, B, in_f = [int(v) for v in np.asarray(in_des)[:4]]
    wb, wsg, out_f, in_w = [int(v) for v in np.asarray(w_des)[:4]]
    assert in_w == in_f
    ex = lambda a, n: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32).reshape(-1), (n,)))
    in_scale, in_zero, w_scale, w_zero = ex(in_scale, B), ex(in_zero, B), ex(w_scale, out_f), ex(w_zero, out_f)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    out = np.empty((B, out_f), dtype=np.float32)
    lib().qo_quantlinear(_p(np.ascontiguousarray(in_packed, dtype=np.uint8)), ib, isg, _p(in_scale), _p(in_zero),
                         _p(np.ascontiguousarray(w_packed, dtype=np.uint8)), wb, wsg, _p(w_scale), _p(w_zero),
                         None if b is None else _p(b), _p(out), B, in_f, out_f)
    return out


def _rows(x, granularity, flag):
    """the estimator's flattened view (range/minmax.py:72-74, :86-88): one row per tensor / channel"""
    x = np.asarray(x, dtype=np.float32)
    if granularity == 0:
        return x.reshape(1, -1)
    if flag == "activation":
        x = np.swapaxes(x, 0, 1)
    return np.ascontiguousarray(x).reshape(x.shape[0], -1)


def minmax(x, granularity, flag, symmetric):
    """range/minmax.py:75-77, :89-91: (xmin, xmax) = (min, max), or (0, max|x|) for symmetric ranges; NaN propagates."""
    r = _rows(x, granularity, flag)
    if symmetric:
        lo, hi = np.zeros(r.shape[0], np.float32), np.abs(r).max(axis=1)
    else:
        lo, hi = r.min(axis=1), r.max(axis=1)
    return (lo[0], hi[0]) if granularity == 0 else (lo, hi)


def kthvalue(x, k, granularity, flag, use_abs=False):
    """range/minmax.py:78-84, :92-98: torch.kthvalue(...)[0] — the k-th smallest (1-based), NaN last."""
    r = _rows(x, granularity, flag)
    if use_abs:
        r = np.abs(r)
    if not (1 <= k <= r.shape[1]):
        raise RuntimeError("kthvalue(): selected number k out of range for dimension")
    v = np.sort(r, axis=1)[:, k - 1]          # numpy sorts NaN last, like torch
    return v[0] if granularity == 0 else v
