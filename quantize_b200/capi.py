"""ctypes binding of the C-ABI library (include/qb200.h) — the boundary a non-torch host binds.

Pointers are passed as integers (torch `tensor.data_ptr()`), sizes as ints, the stream as a cudaStream_t
handle.  Nothing here computes; a missing library raises at load time (no fallback).
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqb200.so")

# names every build of the library must export (tests/test_abi.py checks them against include/qb200.h)
SYMBOLS = [
    "qb200_version", "qb200_last_error", "qb200_launch_count", "qb200_launch_count_reset",
    "qb200_packed_bytes", "qb200_tpack", "qb200_tunpack",
    "qb200_conv_out_hw", "qb200_padded_channels", "qb200_conv_prepared_bytes", "qb200_conv_prepare_weights",
    "qb200_conv_workspace_bytes", "qb200_act_quantize_nhwc", "qb200_set_conv_algo", "qb200_get_conv_algo",
    "qb200_quantconv2d_fused", "qb200_conv2d_q8_nhwc", "qb200_quantconv2d_weightonly",
    "qb200_conv_quantize_input", "qb200_conv_from_workspace", "qb200_conv_is_single_kernel", "qb200_watchdog_code",
    "qb200_quantconv2d_fused_ex", "qb200_conv_from_workspace_ex", "qb200_conv_handoff_supported", "qb200_maxpool2d_f32", "qb200_avgpool_global_f32", "qb200_quantlinear_weightonly", "qb200_fake_quantize_f32",
    "qb200_unpack_act_nhwc", "qb200_dequant_packed_f32", "qb200_quantconv2d_packed", "qb200_quantconv2d_packed_workspace_bytes",
    "qb200_quantlinear_packed", "qb200_minmax_f32", "qb200_minmax_workspace_bytes", "qb200_kthvalue_f32",
    "qb200_kthvalue_workspace_bytes", "qb200_set_rows_chunk", "qb200_conv_rows_chunk",
]

U8, I8, I16, I32, I64, F16, F32, F64, BF16 = range(9)
OUT_F32, OUT_ACC = 0, 1
ALGO_AUTO, ALGO_DIRECT, ALGO_UMMA, ALGO_UMMA_TWO_KERNELS, ALGO_UMMA_FUSED_QUANT, ALGO_UMMA_PAIR = 0, 1, 2, 3, 4, 5


class ConvShape(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("N", "C", "H", "W", "K", "Cg", "R", "S", "stride", "pad", "w_bits", "w_sign")]


class ActQuant(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("scale", "zero", "qmin", "qmax")]


class ConvTail(ctypes.Structure):
    _fields_ = [("residual", ctypes.c_void_p), ("relu", ctypes.c_int32),
                ("next_shape", ctypes.POINTER(ConvShape)), ("next_quant", ctypes.POINTER(ActQuant)),
                ("next_workspace", ctypes.c_void_p)]


class Qb200Error(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing — run `python -m quantize_b200.build` (there is no fallback)")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
        L.qb200_version.restype = ctypes.c_int
        L.qb200_last_error.restype = ctypes.c_char_p
        L.qb200_launch_count.restype = ctypes.c_uint64
        L.qb200_launch_count_reset.restype = None
        L.qb200_packed_bytes.restype = i64
        L.qb200_packed_bytes.argtypes = [i64, ctypes.c_int]
        L.qb200_tpack.argtypes = [vp, ctypes.c_int, i64, ctypes.c_int, ctypes.c_int, vp, vp, vp]
        L.qb200_tunpack.argtypes = [vp, i64, ctypes.c_int, ctypes.c_int, vp, vp]
        sp = ctypes.POINTER(ConvShape)
        ap = ctypes.POINTER(ActQuant)
        L.qb200_conv_out_hw.argtypes = [sp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.qb200_padded_channels.argtypes = [i32]
        L.qb200_padded_channels.restype = i32
        L.qb200_conv_prepared_bytes.argtypes = [sp]
        L.qb200_conv_prepared_bytes.restype = ctypes.c_size_t
        L.qb200_conv_workspace_bytes.argtypes = [sp]
        L.qb200_conv_workspace_bytes.restype = ctypes.c_size_t
        L.qb200_conv_prepare_weights.argtypes = [sp, vp, vp, vp]
        L.qb200_act_quantize_nhwc.argtypes = [vp, i32, i32, i32, i32, ap, vp, vp]
        L.qb200_set_conv_algo.argtypes = [ctypes.c_int]
        L.qb200_set_conv_algo.restype = None
        L.qb200_get_conv_algo.restype = ctypes.c_int
        L.qb200_quantconv2d_fused.argtypes = [sp, vp, vp, vp, i32, vp, ap, vp, vp, i32, vp]
        L.qb200_quantconv2d_fused_ex.argtypes = [sp, vp, vp, vp, i32, vp, ap, ctypes.POINTER(ConvTail), vp, vp, i32, vp]
        L.qb200_conv2d_q8_nhwc.argtypes = [sp, vp, vp, vp, i32, vp, ap, vp, i32, vp]
        L.qb200_conv_is_single_kernel.argtypes = [sp, vp]
        L.qb200_set_rows_chunk.argtypes = [ctypes.c_int]
        L.qb200_set_rows_chunk.restype = None
        L.qb200_conv_rows_chunk.argtypes = [sp]
        L.qb200_conv_quantize_input.argtypes = [sp, vp, ap, vp, vp]
        L.qb200_conv_from_workspace.argtypes = [sp, vp, vp, vp, i32, vp, ap, vp, i32, vp]
        L.qb200_quantconv2d_weightonly.argtypes = [sp, vp, vp, vp, vp, i32, vp, vp, vp]
        L.qb200_conv_from_workspace_ex.argtypes = [sp, vp, vp, vp, i32, vp, ap, ctypes.POINTER(ConvTail), vp, i32, vp]
        L.qb200_conv_handoff_supported.argtypes = [sp, sp]
        L.qb200_maxpool2d_f32.argtypes = [vp, i64, i32, i32, i32, i32, i32, vp, vp]
        L.qb200_avgpool_global_f32.argtypes = [vp, i64, i32, vp, vp]
        L.qb200_fake_quantize_f32.argtypes = [vp, i64, ap, vp, vp]
        L.qb200_quantlinear_weightonly.argtypes = [vp, i64, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp]
        f32 = ctypes.c_float
        L.qb200_unpack_act_nhwc.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp]
        L.qb200_dequant_packed_f32.argtypes = [vp, i32, i32, i64, i64, i32, vp, vp, i32, i32, vp, vp]
        L.qb200_quantconv2d_packed_workspace_bytes.argtypes = [sp]
        L.qb200_quantconv2d_packed_workspace_bytes.restype = ctypes.c_size_t
        L.qb200_quantconv2d_packed.argtypes = [sp, vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, vp, i32, vp]
        L.qb200_quantlinear_packed.argtypes = [vp, i32, i32, vp, vp, i64, i32, i32, vp, i32, i32, vp, vp, vp, vp, vp]
        L.qb200_minmax_workspace_bytes.argtypes = [i64]
        L.qb200_minmax_workspace_bytes.restype = ctypes.c_size_t
        L.qb200_minmax_f32.argtypes = [vp, i64, i64, i64, i32, vp, vp, i32, f32, f32, vp, vp, vp, vp]
        L.qb200_kthvalue_workspace_bytes.argtypes = [i64]
        L.qb200_kthvalue_workspace_bytes.restype = ctypes.c_size_t
        L.qb200_kthvalue_f32.argtypes = [vp, i64, i64, i64, i32, i64, vp, vp, vp]
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise Qb200Error(f"{what} failed ({rc}): {lib().qb200_last_error().decode(errors='replace')}")


def conv_shape(N, C, H, W, K, Cg, R, S, stride, pad, w_bits, w_sign):
    return ConvShape(N, C, H, W, K, Cg, R, S, stride, pad, w_bits, int(bool(w_sign)))


def conv_out_hw(shape):
    P, Q = ctypes.c_int32(), ctypes.c_int32()
    check(lib().qb200_conv_out_hw(ctypes.byref(shape), ctypes.byref(P), ctypes.byref(Q)), "conv_out_hw")
    return P.value, Q.value
