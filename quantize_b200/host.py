"""Host-side mirror of the reference's quantized conv layer, routed through the B200 engine.

The reference's own modules (modelzoo/modules/quantconv2d.py, quantizer.py, range/minmax.py) stay unchanged and can
use this engine directly.  They are not available on a machine without the reference checkout, so
this file restates — same names, same argument meaning, same arithmetic — the part of their interface the hot path
needs:  MinMax / MAMinMax range estimation, Quantizer (calibrate / simulate / pack) and QuantConv2d
(BN folding, calibrate, fake-quant `_forward`, `pack()`, packed `forward`).  The one intended difference is the
packed `forward`: the reference has the op call commented out and runs a float F.conv2d (quantconv2d.py:198-210);
here it calls `quant_engine.quantconv2d_float_input` with the activation quantizer's parameters (the fused path).

All tensor math in this file is calibration-time PyTorch (plumbing); the inference arithmetic is in csrc/.
"""
from copy import deepcopy

import torch
import torch.nn as nn
from torch import Tensor

from . import engine as _engine


class MinMax:
    """reference modelzoo/modules/range/minmax.py:12-145.

    CUDA fp32 tensors are reduced by the engine (`quant_engine.minmax`: min, max / abs-max and the estimator's state
    update in ONE pass over the tensor instead of 2-3 torch reductions plus the update kernels; percentile ranges through
    `quant_engine.kthvalue`, an exact radix select).  Selections are exact, so the values equal torch's bit for bit;
    other tensors (CPU, other dtypes) take the reference's torch expressions."""

    use_engine = True   # class-level switch (tests compare both paths)

    def __init__(self, n_bits=8, symmetric=True, signed=True, granularity="layer", percentile=0.0, **kwargs):
        self.n_bits, self.symmetric, self.signed, self.granularity = n_bits, symmetric, signed, granularity
        self.percentile = percentile

    def update(self, xmin, xmax):                                   # minmax.py:44-60
        if "xmin" not in self.__dict__:
            self.xmin, self.xmax = xmin, xmax
        else:
            self.xmin, self.xmax = torch.min(self.xmin, xmin), torch.max(self.xmax, xmax)
        return self.xmin, self.xmax

    _update_mode = 1    # engine code of this class's update rule (qb200_minmax_f32: 1 running min/max, 2 moving average)

    def _engine_ok(self, x):
        return MinMax.use_engine and x.is_cuda and x.dtype == torch.float32 and x.numel() > 0

    def _gran(self):
        if self.granularity in ("L", "Layer", "layer"):
            return 0
        if self.granularity in ("C", "Channel", "channel"):
            return 1
        raise NotImplementedError(f"Granularity {self.granularity} not implemented.")

    def _range_engine(self, x: Tensor, flag: str, accumulate: bool):
        qe, gran, fl = _engine.load(), self._gran(), 1 if flag == "activation" else 0
        x = x.detach().contiguous()
        if self.percentile == 0.0:
            have = "xmin" in self.__dict__
            if accumulate and have and self.xmin.is_cuda and self.xmin.dtype == torch.float32 and self.xmin.is_contiguous() \
                    and self.xmax.is_contiguous():
                mode, mom = self._update_mode, float(getattr(self, "momentum", 0.0))
                if mode == 2 and not (0.0 <= mom <= 1.0):
                    mode = 1                                        # minmax.py:199-201
                xmin, xmax = qe.minmax(x, gran, fl, self.symmetric, mode, mom, self.xmin, self.xmax)
                self.xmin, self.xmax = xmin, xmax
                return xmin, xmax
            xmin, xmax = qe.minmax(x, gran, fl, self.symmetric)
        else:
            n = x.numel() if gran == 0 else (x.numel() // x.shape[1] if fl else x.numel() // x.shape[0])
            if not self.symmetric:                                  # minmax.py:78-80, :92-94
                hi_k = int(n * (1 - self.percentile)) if gran == 0 else int(n * (1 + self.percentile))
                xmin = qe.kthvalue(x, int(n * self.percentile) + 1, gran, fl, False)
                xmax = qe.kthvalue(x, hi_k, gran, fl, False)
            else:                                                   # :81-84, :95-98
                xmax = qe.kthvalue(x, int(n * (1 - self.percentile)), gran, fl, True)
                xmin = torch.zeros_like(xmax)
        return self.update(xmin, xmax) if accumulate else (xmin, xmax)

    def range(self, x: Tensor, flag: str, accumulate=True):         # minmax.py:62-108
        if self._engine_ok(x):
            return self._range_engine(x, flag, accumulate)
        if self.granularity in ("L", "Layer", "layer"):
            x = x.flatten(0)
            if self.percentile == 0.0:
                xmin = x.min() if not self.symmetric else torch.tensor(0.0).to(x.device)
                xmax = x.max() if not self.symmetric else x.abs().max()
            elif not self.symmetric:
                xmin = x.kthvalue(int(x.numel() * self.percentile) + 1)[0]
                xmax = x.kthvalue(int(x.numel() * (1 - self.percentile)))[0]
            else:
                xmin = torch.tensor(0.0).to(x.device)
                xmax = x.abs().kthvalue(int(x.numel() * (1 - self.percentile)))[0]
        elif self.granularity in ("C", "Channel", "channel"):
            if flag == "activation":
                x = x.transpose(0, 1)
            x = x.flatten(1)
            if self.percentile == 0.0:
                xmin = x.min(dim=1)[0] if not self.symmetric else torch.zeros(x.shape[0]).to(x.device)
                xmax = x.max(dim=1)[0] if not self.symmetric else x.abs().max(dim=1)[0]
            elif not self.symmetric:
                xmin = x.kthvalue(int(x.shape[1] * self.percentile) + 1, dim=1)[0]
                xmax = x.kthvalue(int(x.shape[1] * (1 + self.percentile)), dim=1)[0]
            else:
                xmin = torch.zeros(x.shape[0]).to(x.device)
                xmax = x.abs().kthvalue(int(x.shape[1] * (1 - self.percentile)), dim=1)[0]
        else:
            raise NotImplementedError(f"Granularity {self.granularity} not implemented.")
        return self.update(xmin, xmax) if accumulate else (xmin, xmax)

    def quantize(self, xmin: Tensor, xmax: Tensor):                 # minmax.py:110-145
        n_bits = self.n_bits
        if self.symmetric:
            if self.signed:
                qmax, qmin = (1 << (n_bits - 1)) - 1, -(1 << (n_bits - 1))
                quant_range = float(qmax - qmin - 1) / 2
            else:
                qmax, qmin = (1 << n_bits) - 1, 0
                quant_range = float(qmax - qmin)
            scale = torch.max(xmin.abs(), xmax.abs()) / quant_range
            zero = torch.zeros_like(scale)
        else:
            qmax, qmin = (1 << n_bits) - 1, 0
            scale = (xmax - xmin) / float(qmax - qmin)
            zero = xmin / scale
        return scale, zero, qmin, qmax

    def __call__(self, flag: str, x: Tensor, **kwargs):
        return self.quantize(*self.range(x, flag))


class MAMinMax(MinMax):
    """reference minmax.py:148-203: moving-average min/max."""

    _update_mode = 2

    def __init__(self, momentum=0.1, **kw):
        super().__init__(**kw)
        self.momentum = momentum

    def update(self, xmin, xmax):
        if "xmin" not in self.__dict__:
            self.xmin, self.xmax = xmin, xmax
        elif 0.0 <= self.momentum <= 1.0:
            self.xmin = self.momentum * xmin + (1 - self.momentum) * self.xmin
            self.xmax = self.momentum * xmax + (1 - self.momentum) * self.xmax
        else:
            self.xmin, self.xmax = torch.min(self.xmin, xmin), torch.max(self.xmax, xmax)
        return self.xmin, self.xmax


RANGES = {"minmax": MinMax, "maminmax": MAMinMax}


class Quantizer(nn.Module):
    """reference modelzoo/modules/quantizer.py:43-306 (no adaround / awq / static_scale)."""

    def __init__(self, n_bits=8, symmetric=True, signed=True, granularity="layer", range={"name": "maminmax"},
                 flag="weight", n_channels=1, dim=4):
        super().__init__()
        self.n_bits, self.symmetric, self.signed, self.granularity = n_bits, symmetric, signed, granularity
        self.flag, self.n_channels, self.dim = flag, n_channels, dim
        range = deepcopy(dict(range))
        self.range = range.pop("name")
        self.range_estimator = RANGES[self.range](n_bits=n_bits, symmetric=symmetric, signed=signed,
                                                  granularity=granularity, **range)
        if granularity in ("L", "Layer", "layer"):
            self.n_channels = 1
        self.scale = nn.Parameter(torch.ones(self.n_channels).view(*self.shape))
        self.zero = nn.Parameter(torch.zeros(self.n_channels).view(*self.shape))
        self.register_buffer("qmin", torch.tensor(-(2 ** (n_bits - 1))))
        self.register_buffer("qmax", torch.tensor(2 ** (n_bits - 1) - 1))
        self.quantized = False
        self.packed = False

    use_engine = True   # class-level switch: CUDA per-tensor fake quantization through the engine's one-pass kernel

    @property
    def shape(self):                                                # quantizer.py:137-144
        shape = [1] * self.dim
        shape[1 if self.flag == "activation" else 0] = -1
        return shape

    def quant(self, quantized=True):
        self.quantized = bool(quantized)

    def quantize_int(self, x: Tensor) -> Tensor:                    # quantizer.py:31 + :215
        return (x / self.scale.view(*self.shape) - self.zero.view(*self.shape)).round().clamp(self.qmin, self.qmax)

    def simulate(self, x: Tensor):                                  # quantizer.py:196-226
        self.dim = x.dim()
        scale, zero = self.scale.view(*self.shape), self.zero.view(*self.shape)
        if (Quantizer.use_engine and not self.packed and x.is_cuda and x.dtype == torch.float32 and scale.numel() == 1
                and not (x.requires_grad and torch.is_grad_enabled())):
            # per-tensor fake quantization on the GPU: one engine kernel instead of five torch kernels, same bits
            return _engine.load().fake_quantize(x.detach().contiguous(), scale, zero, self.qmin, self.qmax)
        q = self.quantize_int(x)
        if not self.packed:
            return (q + zero).mul(scale)
        return q, scale, zero

    def pack(self, x):                                              # quantizer.py:228-246
        self.packed = True
        if self.flag == "weight":
            return self.quantize_int(x), self.scale.detach().clone(), self.zero.detach().clone()
        return x, None, None

    @torch.no_grad()
    def calibrate(self, x: Tensor):                                 # quantizer.py:248-263
        device = self.scale.device
        scale, zero, qmin, qmax = self.range_estimator(self.flag, x)
        self.scale.data = scale.view(*self.shape).to(device)
        self.zero.data = zero.view(*self.shape).to(device)
        self.qmin.copy_(torch.tensor(qmin).to(device))
        self.qmax.copy_(torch.tensor(qmax).to(device))

    def forward(self, x: Tensor):                                   # quantizer.py:265-279
        if not self.quantized or self.n_bits >= 32:
            return x
        return self.simulate(x)


DEFAULT_W = dict(n_bits=8, symmetric=True, signed=True, granularity="channel", range={"name": "minmax", "percentile": 0.0})
DEFAULT_A = dict(n_bits=8, symmetric=False, granularity="layer", range={"name": "maminmax", "percentile": 0.0, "momentum": 0.1})
# = the reference's configs/runners/ptq/minmax/base.yaml:1-19


class QuantConv2d(nn.Conv2d):
    """reference modelzoo/modules/quantconv2d.py:20-235, packed forward routed through the engine."""

    def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d = None, w_setting=None, a_setting=None):
        super().__init__(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                         conv.dilation, conv.groups, conv.bias is not None or bn is not None, conv.padding_mode)
        assert conv.padding_mode == "zeros" and conv.dilation == (1, 1), "the op has no dilation / padding modes"
        assert conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1], \
            "the op takes one stride and one padding (quantconv2dop.py:82-85)"
        w = conv.weight.detach().clone()
        b = conv.bias.detach().clone() if conv.bias is not None else (torch.zeros(conv.out_channels) if bn is not None else None)
        if bn is not None:                                          # quantconv2d.py:115-128 (bn_folding)
            std = torch.sqrt(bn.running_var + bn.eps)
            b = bn.bias.detach() + (b - bn.running_mean) * bn.weight.detach() / std   # same operation order as :117-121
            w = w * (bn.weight.detach() / std).reshape(-1, 1, 1, 1)                   # :124-128
        self.weight.data = w
        if b is not None:
            self.bias.data = b.reshape(-1)
        self.w_quantizer = Quantizer(**(w_setting or DEFAULT_W), flag="weight", n_channels=conv.out_channels, dim=4)
        self.a_quantizer = Quantizer(**(a_setting or DEFAULT_A), flag="activation", n_channels=conv.in_channels, dim=4)
        self.calibrating = False
        self.packed = False
        self.use_engine = True
        self.fuse_relu = False      # set by fuse_resnet_blocks: the ReLU that follows this conv runs in its epilogue

    def calibrate(self, x: Tensor):                                 # quantconv2d.py:141-152
        self.a_quantizer.calibrate(x.detach().clone())
        self.w_quantizer.calibrate(self.weight.detach().clone())

    def _forward(self, x: Tensor) -> Tensor:                        # quantconv2d.py:154-168
        if self.calibrating:
            self.calibrate(x)
        x = self.a_quantizer(x)
        weight = self.w_quantizer(self.weight)
        return self._conv_forward(x, weight, self.bias)

    @torch.no_grad()
    def pack(self):                                                 # quantconv2d.py:170-196
        self.requires_grad_(False)
        self.a_quantizer.pack(None)
        weight, w_scale, w_zero = self.w_quantizer.pack(self.weight)
        self.register_buffer("w_scale", w_scale)
        self.register_buffer("w_zero", w_zero)
        qe = _engine.load()
        self.weight.data, w_des = qe.tpack(weight.contiguous(), self.w_quantizer.n_bits, self.w_quantizer.signed)
        self.register_buffer("w_des", w_des)
        self.w_quantizer = None
        self.packed = True

    def symmetric_weights(self) -> bool:
        if not hasattr(self, "_w_sym"):
            self._w_sym = bool((self.w_zero == 0).all().item())
        return self._w_sym

    def byte_activations(self) -> bool:
        """the int8 hand-off stores activations as unsigned bytes: the quantizer's range must lie inside [0, 255]"""
        if not hasattr(self, "_a_u8"):
            a = self.a_quantizer
            self._a_u8 = bool(0 <= float(a.qmin) <= float(a.qmax) <= 255)
        return self._a_u8

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """reference quantconv2d.py:212-235: a checkpoint that holds `w_des` is a PACKED one — pack this (fresh) module first
        so that the packed buffers exist with the right shapes, then load.  The reference then tunpacks the weight again
        (its packed forward is a float conv); this mirror keeps the packed byte stream, which is what the engine consumes."""
        device = self.weight.device
        if prefix + "w_des" in state_dict and not self.packed:
            self.pack()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        for cached in ("_w_sym", "_a_u8"):
            self.__dict__.pop(cached, None)
        self.to(device)

    def chain_args(self):
        """this layer as an element of engine.quantconv2d_chain's `layers` (the op's arguments + relu-after flag)."""
        a = self.a_quantizer
        return (self.weight, self.w_des, self.w_scale, self.w_zero, self.bias, self.stride[0], self.padding[0],
                a.scale, a.zero, a.qmin, a.qmax, bool(self.fuse_relu))

    def forward(self, x: Tensor, residual: Tensor = None) -> Tensor:    # quantconv2d.py:198-210
        """`residual` / `self.fuse_relu` are this mirror's extension: out = relu(out + residual) in the conv's epilogue
        (bit-identical to the separate torch ops, which is what every non-engine branch below executes)."""
        if self.packed and self.use_engine:
            a = self.a_quantizer
            return _engine.load().quantconv2d_float_input(
                x.contiguous(), self.weight, self.w_des, self.w_scale, self.w_zero, self.bias, self.stride[0],
                self.padding[0], input_scale=a.scale, input_zero=a.zero, input_qmin=a.qmin, input_qmax=a.qmax,
                residual=None if residual is None else residual.contiguous(), fuse_relu=self.fuse_relu)
        if not self.packed:
            out = self._forward(x)
        else:
            # the reference's packed forward, kept for cross-checks: float conv on dequantized operands
            q, a_scale, a_zero = self.a_quantizer.simulate(x)
            w = _engine.load().tunpack(self.weight, self.w_des)
            out = self._conv_forward((q + a_zero).mul_(a_scale), (w + self.w_zero).mul_(self.w_scale), self.bias)
        if residual is not None:
            out = out + residual
        return torch.relu(out) if self.fuse_relu else out


class QuantLinear(nn.Linear):
    """reference modelzoo/modules/quantlinear.py:17-186 (no bias correction), packed forward routed through the engine's
    quantlinear_float_input with the activation quantizer's parameters (the integer path: the 1x1 case of the conv)."""

    def __init__(self, lin: nn.Linear, w_setting=None, a_setting=None):
        super().__init__(lin.in_features, lin.out_features, lin.bias is not None)
        self.weight.data = lin.weight.detach().clone()                                # quantlinear.py:77-79
        if lin.bias is not None:
            self.bias.data = lin.bias.detach().clone()
        self.w_quantizer = Quantizer(**(w_setting or DEFAULT_W), flag="weight", n_channels=lin.out_features, dim=2)
        self.a_quantizer = Quantizer(**(a_setting or DEFAULT_A), flag="activation", n_channels=lin.in_features, dim=2)
        self.calibrating = False
        self.packed = False
        self.use_engine = True

    def calibrate(self, x: Tensor):                                 # quantlinear.py:93-104
        self.a_quantizer.calibrate(x.detach().clone())
        self.w_quantizer.calibrate(self.weight.detach().clone())

    def _forward(self, x: Tensor) -> Tensor:                        # quantlinear.py:106-121
        if self.calibrating:
            self.calibrate(x)
        x = self.a_quantizer(x)
        weight = self.w_quantizer(self.weight)
        return nn.functional.linear(x, weight, self.bias)

    @torch.no_grad()
    def pack(self):                                                 # quantlinear.py:123-150
        self.requires_grad_(False)
        self.a_quantizer.pack(None)
        weight, w_scale, w_zero = self.w_quantizer.pack(self.weight)
        self.register_buffer("w_scale", w_scale)
        self.register_buffer("w_zero", w_zero)
        self.weight.data, w_des = _engine.load().tpack(weight.contiguous(), self.w_quantizer.n_bits, self.w_quantizer.signed)
        self.register_buffer("w_des", w_des)
        self.w_quantizer = None
        self.packed = True

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """reference quantlinear.py:170-186 (see QuantConv2d._load_from_state_dict: the packed stream is kept)."""
        device = self.weight.device
        if prefix + "w_des" in state_dict and not self.packed:
            self.pack()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        self.to(device)

    def forward(self, x: Tensor) -> Tensor:                         # quantlinear.py:152-163
        if not self.packed:
            return self._forward(x)
        if self.use_engine and x.is_cuda:
            a = self.a_quantizer
            x2 = x.reshape(-1, self.in_features).contiguous()
            out = _engine.load().quantlinear_float_input(x2, self.weight, self.w_des, self.w_scale.reshape(-1),
                                                         self.w_zero.reshape(-1), self.bias, input_scale=a.scale,
                                                         input_zero=a.zero, input_qmin=a.qmin, input_qmax=a.qmax)
            return out.reshape(*x.shape[:-1], self.out_features)
        # the reference's packed forward: float linear on dequantized operands
        q, a_scale, a_zero = self.a_quantizer.simulate(x)
        w = _engine.load().tunpack(self.weight, self.w_des)
        return nn.functional.linear((q + a_zero).mul_(a_scale), (w + self.w_zero).mul_(self.w_scale), self.bias)


class EngineMaxPool2d(nn.Module):
    """nn.MaxPool2d (square kernel / stride, no dilation, floor mode) on the engine's kernel; same values as torch."""

    def __init__(self, pool: nn.MaxPool2d):
        super().__init__()
        one = lambda v: v if isinstance(v, int) else v[0]
        assert not pool.ceil_mode and one(pool.dilation) == 1 and not pool.return_indices
        self.kernel_size, self.stride, self.padding = one(pool.kernel_size), one(pool.stride or pool.kernel_size), one(pool.padding)

    def forward(self, x, out=None):
        if x.is_cuda and x.dtype == torch.float32:
            return _engine.load().max_pool2d(x.contiguous(), self.kernel_size, self.stride, self.padding, out=out)
        y = torch.nn.functional.max_pool2d(x, self.kernel_size, self.stride, self.padding)
        return y if out is None else out.copy_(y)


class EngineGlobalAvgPool2d(nn.Module):
    """nn.AdaptiveAvgPool2d((1, 1)) on the engine's kernel (one warp per plane); torch's op elsewhere."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float32 and x.dim() == 4:
            return _engine.load().avg_pool_global(x.contiguous())
        return torch.nn.functional.adaptive_avg_pool2d(x, (1, 1))


def _chainable(convs):
    """int8 hand-off between consecutive convs needs the engine path, symmetric weights (w_zero == 0) and activation
    ranges that fit an unsigned byte."""
    return all(c.packed and c.use_engine and c.symmetric_weights() and c.byte_activations() for c in convs)


def _shortcut_shares_handoff(block) -> bool:
    """Can the 1x1 shortcut conv of a down-sampling block read the int8 hand-off written for the block's conv1?  Needs: an
    engine-side packed shortcut conv (conv + folded BN) without padding; conv1 a 1x1 / stride-1 / pad-0 conv (its hand-off
    workspace is then the plain NHWC byte tensor of the block's input); bit-identical activation quantizer parameters
    (true after calibration on the same data: both quantizers observe the block's input).  Checked once per block and again
    whenever a parameter tensor is replaced or modified in place."""
    ds = block.downsample
    c1 = getattr(block, "conv1", None)

    def stamp(q):   # identity + in-place version of the quantizer's parameter tensors (host-side, no synchronisation)
        return tuple((t.data_ptr(), t._version) if isinstance(t, Tensor) else t
                     for t in (getattr(q, n, None) for n in ("scale", "zero", "qmin", "qmax")))
    key = None
    if isinstance(ds, nn.Sequential) and len(ds) >= 1 and isinstance(ds[0], QuantConv2d) and isinstance(c1, QuantConv2d):
        key = (stamp(c1.a_quantizer), stamp(ds[0].a_quantizer))
    cached = getattr(block, "_ds_shared", None)
    if cached is not None and getattr(block, "_ds_shared_key", None) == key:
        return cached
    ok = False
    if (isinstance(ds, nn.Sequential) and len(ds) >= 1 and isinstance(ds[0], QuantConv2d) and isinstance(c1, QuantConv2d)
            and all(isinstance(m, nn.Identity) for m in list(ds)[1:])):
        d = ds[0]
        ok = (d.packed and d.use_engine and not d.fuse_relu and d.groups == 1 and d.kernel_size == (1, 1) and d.padding == (0, 0)
              and d.symmetric_weights() and d.byte_activations()
              and c1.groups == 1 and c1.kernel_size == (1, 1) and c1.stride == (1, 1) and c1.padding == (0, 0)
              and c1.in_channels == d.in_channels and _chainable((c1,)))
        if ok:
            a, b = c1.a_quantizer, d.a_quantizer
            ok = all(bool(torch.equal(torch.as_tensor(getattr(a, n)).float().reshape(-1).cpu(),
                                      torch.as_tensor(getattr(b, n)).float().reshape(-1).cpu()))
                     for n in ("scale", "zero", "qmin", "qmax"))
    block._ds_shared = ok
    block._ds_shared_key = key
    return ok


def _block_convs(block):
    return (block.conv1, block.conv2, block.conv3) if hasattr(block, "conv3") else (block.conv1, block.conv2)


def _block_forward(self, x, handoff=None, next_conv=None, bytes_only=False):
    """torchvision Bottleneck / BasicBlock forward with the ReLUs and the residual add folded into the convs.
    With `self.chain` the convs run as ONE engine chain: each conv hands its (ReLU'd) result to the next one already
    quantized, so the intermediates never exist as fp32 tensors (same bits as the unchained path).
    handoff / next_conv (used by _stage_forward): the block's input as the int8 workspace the previous block wrote for
    conv1, and the conv that will consume this block's output — the last epilogue then writes fp32 + int8."""
    convs = _block_convs(self)
    identity = None
    if handoff is not None and self.downsample is not None and _shortcut_shares_handoff(self):
        # the shortcut conv's quantizer has the same parameters as conv1's (both were calibrated on this very tensor): the
        # bytes the previous block wrote for conv1 are its input too — no quantizer pass over the fp32 tensor
        identity = _engine.load().quantconv2d_u8_nhwc(handoff, list(x.shape), self.downsample[0].chain_args())
    if identity is None:
        identity = x if self.downsample is None else self.downsample(x)
    if getattr(self, "chain", False) and _chainable(convs):
        qe = _engine.load()
        layers = [c.chain_args() for c in convs]
        if next_conv is not None and _chainable((next_conv,)):
            # bytes_only: the next block reads this block's output only as bytes (see _reads_only_bytes) — the fp32 store of
            # the last epilogue is skipped and the returned fp32 tensor is uninitialised (shape only)
            return qe.quantconv2d_chain(x.contiguous(), layers, residual=identity.contiguous(), input_handoff=handoff,
                                        emit_next=next_conv.chain_args(), emit_only=bool(bytes_only))
        out = qe.quantconv2d_chain(x.contiguous(), layers, residual=identity.contiguous(), input_handoff=handoff)
        return out if next_conv is None else (out, None)
    out = x
    for c in convs[:-1]:
        out = c(out)
    out = convs[-1](out, residual=identity)
    return out if next_conv is None else (out, None)


def _reads_only_bytes(block) -> bool:
    """Does `block` read its input only through conv1's int8 hand-off?  True for a chained down-sampling block whose
    shortcut conv shares that hand-off: conv1 and the shortcut are the only readers of the input (an identity shortcut would
    add the fp32 tensor itself).  The producing block may then skip its fp32 store."""
    return (block.downsample is not None and getattr(block, "chain", False) and _chainable(_block_convs(block))
            and _shortcut_shares_handoff(block))


def _resnet_forward_chained(self, x):
    """torchvision ResNet._forward_impl with the int8 hand-off running through ALL residual blocks, also across stage
    boundaries (the first conv of a down-sampling block consumes the previous stage's output like any other conv1;
    its 1x1 stride-2 shortcut conv reads the fp32 tensor)."""
    return _resnet_body_chained(self, _resnet_stem(self, x))


def _resnet_stem(self, x, out=None):
    """conv1 -> (folded bn) -> relu -> maxpool; `out`: where the pooled tensor goes (a batch slice when the input arrives in
    chunks of images — GraphedForward(stem_chunks=...))"""
    y = self.relu(self.bn1(self.conv1(x)))
    if out is not None and isinstance(self.maxpool, EngineMaxPool2d):
        return self.maxpool(y, out=out)
    y = self.maxpool(y)
    return y if out is None else out.copy_(y)


def _resnet_body_chained(self, x):
    """everything after the stem's pooling (see _resnet_forward_chained)"""
    blocks = [b for stage in (self.layer1, self.layer2, self.layer3, self.layer4) for b in stage]
    handoff = None
    for i, blk in enumerate(blocks):
        nxt = blocks[i + 1] if i + 1 < len(blocks) else None
        if nxt is not None and getattr(blk, "chain", False) and getattr(nxt, "chain", False):
            # (the engine skips the fp32 store only when it does write the hand-off: x is valid whenever handoff is None)
            x, handoff = blk(x, handoff, nxt.conv1, _reads_only_bytes(nxt))
        else:
            x = blk(x, handoff)
            handoff = None
    return self.fc(torch.flatten(self.avgpool(x), 1))


def _stage_forward(self, x):
    """nn.Sequential of residual blocks: block i's last conv also writes the int8 input of block i+1's conv1 (when that
    block has an identity shortcut, i.e. conv1 is the only quantizing consumer besides the fp32 residual path)."""
    blocks = list(self)
    handoff = None
    for i, blk in enumerate(blocks):
        nxt = blocks[i + 1] if i + 1 < len(blocks) else None
        if nxt is not None and getattr(blk, "chain", False) and getattr(nxt, "chain", False):
            x, handoff = blk(x, handoff, nxt.conv1)
        else:
            x = blk(x, handoff)
            handoff = None
    return x


def fuse_resnet_blocks(model, chain=False, cross_block=False):
    """For torchvision ResNets rebuilt with QuantConv2d: run `relu` and `+ identity` in the conv epilogues.
    Every fused conv applies relu(out [+ residual]) exactly where the original block does, so the network function is
    unchanged (tests/test_models_gpu.py asserts bit-identical logits).  chain=True additionally keeps the activations
    between the convs of a block quantized (engine.quantconv2d_chain); cross_block=True also lets a block's last conv
    write the int8 input of the next block's conv1 next to its fp32 result (measured slower on ResNet-50 at batch 256:
    the dual-output epilogue costs more than the quantizer launch it replaces — kept for tests and smaller batches)."""
    import types
    import torchvision.models.resnet as R
    for m in model.modules():
        if isinstance(m, R.Bottleneck) and all(isinstance(c, QuantConv2d) for c in (m.conv1, m.conv2, m.conv3)):
            for c in (m.conv1, m.conv2, m.conv3):
                c.fuse_relu = True
            m.chain = chain
            m.forward = types.MethodType(_block_forward, m)
        elif isinstance(m, R.BasicBlock) and all(isinstance(c, QuantConv2d) for c in (m.conv1, m.conv2)):
            m.conv1.fuse_relu = m.conv2.fuse_relu = True
            m.chain = chain
            m.forward = types.MethodType(_block_forward, m)
    if chain and cross_block and isinstance(model, R.ResNet):
        stages = (model.layer1, model.layer2, model.layer3, model.layer4)
        if all(hasattr(b, "chain") for stage in stages for b in stage):
            model.forward = types.MethodType(_resnet_forward_chained, model)
            # stem / body as separate callables: GraphedForward(stem_chunks=...) runs the stem per chunk of images
            model.qb_stem = types.MethodType(_resnet_stem, model)
            model.qb_body = types.MethodType(_resnet_body_chained, model)
        else:
            for stage in stages:
                if all(hasattr(b, "chain") for b in stage):
                    stage.forward = types.MethodType(_stage_forward, stage)
    if isinstance(model, R.ResNet) and isinstance(model.conv1, QuantConv2d):
        model.conv1.fuse_relu = True        # stem: conv1 -> (folded bn) -> relu -> maxpool
        model.relu = nn.Identity()
        if isinstance(model.maxpool, nn.MaxPool2d) and not model.maxpool.ceil_mode:
            model.maxpool = EngineMaxPool2d(model.maxpool)
    return model


class GraphedForward:
    """The packed forward captured as CUDA graphs ("CUDA-graph the stack").  The engine's ops are capture-safe
    once their caches are warm (no host synchronisation, tensor maps encoded on the host, allocations from the graph's
    pool), so one replay launches the whole network: what matters at small per-GPU batches (strong scaling, CIFAR-sized
    inputs), where ~100 launches through Python cost more than the kernels run.
    `n_buffers` input buffers, one graph each (the input address is baked into a graph): the H2D copy of step k+1 can fill
    buffer (k+1) % n while graph k % n runs.  Usage:  g = GraphedForward(model, example);  g.input(i).copy_(x);  y = g(i)"""

    def __init__(self, model, example: Tensor, n_buffers=2, warmup=2, stem_chunks=1):
        """stem_chunks > 1 (models with `qb_stem` / `qb_body`, see fuse_resnet_blocks): the stem (conv1 -> relu -> maxpool)
        is captured once per chunk of images and writes its slice of the pooled tensor, the rest of the network is one more
        graph — so the H2D copy of a batch can arrive in chunks and the stem of chunk c runs while chunk c + 1 is still on
        the wire (the first batch no longer waits for its whole input).  Images are independent: the logits are bit-identical
        to the one-graph forward.  replay_chunk(i, c) / replay_body(i); __call__ replays everything in order."""
        assert example.is_cuda, "CUDA graphs need CUDA tensors"
        self.model = model
        self.inputs = [torch.empty_like(example) for _ in range(n_buffers)]
        self.inputs[0].copy_(example)
        n = example.shape[0]
        chunked = stem_chunks > 1 and hasattr(model, "qb_stem") and hasattr(model, "qb_body") and n % stem_chunks == 0
        self.stem_chunks = stem_chunks if chunked else 1
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):                     # fills the engine's caches (prepared weights, ranges)
                y = model(self.inputs[0])
            if chunked:
                step = n // stem_chunks
                pooled_shape = model.qb_stem(self.inputs[0][:step]).shape[1:]
                for _ in range(max(warmup, 1)):
                    model.qb_stem(self.inputs[0][:step])
        torch.cuda.current_stream(example.device).wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graphs, self.outputs, self.chunk_graphs = [], [], []
        pool = None
        for x in self.inputs:
            if not chunked:
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                    out = model(x)
                pool = g.pool()                                    # the graphs run one after the other: share the memory pool
                self.graphs.append(g)
                self.outputs.append(out)
                self.chunk_graphs.append([])
                continue
            pooled = torch.empty((n,) + tuple(pooled_shape), device=x.device, dtype=x.dtype)
            cgs = []
            for c in range(stem_chunks):
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                    model.qb_stem(x[c * step:(c + 1) * step], out=pooled[c * step:(c + 1) * step])
                pool = g.pool()
                cgs.append(g)
            g = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(g, pool=pool):
                out = model.qb_body(pooled)
            pool = g.pool()
            self.chunk_graphs.append(cgs)
            self.graphs.append(g)
            self.outputs.append(out)

    def input(self, i=0) -> Tensor:
        return self.inputs[i % len(self.inputs)]

    def replay_chunk(self, i, c):
        """stem of chunk c of input buffer i (stem_chunks > 1)"""
        self.chunk_graphs[i % len(self.graphs)][c].replay()

    def replay_body(self, i=0) -> Tensor:
        self.graphs[i % len(self.graphs)].replay()
        return self.outputs[i % len(self.outputs)]

    def __call__(self, i=0) -> Tensor:
        """replays the forward of input buffer i on the current stream; the returned tensor is overwritten by the next replay"""
        for g in self.chunk_graphs[i % len(self.graphs)]:
            g.replay()
        return self.replay_body(i)


def reconstruct(model: nn.Module, w_setting=None, a_setting=None) -> nn.Module:
    """reference modelzoo/reconstruct.py:15-41, :94-132 for Conv2d(+BatchNorm2d): a conv directly followed (in
    child order) by a BatchNorm2d is folded and the BN becomes Identity; other children are visited recursively.
    nn.Linear -> QuantLinear (reconstruct.py:115-117)."""
    names = [n for n, _ in model.named_children()]
    mods = dict(model.named_children())
    skip = set()
    for i, name in enumerate(names):
        m = mods[name]
        if name in skip:
            setattr(model, name, nn.Identity())
            continue
        if isinstance(m, nn.Conv2d) and not isinstance(m, QuantConv2d):
            nxt = mods[names[i + 1]] if i + 1 < len(names) else None
            if isinstance(nxt, nn.BatchNorm2d):
                setattr(model, name, QuantConv2d(m, nxt, w_setting, a_setting))
                skip.add(names[i + 1])
            else:
                setattr(model, name, QuantConv2d(m, None, w_setting, a_setting))
        elif isinstance(m, nn.Linear) and not isinstance(m, QuantLinear):
            setattr(model, name, QuantLinear(m, w_setting, a_setting))
        else:
            reconstruct(m, w_setting, a_setting)
    return model


def quant_layers(model):
    return [m for m in model.modules() if isinstance(m, QuantConv2d)]


def set_mode(model, calibrating=False, quantized=False):
    """reference runner/ptq.py:51-63."""
    model.train(False)
    for m in model.modules():
        if hasattr(m, "calibrating"):
            m.calibrating = calibrating
        if isinstance(m, Quantizer):
            m.quant(quantized)


@torch.no_grad()
def calibrate(model, batch):
    """one PTQ calibration pass (runner/ptq.py:70-78 with the modules in calibrating mode)."""
    set_mode(model, calibrating=True, quantized=False)
    model(batch)
    set_mode(model, calibrating=False, quantized=True)


def pack(model):
    """the packing loop the reference keeps commented out (runner/ptq.py:106-114)."""
    for m in model.modules():
        if isinstance(m, (QuantConv2d, QuantLinear)):
            m.pack()
    # the packed model runs on the engine: its global average pool too (torch's generic reduction is 4x slower there)
    pool = getattr(model, "avgpool", None)
    if isinstance(pool, nn.AdaptiveAvgPool2d) and pool.output_size in (1, (1, 1)):
        model.avgpool = EngineGlobalAvgPool2d()
    return model
