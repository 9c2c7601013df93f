"""Generates the committed golden fixtures from the REFERENCE ITSELF (run in the build container only).

  python tests/golden/make_golden.py

  pack_kat.json   tpack/tunpack known-answer vectors produced by the reference's own compiled tpack
                  (oracle/_ref/quant_engine_ref.so = unmodified engine/kernels/tpack/tpack.cu, CPU path :140-190)
  conv_*.npz      per-layer fixtures produced by importing the reference's unmodified
                  modelzoo/modules/quantconv2d.py: calibrate -> fake-quant forward -> pack -> state_dict round
                  trip -> packed forward (quantconv2d.py:154-235), on seeded synthetic tensors.
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it, the CUDA path (tests/test_conv_gpu.py).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import build_ref  # noqa: E402
import refshim  # noqa: E402


def make_pack_kat(ref):
    rng = np.random.default_rng(1234)
    cases = []
    fixed = [
        (4, True, [-8, -7, -1, 0, 1, 7, 3, -4], "float32"),
        (8, True, [[-128, -1, 0, 127], [5, -5, 64, -64]], "float32"),
        (3, False, [0, 1, 2, 3, 4, 5, 6, 7, 7, 0], "float32"),
        (5, True, [-16, 15, -1, 0, 7], "float32"),
        (6, True, [-32, 31, -1, 0, 7, -9], "float32"),
        (4, True, [1, -2, 3, -4], "int8"),
        (4, True, [1, -2, 3, -4], "int32"),
        (4, True, [1, -2, 3, -4], "float16"),
        (4, True, [1, -2, 3, -4], "float64"),
        (1, False, [1, 0, 0, 1, 1, 1, 0, 1, 1], "float32"),
        (2, True, [-2, -1, 0, 1, 1, 0, -1], "float32"),
        (7, False, [0, 127, 64, 1, 2, 3, 100, 99, 98], "float32"),
        (8, False, [0, 255, 128, 127, 1], "float32"),
    ]
    for n_bits, sign, vals, dt in fixed:
        cases.append((n_bits, sign, np.array(vals), dt))
    for n_bits in range(1, 9):
        for sign in (False, True):
            lo, hi = (-(1 << (n_bits - 1)), (1 << (n_bits - 1)) - 1) if sign else (0, (1 << n_bits) - 1)
            for shape in ((37,), (3, 5, 7)):
                cases.append((n_bits, sign, rng.integers(lo, hi + 1, size=shape), "float32"))
    out = []
    for n_bits, sign, vals, dt in cases:
        x = torch.tensor(vals).to(getattr(torch, dt))
        packed, des = ref.tpack(x.contiguous(), n_bits, sign)
        back = ref.tunpack(packed, des)
        assert torch.equal(back.to(torch.int64), torch.tensor(vals).to(torch.int64))
        out.append({"n_bits": n_bits, "sign": bool(sign), "dtype": dt, "shape": list(x.shape),
                    "values": np.asarray(vals).reshape(-1).tolist(), "packed_hex": bytes(packed.numpy().tolist()).hex(),
                    "des": des.numpy().tolist(), "unpacked_dtype": str(back.dtype).replace("torch.", "")})
    with open(os.path.join(HERE, "pack_kat.json"), "w") as f:
        json.dump({"generator": "reference engine/kernels/tpack/tpack.cu compiled unmodified (oracle/build_ref.py)",
                   "cases": out}, f, indent=0)
    print("pack_kat.json:", len(out), "cases")


CONV_CASES = [
    # name, N, C, H, W, K, k, stride, pad, groups, w_bits, a_bits, relu_input, bn
    ("w8a8_k3s1p1_neg", 2, 16, 9, 11, 24, 3, 1, 1, 1, 8, 8, False, True),
    ("w8a8_k1s1p0_relu", 2, 32, 8, 8, 16, 1, 1, 0, 1, 8, 8, True, True),
    ("w4a8_k3s2p1", 2, 16, 12, 12, 16, 3, 2, 1, 1, 4, 8, False, True),
    ("w4a4_k7s2p3_stem", 1, 3, 20, 20, 8, 7, 2, 3, 1, 4, 4, False, False),
    ("w8a8_dw_k3s1p1", 2, 16, 10, 10, 16, 3, 1, 1, 16, 8, 8, False, True),
    ("w8a8_k1s2p0", 2, 24, 9, 9, 40, 1, 2, 0, 1, 8, 8, True, False),
    ("w6a6_k3s1p1_tensor", 1, 8, 7, 7, 8, 3, 1, 1, 1, 6, 6, False, False),
]


def make_conv_fixtures(mods):
    QuantConv2d = mods.QuantConv2d
    for (name, N, C, H, W, K, k, stride, pad, groups, wb, ab, relu_in, bn) in CONV_CASES:
        torch.manual_seed(abs(hash(name)) % (2 ** 31))
        torch.manual_seed(sum(ord(c) for c in name))
        conv = nn.Conv2d(C, K, k, stride, pad, groups=groups, bias=not bn)
        gran = "layer" if "tensor" in name else "channel"
        w_setting = dict(n_bits=wb, symmetric=True, signed=True, granularity=gran,
                         range={"name": "minmax", "percentile": 0.0})
        a_setting = dict(n_bits=ab, symmetric=False, granularity="layer",
                         range={"name": "maminmax", "percentile": 0.0, "momentum": 0.1})
        bn_folding = {}
        if bn:
            bn_folding = dict(running_mean=torch.randn(K), running_var=torch.rand(K) * 1.5 + 0.5,
                              weight=1 + 0.2 * torch.randn(K), bias=torch.randn(K), eps=1e-5)

        def build():
            params = {"weight": conv.weight.detach().clone(), "bias": None if bn else conv.bias.detach().clone()}
            return QuantConv2d(C, K, k, stride, pad, 1, groups, w_setting=w_setting, a_setting=a_setting,
                               bn_folding=dict(bn_folding), _parameters=params)

        m = build()
        x = torch.randn(N, C, H, W)
        if relu_in:
            x = torch.relu(x)
        with torch.no_grad():
            # calibration pass (runner/ptq.py:51-63: calibrating=True, quantizers in pass-through mode)
            m.calibrating = True
            m.w_quantizer.quant(False)
            m.a_quantizer.quant(False)
            m(x)
            # quantized evaluation (fake-quant path, quantconv2d.py:154-168)
            m.calibrating = False
            m.w_quantizer.quant(True)
            m.a_quantizer.quant(True)
            out_fake = m(x).clone()
            a_scale = m.a_quantizer.scale.detach().clone()
            a_zero = m.a_quantizer.zero.detach().clone()
            qmin, qmax = m.a_quantizer.qmin.clone(), m.a_quantizer.qmax.clone()
            q_x = (x / a_scale - a_zero).round().clamp(qmin, qmax)
            # pack -> state_dict -> load into a fresh module -> packed forward (quantconv2d.py:170-235)
            m.pack()
            sd = {k_: v.clone() for k_, v in m.state_dict().items()}
            m2 = build()
            m2.load_state_dict(sd)
            m2.a_quantizer.quant(True)
            out_packed = m2(x).clone()
        np.savez_compressed(
            os.path.join(HERE, f"conv_{name}.npz"),
            x=x.numpy(), w_packed=sd["weight"].numpy(), w_des=sd["w_des"].numpy(),
            w_scale=sd["w_scale"].numpy().reshape(-1), w_zero=sd["w_zero"].numpy().reshape(-1),
            bias=sd["bias"].numpy() if "bias" in sd and sd["bias"] is not None else np.zeros(0, np.float32),
            a_scale=a_scale.numpy().reshape(-1), a_zero=a_zero.numpy().reshape(-1),
            qmin=np.array([float(qmin)], np.float32), qmax=np.array([float(qmax)], np.float32),
            stride=np.array([stride]), pad=np.array([pad]), groups=np.array([groups]),
            q_x=q_x.numpy().astype(np.uint8), q_w=m2.weight.detach().numpy().astype(np.int8),
            out_fake=out_fake.numpy(), out_packed=out_packed.numpy())
        print(f"conv_{name}.npz  a_zero={float(a_zero):.3f} max|fake-packed|={float((out_fake - out_packed).abs().max()):.3e}")


if __name__ == "__main__":
    build_ref.build()
    ref = build_ref.load()
    make_pack_kat(ref)
    mods = refshim.load_reference(ref)
    make_conv_fixtures(mods)
