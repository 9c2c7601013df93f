"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement of the reference's module-level fake-quant conv path.

This is "the reference's CPU fake-quant conv path" of BASELINE.md §4: QuantConv2d._forward
(reference modelzoo/modules/quantconv2d.py:154-168) = Quantizer.simulate on the activations and on the weights
(modelzoo/modules/quantizer.py:196-226: q = clamp(round(x/s - z)); (q + z) * s) followed by an fp32 F.conv2d
(quantconv2d.py:166-168 -> nn.Conv2d._conv_forward, oneDNN on the host).  The reference package itself cannot travel
to the GPU box, so bench.py's cpu_baseline / `--impl reference` legs time this port (kind: "port").  tests/test_host_cpu.py
pins the same arithmetic bit-for-bit against the unmodified reference modules in the build container.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this file.
"""
import time

import torch
import torch.nn.functional as F


def fake_quant(x, scale, zero, qmin, qmax):
    """quantizer.py:215-218 (unpacked mode)."""
    q = (x / scale - zero).round().clamp(qmin, qmax)
    return (q + zero).mul(scale)


def minmax_asym(x, n_bits):
    """range/minmax.py:136-143 on one batch."""
    qmax = float((1 << n_bits) - 1)
    xmin, xmax = x.min(), x.max()
    scale = (xmax - xmin) / qmax
    return scale, xmin / scale, 0.0, qmax


def minmax_sym_channel(w, n_bits):
    """range/minmax.py:123-135, granularity 'channel', signed."""
    qmax, qmin = (1 << (n_bits - 1)) - 1, -(1 << (n_bits - 1))
    scale = w.flatten(1).abs().max(dim=1)[0] / (float(qmax - qmin - 1) / 2)
    return scale.view(-1, 1, 1, 1), torch.zeros_like(scale).view(-1, 1, 1, 1), float(qmin), float(qmax)


class FakeQuantConvStack:
    """The conv layers of a CNN (specs from quantize_b200.models.conv_layer_specs — plain dicts, passed in by the
    caller so that this file imports nothing from the product), each with its own synthetic input."""

    def __init__(self, specs, batch, w_bits=8, a_bits=8, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.layers = []
        for s in specs:
            x = torch.randn(batch, s["C"], s["H"], s["W"], generator=g)
            if s["relu_input"]:
                x = torch.relu(x)
            cg = s["C"] // s["groups"]
            w = torch.randn(s["K"], cg, s["R"], s["R"], generator=g) * (2.0 / (cg * s["R"] * s["R"])) ** 0.5
            b = torch.randn(s["K"], generator=g) * 0.1
            self.layers.append(dict(x=x, w=w, b=b, aq=minmax_asym(x, a_bits), wq=minmax_sym_channel(w, w_bits),
                                    stride=s["stride"], pad=s["pad"], groups=s["groups"]))
        self.batch = batch

    @torch.no_grad()
    def step(self):
        out = None
        for L in self.layers:
            xq = fake_quant(L["x"], *L["aq"])
            wq = fake_quant(L["w"], *L["wq"])
            out = F.conv2d(xq, wq, L["b"], L["stride"], L["pad"], 1, L["groups"])
        return out

    def time_steps(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / max(steps, 1)
