"""CPU: the host-side mirror (quantize_b200/host.py) against the reference's own modules (container only) and
its calibration arithmetic against the committed fixtures."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import refshim
from quantize_b200 import host, models

needs_ref = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")


def _conv_bn(C, K, k, stride, pad, groups=1, seed=0):
    torch.manual_seed(seed)
    conv = nn.Conv2d(C, K, k, stride, pad, groups=groups, bias=False)
    bn = nn.BatchNorm2d(K)
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(K))
        bn.running_var.copy_(torch.rand(K) * 1.5 + 0.5)
        bn.weight.copy_(1 + 0.2 * torch.randn(K))
        bn.bias.copy_(torch.randn(K))
    return conv, bn.eval()


@needs_ref
@pytest.mark.parametrize("cfg", [(16, 24, 3, 1, 1, 1, 8, 8), (8, 8, 3, 2, 1, 8, 4, 8), (12, 20, 1, 1, 0, 1, 4, 4)])
def test_layer_mirror_is_bit_identical_to_reference_module(cfg):
    C, K, k, stride, pad, groups, wb, ab = cfg
    mods = refshim.load_reference(refshim.oracle_engine_module())
    conv, bn = _conv_bn(C, K, k, stride, pad, groups)
    w_setting = dict(host.DEFAULT_W, n_bits=wb)
    a_setting = dict(host.DEFAULT_A, n_bits=ab)
    mine = host.QuantConv2d(conv, bn, w_setting, a_setting)
    ref = mods.QuantConv2d(C, K, k, stride, pad, 1, groups, w_setting=w_setting, a_setting=a_setting,
                           bn_folding=dict(running_mean=bn.running_mean, running_var=bn.running_var, weight=bn.weight.detach(),
                                           bias=bn.bias.detach(), eps=bn.eps),
                           _parameters={"weight": conv.weight.detach().clone(), "bias": None})
    assert torch.equal(mine.weight, ref.weight) and torch.equal(mine.bias, ref.bias)     # BN folding
    x1, x2 = torch.randn(2, C, 9, 9), torch.randn(2, C, 9, 9) * 2
    with torch.no_grad():
        for m in (mine, ref):
            m.calibrating = True
            m.w_quantizer.quant(False); m.a_quantizer.quant(False)
            m(x1); m(x2)                      # two batches: exercises the moving-average update
            m.calibrating = False
            m.w_quantizer.quant(True); m.a_quantizer.quant(True)
        for q in ("a_quantizer", "w_quantizer"):
            for p in ("scale", "zero", "qmin", "qmax"):
                assert torch.equal(getattr(getattr(mine, q), p), getattr(getattr(ref, q), p)), (q, p)
        assert torch.equal(mine(x1), ref(x1))                                            # fake-quant forward
        wi, ws, wz = mine.w_quantizer.pack(mine.weight)
        ri, rs, rz = ref.w_quantizer.pack(ref.weight)
        assert torch.equal(wi, ri) and torch.equal(ws, rs) and torch.equal(wz, rz)       # integers handed to tpack


@needs_ref
def test_resnet20_reconstruct_matches_reference_reconstruct():
    refshim.load_reference(refshim.oracle_engine_module())
    from modelzoo.reconstruct import reconstruct as ref_reconstruct
    from utils import Configs
    import os
    cfg = Configs()
    cfg.merge_from_yaml(os.path.join(refshim.REF_ROOT, "configs/runners/ptq/minmax/base.yaml"))
    cfg.freeze()
    torch.manual_seed(0)
    float_model = models.ResNet20()
    with torch.no_grad():
        models.perturb_bn(float_model, torch.Generator().manual_seed(1))
    float_model.eval()
    import copy
    ref_model = ref_reconstruct(copy.deepcopy(float_model), cfg.quant)
    # (the reference also rebuilds nn.Linear as QuantLinear, reconstruct.py:115-117; so does the mirror)
    mine = host.reconstruct(copy.deepcopy(float_model))
    assert len(host.quant_layers(mine)) == 21 and isinstance(mine.fc, host.QuantLinear)
    assert ref_model.fc.__class__.__name__ == "QuantLinear"
    x = torch.randn(4, 3, 32, 32)
    with torch.no_grad():
        host.calibrate(mine, x)
        for m in ref_model.modules():            # runner/ptq.py:51-63
            if hasattr(m, "calibrating"):
                m.calibrating = True
            if m.__class__.__name__ == "Quantizer":
                m.quant(False)
        ref_model(x)
        for m in ref_model.modules():
            if hasattr(m, "calibrating"):
                m.calibrating = False
            if m.__class__.__name__ == "Quantizer":
                m.quant(True)
        assert torch.equal(mine(x), ref_model(x))


def test_calibration_matches_fixture_parameters():
    """MinMax / MAMinMax arithmetic against the parameters the reference produced for the committed fixtures."""
    from conftest import load_conv_fixture
    f = load_conv_fixture("w8a8_k3s1p1_neg")
    q = host.Quantizer(**host.DEFAULT_A, flag="activation", n_channels=16, dim=4)
    q.calibrate(torch.from_numpy(f["x"]))
    assert np.array_equal(q.scale.detach().numpy().reshape(-1), f["a_scale"])
    assert np.array_equal(q.zero.detach().numpy().reshape(-1), f["a_zero"])
    assert float(q.qmin) == f["qmin"][0] and float(q.qmax) == f["qmax"][0]
    q.quant(True)
    assert np.array_equal(q.quantize_int(torch.from_numpy(f["x"])).detach().numpy().astype(np.uint8), f["q_x"])


def test_conv_specs_and_work_match_survey():
    specs = models.conv_layer_specs("resnet50", 256)
    ops, nbytes = models.conv_stack_work(specs)
    assert len(specs) == 53
    assert abs(ops / 256 / 1e9 - 8.1743) < 1e-3          # 8.1743 Gop / image
    assert abs(nbytes / 1e9 - 22.32) < 0.01              # 22.32 GB / batch-256 under the op contract
    assert not specs[0]["relu_input"] and all(s["relu_input"] for s in specs[1:])
    ops18, _ = models.conv_stack_work(models.conv_layer_specs("resnet18", 128))
    assert abs(ops18 / 128 / 1e9 - 3.6271) < 1e-3


@pytest.mark.parametrize("name", ["resnet18", "resnet50"])
def test_fused_and_chained_block_forwards_keep_the_network_function_on_cpu(name):
    """host.fuse_resnet_blocks rewires the block / stage / network forwards (ReLU and residual add folded into the
    convs, hand-off plumbing between blocks).  Without the engine (CPU, fake-quant mode) every branch must fall back to
    the torch ops and give the same output as the untouched module graph."""
    import copy
    from quantize_b200 import models
    torch.manual_seed(0)
    model = models.build_quantized(name, 8, 8, seed=0)
    host.calibrate(model, torch.randn(2, 3, 64, 64))
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        want = model(x)
        for kw in (dict(), dict(chain=True), dict(chain=True, cross_block=True)):
            fused = host.fuse_resnet_blocks(copy.deepcopy(model), **kw)
            got = fused(x)
            assert torch.equal(got, want), kw


def test_cpu_baseline_port_equals_the_module_fake_quant_forward():
    """oracle/fakequant.py (bench.py's CPU arm) is pinned to host.QuantConv2d._forward — itself pinned bit-for-bit to the
    reference module above: same quantizer parameters, same fake-quantized operands, same fp32 conv."""
    from oracle import fakequant
    torch.manual_seed(3)
    conv = nn.Conv2d(12, 20, 3, 1, 1, bias=True)
    x = torch.relu(torch.randn(4, 12, 10, 10))
    layer = host.QuantConv2d(conv, None, dict(host.DEFAULT_W, n_bits=8), dict(host.DEFAULT_A, n_bits=8))
    with torch.no_grad():
        layer.calibrating = True
        layer(x)
        layer.calibrating = False
        layer.w_quantizer.quant(True); layer.a_quantizer.quant(True)
        want = layer(x)
        aq = fakequant.minmax_asym(x, 8)
        wq = fakequant.minmax_sym_channel(conv.weight.detach(), 8)
        assert torch.equal(aq[0].reshape(-1), layer.a_quantizer.scale.reshape(-1)) and torch.equal(aq[1].reshape(-1), layer.a_quantizer.zero.reshape(-1))
        assert torch.equal(wq[0].reshape(-1), layer.w_quantizer.scale.reshape(-1))
        got = torch.nn.functional.conv2d(fakequant.fake_quant(x, *aq), fakequant.fake_quant(conv.weight.detach(), *wq), conv.bias, 1, 1)
        assert torch.equal(got, want)


@needs_ref
@pytest.mark.parametrize("cfg", [dict(symmetric=False, granularity="layer", percentile=0.01),
                                 dict(symmetric=True, granularity="layer", percentile=0.001),
                                 dict(symmetric=True, granularity="channel", percentile=0.02),
                                 dict(symmetric=False, granularity="channel", percentile=0.0)])
def test_percentile_ranges_mirror_the_reference(cfg):
    """range/minmax.py:78-84, :92-98 (kthvalue ranges), incl. the state update over two batches."""
    refshim.load_reference(refshim.oracle_engine_module())
    from modelzoo.modules.range.minmax import MinMax as RefMinMax
    torch.manual_seed(1)
    for flag in ("weight", "activation"):
        mine, ref = host.MinMax(n_bits=8, signed=True, **cfg), RefMinMax(n_bits=8, signed=True, **cfg)
        for i in range(2):
            x = torch.randn(6, 10, 5, 5) * (i + 1)
            a, b = mine(flag, x), ref(flag, x)
            for u, v in zip(a, b):
                assert (torch.equal(u, v) if torch.is_tensor(u) else u == v)


def test_oracle_reductions_equal_torch():
    import oracle
    torch.manual_seed(2)
    x = torch.randn(5, 7, 3, 3)
    for gran, flag in ((0, "weight"), (1, "weight"), (1, "activation")):
        rows = x.reshape(1, -1) if gran == 0 else (x.transpose(0, 1).flatten(1) if flag == "activation" else x.flatten(1))
        lo, hi = oracle.minmax(x.numpy(), gran, flag, False)
        assert np.array_equal(np.atleast_1d(lo), rows.min(dim=1)[0].numpy()) and np.array_equal(np.atleast_1d(hi), rows.max(dim=1)[0].numpy())
        lo, hi = oracle.minmax(x.numpy(), gran, flag, True)
        assert np.array_equal(np.atleast_1d(hi), rows.abs().max(dim=1)[0].numpy()) and not np.any(lo)
        for k in (1, 3, rows.shape[1]):
            assert np.array_equal(np.atleast_1d(oracle.kthvalue(x.numpy(), k, gran, flag)), rows.kthvalue(k, dim=1)[0].numpy())
            assert np.array_equal(np.atleast_1d(oracle.kthvalue(x.numpy(), k, gran, flag, True)), rows.abs().kthvalue(k, dim=1)[0].numpy())
