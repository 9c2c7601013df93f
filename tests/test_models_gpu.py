"""GPU: whole networks of BASELINE.json's configs — packed inference through the engine vs the fake-quant path
(the restated reference `_forward`, torch fp32 on the same GPU, TF32 off).

Bar: (1) every conv layer of the network, fed the SAME input, matches the reference's packed forward (float conv on
dequantized operands, quantconv2d.py:207-210) within 1e-3 relative — the north star's per-op tolerance;
(2) top-1 agreement 100 % on the synthetic batch;  (3) end-to-end logits within 1e-2 of the logit scale: a network
of fake-quantizers is discontinuous — a 1e-7 difference in one layer's output can move an activation across a
rounding boundary of the next quantizer (one full quantization step), so the per-op tolerance does not compose."""
import copy

import pytest
import torch

from quantize_b200 import host, models

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _compare(name, batch, w_bits, a_bits):
    dev = "cuda"
    model = models.build_quantized(name, w_bits, a_bits).to(dev)
    x = models.synthetic_batch(name, batch, device=dev)
    host.calibrate(model, x)
    with torch.no_grad():
        ref = model(x).clone()                       # fake-quant forward (quantconv2d.py:154-168)
    packed = host.pack(copy.deepcopy(model))
    layers = [m for m in packed.modules() if isinstance(m, (host.QuantConv2d, host.QuantLinear))]
    worst = [0.0]

    def check_layer(m, inp, out):
        # same input, the reference's packed forward on it (float conv on dequantized operands)
        m.use_engine = False
        want = m.forward(inp[0])            # .forward: no hooks, no recursion
        m.use_engine = True
        tol = 1e-3 * want.abs() + 1e-4 * want.abs().max()
        err = (out - want).abs()
        worst[0] = max(worst[0], float((err / (want.abs().max() + 1e-30)).max()))
        assert bool((err <= tol).all()), f"layer {m}: {float((err - tol).max())}"

    hooks = [m.register_forward_hook(check_layer) for m in layers]
    with torch.no_grad():
        out = packed(x)
    for h in hooks:
        h.remove()
    with torch.no_grad():
        for m in layers:
            m.use_engine = False
        out_ref_packed = packed(x)
    scale = ref.abs().max()
    # network level (the rigorous check is the per-layer one above): a value on a rounding boundary may flip by one
    # level in one path and not the other; one 4-bit level is 1/15 of a layer's range and, now that the classifier is
    # quantized too (QuantLinear), a flip in its input reaches the logits directly
    e2e_tol = 1e-2 if a_bits >= 8 else 1e-1
    assert (out - ref).abs().max() <= e2e_tol * scale, float((out - ref).abs().max() / scale)
    assert (out - out_ref_packed).abs().max() <= e2e_tol * scale
    assert torch.equal(out.argmax(1), ref.argmax(1))           # top-1 agreement 100 %
    assert torch.equal(out.argmax(1), out_ref_packed.argmax(1))
    return out


def test_resnet20_cifar_w8a8():
    _compare("resnet20", 32, 8, 8)


def test_resnet18_w8a8():
    _compare("resnet18", 8, 8, 8)


def test_resnet50_w8a8():
    _compare("resnet50", 4, 8, 8)


def test_resnet18_w4a4():
    _compare("resnet18", 4, 4, 4)


def test_mobilenet_v2_w4a8_depthwise_and_pointwise():
    _compare("mobilenet_v2", 4, 4, 8)


@pytest.mark.parametrize("name,batch", [("resnet18", 4), ("resnet50", 2)])
def test_fused_blocks_are_bit_identical(name, batch):
    """relu / residual-add folded into the conv epilogues == the separate torch ops, bit for bit."""
    model = models.build_packed(name, 8, 8, calib_batch=4)
    x = models.synthetic_batch(name, batch, device="cuda")
    with torch.no_grad():
        want = model(x)
        fused = host.fuse_resnet_blocks(copy.deepcopy(model))
        assert sum(m.fuse_relu for m in host.quant_layers(fused)) >= 16
        got = fused(x)
        # and with the engine off (torch ops everywhere) the fused forward is still the same network
        for m in host.quant_layers(fused):
            m.use_engine = False
        ref_path = fused(x)
    assert torch.equal(got, want)
    assert (ref_path - want).abs().max() <= 1e-2 * want.abs().max()
    assert torch.equal(ref_path.argmax(1), want.argmax(1))


@pytest.mark.parametrize("cross", [False, True], ids=["in-block", "cross-block"])
@pytest.mark.parametrize("name,batch,wb,ab", [("resnet18", 4, 8, 8), ("resnet50", 3, 8, 8), ("resnet50", 2, 4, 4),
                                              ("resnet20", 5, 8, 8)])
def test_chained_blocks_are_bit_identical(name, batch, wb, ab, cross):
    """int8 activations handed from one conv's epilogue to the next conv (quantconv2d_chain)
    give the same logits, bit for bit, as the layer-by-layer fp32 hand-over."""
    import quantize_b200.engine as E
    model = models.build_packed(name, wb, ab, calib_batch=4, fuse_blocks=True)
    x = models.synthetic_batch(name, batch, device="cuda")
    with torch.no_grad():
        want = model(x)
        chained = host.fuse_resnet_blocks(copy.deepcopy(model), chain=True, cross_block=cross)
        qe = E.load()
        chained(x)                    # (the first call prepares the copied weights)
        qe._launch_count_reset()
        got = chained(x)
        n_chain = qe._launch_count()
        qe._launch_count_reset()
        model(x)
        n_plain = qe._launch_count()
    assert torch.equal(got, want)
    if name != "resnet20":      # (its blocks are not torchvision's: nothing to chain, still must run)
        assert n_chain < n_plain      # the intermediate quantizer launches are gone


def test_chain_op_level_all_layouts(engine):
    """quantconv2d_chain == quantconv2d_float_input applied layer by layer, for every consumer workspace layout:
    NHWC (1x1 and strided 3x3 consumers), zero-padded NHWC (halo variant), ragged channel counts, zero points != 0,
    and a pair the engine cannot chain (depthwise consumer -> fp32 fallback inside the chain)."""
    from gpu_util import random_conv_case
    import numpy as np
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    # (C0, H, [ (K, R, stride, pad, groups), ... ])
    nets = [
        (64, 28, [(64, 1, 1, 0, 1), (64, 3, 1, 1, 1), (256, 1, 1, 0, 1)]),      # bottleneck, halo consumer (C=64, 28x28)
        (128, 14, [(128, 3, 2, 1, 1), (128, 3, 1, 1, 1)]),                       # strided producer, plain consumers
        (24, 9, [(40, 3, 1, 1, 1), (72, 1, 1, 0, 1), (24, 3, 1, 1, 1)]),         # ragged channels (Cp > C)
        (3, 33, [(64, 7, 2, 3, 1), (32, 3, 1, 1, 1)]),                           # im2col-rows producer (stem)
        (32, 10, [(32, 1, 1, 0, 1), (32, 3, 1, 1, 32), (16, 1, 1, 0, 1)]),       # depthwise in the middle: fp32 fallback
        (256, 7, [(64, 1, 2, 0, 1), (64, 1, 1, 0, 1)]),                          # strided 1x1 producer
    ]
    for algo in (0, 4):
        engine._set_conv_algo(algo)
        try:
            for seed, (C0, H, layers) in enumerate(nets):
                if algo == 4 and any(l[4] > 1 for l in layers):
                    continue              # (forcing the tensor-core variants excludes grouped convs)
                N = 3
                g = torch.Generator().manual_seed(seed)
                x = torch.randn(N, C0, H, H, generator=g).cuda()
                C, Hc = C0, H
                tuples, cur = [], x
                want = x
                for li, (K, R, stride, pad, groups) in enumerate(layers):
                    c = random_conv_case(100 * seed + li, N, C, Hc, Hc, K, R, stride, pad, groups)
                    relu = li % 2 == 0
                    # asymmetric per-tensor quantizer fitted to this layer's input (range/minmax.py:136-143)
                    lo_, hi_ = float(want.min()), float(want.max())
                    a_scale = torch.tensor([(hi_ - lo_) / 255.0], dtype=torch.float32).cuda()
                    a_zero = (torch.tensor([lo_], dtype=torch.float32).cuda() / a_scale) if li != 1 else torch.zeros(1).cuda()
                    tup = (t(c["packed"]), t(c["des"]), t(c["w_scale"]), torch.zeros(K).cuda(), t(c["bias"]), stride, pad,
                           a_scale, a_zero, 0, 255, relu)
                    tuples.append(tup)
                    want = engine.quantconv2d_float_input(want, *tup[:7], input_scale=a_scale, input_zero=a_zero,
                                                          input_qmin=0, input_qmax=255, fuse_relu=relu)
                    C, Hc = K, want.shape[2]
                got = engine.quantconv2d_chain(x, tuples)
                assert torch.equal(got, want), (algo, seed)
                res = torch.randn_like(want)
                last = tuples[-1]
                want_r = None
                # residual on the last layer
                prev = x
                for tup in tuples[:-1]:
                    prev = engine.quantconv2d_float_input(prev, *tup[:7], input_scale=tup[7], input_zero=tup[8], input_qmin=0,
                                                          input_qmax=255, fuse_relu=tup[11])
                want_r = engine.quantconv2d_float_input(prev, *last[:7], input_scale=last[7], input_zero=last[8], input_qmin=0,
                                                        input_qmax=255, residual=res, fuse_relu=last[11])
                assert torch.equal(engine.quantconv2d_chain(x, tuples, residual=res), want_r), (algo, seed, "residual")
        finally:
            engine._set_conv_algo(0)


def test_fused_tail_op_level(engine):
    """quantconv2d_float_input(..., residual=, fuse_relu=) == relu(op(...) + residual), both conv kernels."""
    from gpu_util import random_conv_case
    import numpy as np
    # (planes of 49 / 81 / 289 pixels are not a multiple of 4: the identity stream uses 4-byte copies there)
    for cfg in ((2, 64, 14, 14, 64, 3, 1, 1, 1), (2, 64, 28, 28, 256, 1, 1, 0, 1), (2, 3, 33, 33, 64, 7, 2, 3, 1),
                (2, 32, 8, 8, 32, 3, 1, 1, 32), (5, 128, 7, 7, 512, 1, 1, 0, 1), (3, 64, 7, 7, 128, 3, 1, 1, 1),
                (2, 64, 9, 9, 96, 1, 1, 0, 1)):
        N, C, H, W, K, R, stride, pad, groups = cfg
        c = random_conv_case(sum(cfg), N, C, H, W, K, R, stride, pad, groups)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        kw = dict(input_scale=torch.tensor([c["a_scale"]]).cuda(), input_zero=torch.tensor([c["a_zero"]]).cuda(),
                  input_qmin=0, input_qmax=255)
        args = (t(c["x"]), t(c["packed"]), t(c["des"]), t(c["w_scale"]), torch.zeros(K).cuda(), t(c["bias"]), stride, pad)
        base = engine.quantconv2d_float_input(*args, **kw)
        res = torch.randn_like(base)
        got = engine.quantconv2d_float_input(*args, **kw, residual=res, fuse_relu=True)
        assert torch.equal(got, torch.relu(base + res))
        assert torch.equal(engine.quantconv2d_float_input(*args, **kw, fuse_relu=True), torch.relu(base))
        assert torch.equal(engine.quantconv2d_float_input(*args, **kw, residual=res), base + res)


def test_batch_shards_equal_full_batch():
    """rows [k*B/G, (k+1)*B/G) of the full batch == the shard run on its own, bit for bit, for every
    conv of the network (the cuBLAS FC layer after the conv stack picks batch-dependent algorithms and is not ours)."""
    model = models.build_packed("resnet18", 8, 8, calib_batch=4)
    x = models.synthetic_batch("resnet18", 8, device="cuda")
    feats = []
    h = model.avgpool.register_forward_pre_hook(lambda m, inp: feats.append(inp[0].clone()))
    with torch.no_grad():
        full_logits = model(x)
        parts = [model(x[i:i + 2].contiguous()) for i in range(0, 8, 2)]
    h.remove()
    assert torch.equal(feats[0], torch.cat(feats[1:]))
    assert torch.allclose(full_logits, torch.cat(parts), rtol=1e-5, atol=1e-5)


def test_state_dict_round_trip_keeps_reference_format():
    """packed weights are the reference's byte stream + w_des (quantconv2d.py:186-191) and survive a checkpoint."""
    import io
    model = models.build_packed("resnet20", 8, 8, calib_batch=4)
    x = models.synthetic_batch("resnet20", 4, device="cuda")
    with torch.no_grad():
        want = model(x)
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)
    buf.seek(0)
    sd = torch.load(buf)
    l0 = host.quant_layers(model)[0]
    assert l0.weight.dtype == torch.uint8 and l0.weight.dim() == 1
    assert l0.w_des.tolist() == [8, 1, 16, 3, 3, 3]
    fresh = models.build_quantized("resnet20", 8, 8).to("cuda")
    host.set_mode(fresh, quantized=True)
    host.pack(fresh)
    fresh.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(fresh(x), want)


def test_fake_quant_forward_through_the_engine_kernel_is_bit_identical():
    """the unpacked (fake-quant) GPU forward with Quantizer.simulate on the engine's one-pass kernel
    == the same forward on the five torch kernels per quantizer."""
    model = models.build_quantized("resnet18", 8, 8, seed=0).cuda()
    host.calibrate(model, models.synthetic_batch("resnet18", 4, 1, "cuda"))
    x = models.synthetic_batch("resnet18", 3, device="cuda")
    with torch.no_grad():
        host.Quantizer.use_engine = False
        try:
            want = model(x)
        finally:
            host.Quantizer.use_engine = True
        import quantize_b200.engine as E
        qe = E.load()
        qe._launch_count_reset()
        got = model(x)
        assert qe._launch_count() >= 20          # one engine kernel per activation quantizer
    assert torch.equal(got, want)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json's own sizes (configs[1..3]): ResNet-18 bs 128, ResNet-50 bs 256, MobileNetV2 W4A8 bs 128
# ---------------------------------------------------------------------------------------------------------------------
def _full_size_parity(name, batch, w_bits, a_bits, chain):
    """Every conv / linear layer of the packed network at the BASELINE batch, fed the same input as the reference's packed
    forward (float conv on dequantized operands, quantconv2d.py:207-210): |err| <= 1e-3*|ref| + 1e-4*max|ref| per element,
    with the share of elements that needed the absolute floor reported (pure-relative 1e-3 fails only where the fp32
    reference itself cancels to ~0); top-1 over ALL images; chained (int8 hand-off) logits torch.equal to layer-by-layer."""
    import json
    import os
    dev = "cuda"
    model = models.build_quantized(name, w_bits, a_bits).to(dev)
    x = models.synthetic_batch(name, batch, device=dev)
    host.calibrate(model, x[:32])
    with torch.no_grad():
        ref = model(x).clone()                       # fake-quant forward of the whole batch
    packed = host.pack(copy.deepcopy(model))
    layers = [m for m in packed.modules() if isinstance(m, (host.QuantConv2d, host.QuantLinear))]
    stats = {"elements": 0, "needed_floor": 0, "worst_rel_to_layer_max": 0.0, "layers": len(layers)}

    def check_layer(m, inp, out):
        m.use_engine = False
        want = m.forward(inp[0])
        m.use_engine = True
        err = (out - want).abs()
        rel_ok = err <= 1e-3 * want.abs()
        tol = 1e-3 * want.abs() + 1e-4 * want.abs().max()
        stats["elements"] += err.numel()
        stats["needed_floor"] += int((~rel_ok).sum())
        stats["worst_rel_to_layer_max"] = max(stats["worst_rel_to_layer_max"], float(err.max() / (want.abs().max() + 1e-30)))
        assert bool((err <= tol).all()), f"layer {m}: {float((err - tol).max())}"

    hooks = [m.register_forward_hook(check_layer) for m in layers]
    with torch.no_grad():
        out = packed(x)
    for h in hooks:
        h.remove()
    agree = float((out.argmax(1) == ref.argmax(1)).float().mean())
    stats.update(model=name, batch=batch, w_bits=w_bits, a_bits=a_bits, top1_agreement=agree,
                 floor_share=stats["needed_floor"] / max(stats["elements"], 1),
                 logits_max_err_over_scale=float((out - ref).abs().max() / ref.abs().max()))
    if chain:
        with torch.no_grad():
            fused = host.fuse_resnet_blocks(copy.deepcopy(packed), chain=True, cross_block=True)
            got = fused(x)
        stats["chained_equal"] = bool(torch.equal(got, out))
        assert stats["chained_equal"]
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, f"parity_full_{name}_w{w_bits}a{a_bits}.json"), "w") as f:
            json.dump(stats, f)
    print("full-size parity:", stats)
    # a boundary flip in one quantizer moves one activation by a whole step, so logits are compared on the argmax: every
    # image whose reference top-1 margin exceeds the logit noise must agree, and in practice all do
    margin = ref.topk(2, dim=1).values
    clear = (margin[:, 0] - margin[:, 1]) > 1e-2 * ref.abs().max()
    assert bool((out.argmax(1) == ref.argmax(1))[clear].all())
    if w_bits >= 8 and a_bits >= 8:
        assert agree >= 0.98, agree
    return stats


def test_resnet18_w8a8_batch128_full_size():
    _full_size_parity("resnet18", 128, 8, 8, chain=True)


def test_resnet50_w8a8_batch256_full_size():
    _full_size_parity("resnet50", 256, 8, 8, chain=True)


def test_mobilenet_v2_w4a8_batch128_full_size():
    _full_size_parity("mobilenet_v2", 128, 4, 8, chain=False)


def test_graphed_forward_on_a_fresh_model_equals_eager():
    """host.GraphedForward right after build_packed (no eager call first): the prepared weights are created by the
    warm-up on a side stream and first used by the capture on another stream — no event query / wait may happen inside
    the capture.  Replays of both input buffers == the eager forward, bit for bit."""
    net = models.build_packed("resnet18", 8, 8, calib_batch=4, device="cuda", seed=0, fuse_blocks=True, chain_blocks=True,
                              cross_block=True)
    x = models.synthetic_batch("resnet18", 4, device="cuda")
    g = host.GraphedForward(net, torch.zeros_like(x), n_buffers=2)
    outs = []
    for i in range(2):
        g.input(i).copy_(x)
        outs.append(g(i).clone())
    with torch.no_grad():
        want = net(x)
    assert torch.equal(outs[0], want) and torch.equal(outs[1], want)


def test_shortcut_conv_reads_the_hand_off_of_conv1():
    """Down-sampling bottlenecks: the 1x1 / stride-2 shortcut conv's quantizer has the same parameters as conv1's (both
    calibrated on the block's input), so it consumes conv1's int8 hand-off (engine.quantconv2d_u8_nhwc) instead of
    quantizing the fp32 tensor again.  Logits stay bit-identical to the unchained forward; a block whose shortcut quantizer
    differs keeps the fp32 path (and the same logits as its own unchained forward)."""
    net = models.build_packed("resnet50", 8, 8, calib_batch=4, device="cuda", seed=1, fuse_blocks=True, chain_blocks=True,
                              cross_block=True)
    ref = models.build_packed("resnet50", 8, 8, calib_batch=4, device="cuda", seed=1, fuse_blocks=True, chain_blocks=False)
    x = models.synthetic_batch("resnet50", 3, device="cuda")
    with torch.no_grad():
        got, want = net(x), ref(x)
    assert torch.equal(got, want)
    shared = [b for stage in (net.layer2, net.layer3, net.layer4) for b in stage if getattr(b, "_ds_shared", False)]
    assert len(shared) == 3, "layer2.0 / layer3.0 / layer4.0 should share conv1's hand-off"
    assert not getattr(net.layer1[0], "_ds_shared", False)          # fed by the max pool: no hand-off to share
    # different quantizer parameters -> no sharing, still the right answer
    blk, rblk = net.layer3[0], ref.layer3[0]
    for b in (blk, rblk):
        with torch.no_grad():
            b.downsample[0].a_quantizer.scale.mul_(1.25)
    with torch.no_grad():                       # (the in-place change bumps the tensor's version: the cached answer is dropped)
        got2, want2 = net(x), ref(x)
    assert blk._ds_shared is False
    assert torch.equal(got2, want2) and not torch.equal(got2, got)


def test_graphed_forward_in_stem_chunks_is_bit_identical():
    """host.GraphedForward(stem_chunks=4): the stem (conv1 + max pool) replayed per quarter of the batch into slices of the
    pooled tensor, then the body — same logits, bit for bit, as the eager forward (images are independent)."""
    net = models.build_packed("resnet50", 8, 8, calib_batch=4, device="cuda", seed=2, fuse_blocks=True, chain_blocks=True,
                              cross_block=True)
    x = models.synthetic_batch("resnet50", 8, device="cuda")
    g = host.GraphedForward(net, torch.zeros_like(x), n_buffers=2, stem_chunks=4)
    assert g.stem_chunks == 4
    with torch.no_grad():
        want = net(x)
    for i in range(2):
        g.input(i).copy_(x)
        for c in range(4):
            g.replay_chunk(i, c)
        assert torch.equal(g.replay_body(i), want)
        g.input(i).zero_()
        g.input(i).copy_(x)
        assert torch.equal(g(i), want)
