"""Synthetic-weight CNNs of BASELINE.json's configs, rebuilt around host.QuantConv2d.

ResNet-18/50 and MobileNetV2 come from torchvision's constructors (what the reference's model zoo wraps,
modelzoo/cnns/resnet.py:10-21, cnns/mobilenet/__init__.py:11-16); ResNet-20 (CIFAR) is not in the reference zoo and
is defined here with conv -> bn child ordering so that BN folding applies.
Weights are random (no network for checkpoints): torch default init with seed 0, BN statistics perturbed so that
folding is non-trivial.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import host


class _BasicBlock20(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.down = None
        if stride != 1 or cin != cout:
            self.down = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, 0, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + (x if self.down is None else self.down(x)))


class ResNet20(nn.Module):
    """He et al. CIFAR ResNet-20: 3 stages x 3 basic blocks, 16/32/64 channels."""

    def __init__(self, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        blocks, cin = [], 16
        for cout, stride in ((16, 1), (32, 2), (64, 2)):
            for i in range(3):
                blocks.append(_BasicBlock20(cin, cout, stride if i == 0 else 1))
                cin = cout
        self.layers = nn.Sequential(*blocks)
        self.fc = nn.Linear(64, num_classes)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layers(x)
        return self.fc(F.adaptive_avg_pool2d(x, 1).flatten(1))


def _float_model(name):
    if name == "resnet20":
        return ResNet20()
    import torchvision
    return {"resnet18": torchvision.models.resnet18, "resnet50": torchvision.models.resnet50,
            "mobilenet_v2": torchvision.models.mobilenet_v2}[name]()


INPUT_HW = {"resnet20": 32, "resnet18": 224, "resnet50": 224, "mobilenet_v2": 224}


def perturb_bn(model, gen):
    """BN running stats / affine parameters away from the identity, so that folding changes the weights."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=gen) * 1.5 + 0.5)
            m.weight.data.copy_(1 + 0.2 * torch.randn(n, generator=gen))
            m.bias.data.copy_(torch.randn(n, generator=gen) * 0.1)


def build_quantized(name, w_bits=8, a_bits=8, seed=0):
    """float model (seeded random init) -> QuantConv2d layers (BN folded).  Not yet calibrated / packed."""
    torch.manual_seed(seed)
    model = _float_model(name)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        perturb_bn(model, gen)
    model.eval()
    w_setting = dict(host.DEFAULT_W, n_bits=w_bits)
    a_setting = dict(host.DEFAULT_A, n_bits=a_bits)
    return host.reconstruct(model, w_setting, a_setting)


def synthetic_batch(name, batch, seed=1, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    hw = INPUT_HW[name]
    return torch.randn(batch, 3, hw, hw, generator=g).to(device)


def build_packed(name, w_bits=8, a_bits=8, calib_batch=8, device="cuda", seed=0, fuse_blocks=False, chain_blocks=False, cross_block=False):
    """calibrate on a synthetic batch (one PTQ pass), pack every conv with the engine's tpack; ready for inference.
    fuse_blocks: fold ReLU / residual add of torchvision ResNet blocks into the conv epilogues (same function);
    chain_blocks: also hand the activations between the convs of a block over as int8 (engine.quantconv2d_chain)."""
    model = build_quantized(name, w_bits, a_bits, seed).to(device)
    host.calibrate(model, synthetic_batch(name, calib_batch, seed + 1, device))
    host.pack(model)
    if fuse_blocks or chain_blocks:
        host.fuse_resnet_blocks(model, chain=chain_blocks, cross_block=cross_block)
    return model


def conv_layer_specs(name, batch):
    """[{N, C, H, W, K, R, stride, pad, groups, relu_input}] of every conv in forward order (shapes traced on a
    random image; relu_input = the layer's input is non-negative, i.e. its calibrated zero point is 0)."""
    torch.manual_seed(0)
    model = _float_model(name).eval()
    specs = []

    def hook(m, inp, out):
        x = inp[0]
        specs.append(dict(N=batch, C=m.in_channels, H=x.shape[2], W=x.shape[3], K=m.out_channels, R=m.kernel_size[0],
                          stride=m.stride[0], pad=m.padding[0], groups=m.groups, relu_input=bool(x.min() >= 0)))

    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, nn.Conv2d)]
    with torch.no_grad():
        model(torch.randn(1, 3, INPUT_HW[name], INPUT_HW[name]))
    for h in hs:
        h.remove()
    return specs


def conv_stack_work(specs, w_bits=8):
    """algorithmic work of a conv stack under the op contract:
    ops = 2*N*K*P*Q*(C/g)*R*S;  bytes = 4*N*C*H*W + 4*N*K*P*Q + ceil(K*(C/g)*R*S*wb/8) + 12*K."""
    ops = 0
    nbytes = 0
    for s in specs:
        P = (s["H"] + 2 * s["pad"] - s["R"]) // s["stride"] + 1
        Q = (s["W"] + 2 * s["pad"] - s["R"]) // s["stride"] + 1
        cg = s["C"] // s["groups"]
        ops += 2 * s["N"] * s["K"] * P * Q * cg * s["R"] * s["R"]
        nbytes += 4 * s["N"] * s["C"] * s["H"] * s["W"] + 4 * s["N"] * s["K"] * P * Q + \
            (s["K"] * cg * s["R"] * s["R"] * w_bits + 7) // 8 + 12 * s["K"]
    return ops, nbytes
