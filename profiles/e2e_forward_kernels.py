"""Per-kernel GPU time of one forward of the packed, chained ResNet-50 (256 images) -> profiles/r01_e2e_forward_kernels.md.
Run with QB200_PDL=0 so that kernel durations do not overlap:  QB200_PDL=0 python profiles/e2e_forward_kernels.py"""
import os, sys, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantize_b200 import models
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0')
net = models.build_packed('resnet50', 8, 8, calib_batch=8, device=dev, seed=0, fuse_blocks=True, chain_blocks=True, cross_block=True)
x = torch.randn(256, 3, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(3): net(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(x); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.OrderedDict()
for e in ev:
    a = agg.setdefault(e.name[:90], [0, 0.0]); a[0] += 1; a[1] += e.device_time
tot = sum(a[1] for a in agg.values())
print("packed ResNet-50 forward, 256 images, chained blocks: %d kernels, %.1f us of GPU time (torch profiler, one forward)" % (len(ev), tot))
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.3f |" % (k, a[0], a[1], a[1] / tot))
