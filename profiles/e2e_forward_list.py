import os, sys, torch
sys.path.insert(0, "/root/repo")
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
ev.sort(key=lambda e: e.time_range.start)
import re
out=[]
for e in ev:
    m=re.search(r"conv_umma_kernel<([^>]*)>", e.name)
    nm = "umma<"+m.group(1).replace("false","0").replace("true","1").replace(" ","")+">" if m else e.name.split("(")[0][-40:]
    out.append("%s:%.0f"%(nm,e.device_time))
print(" ".join(out))
