"""Runs two forwards of the packed, chained ResNet-50 (256 images) — the command profiled for
profiles/r01_launches_e2e_forward.csv:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -k regex:"conv_umma|act_quantize|maxpool|zero_pad|linear" -c 400 --csv --log-file launches.csv python profiles/e2e_forward_run.py
(the last 63 matching launches are the second forward)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantize_b200 import models
dev = torch.device('cuda:0')
net = models.build_packed('resnet50', 8, 8, calib_batch=8, device=dev, seed=0, fuse_blocks=True, chain_blocks=True, cross_block=True)
x = torch.randn(256, 3, 224, 224, device=dev)
with torch.no_grad():
    for _ in range(2):
        y = net(x)
torch.cuda.synchronize()
print(float(y.abs().max()))
