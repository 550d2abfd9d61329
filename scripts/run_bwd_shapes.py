"""Timing: Chamfer backward (C ABI through chamfer.backward) at the loss shapes, list path and scan-only path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200 import chamfer
from imgadapointr_b200.utils import synthetic
dev = torch.device("cuda:0")
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
gt = synthetic.pcn_gt(48, 16384, 303).to(dev)
for (n, m) in [(16384, 16384), (512, 16384), (2048, 2048), (64, 16384)]:
    p = (gt[:, torch.randperm(16384)[:n]] + 0.01 * torch.randn(48, n, 3, device=dev)).contiguous()
    g = gt[:, :m].contiguous()
    d1, d2, i1, i2 = chamfer.forward(p, g)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    print("B=48 %d x %d: backward %.4f ms" % (n, m, t(lambda: chamfer.backward(p, g, i1, i2, g1, g2))), flush=True)
# collapsed prediction: every gt point maps to one of 4 pred points
p = torch.rand(48, 4, 3, device=dev).repeat(1, 128, 1).contiguous()
d1, d2, i1, i2 = chamfer.forward(p, gt)
g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
print("B=48 512(collapsed) x 16384: backward %.4f ms" % t(lambda: chamfer.backward(p, gt, i1, i2, g1, g2)))
