import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
from imgadapointr_b200.utils import synthetic
dev = torch.device("cuda:0")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
l1 = ChamferDistanceL1()
for (B, N, M) in [(48, 512, 16384), (48, 2048, 2048), (48, 2048, 16384), (48, 4096, 4096), (16, 16384, 16384)]:
    x = synthetic.uniform(B, N, 1).to(dev); y = synthetic.uniform(B, M, 2).to(dev)
    f = t(lambda: chamfer.forward(x, y))
    d1, d2, i1, i2 = chamfer.forward(x, y)
    bw = t(lambda: chamfer.backward(x, y, i1, i2, d1, d2))
    def fb():
        p = x.detach().requires_grad_(True); l1(p, y).backward()
    tot = t(fb)
    ideal = B * N * M / 4.17e12 * 1e3
    print(f"B={B} N={N} M={M}: fwd {f:.3f} ms  bwd {bw:.3f} ms  L1 fwd+bwd {tot:.3f} ms  (pair kernel at full rate would be {ideal:.3f} ms)")
