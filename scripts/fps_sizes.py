import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic
dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for (B, N, m) in [(48, 512, 256), (48, 2048, 512), (48, 4096, 1024), (48, 8192, 2048), (48, 16384, 2048), (8, 65536, 2048), (148, 2048, 512), (296, 2048, 512)]:
    x = synthetic.uniform(B, N, 1).to(dev)
    ms = t(lambda: pointnet2_utils.furthest_point_sample(x, m))
    print(f"B={B} N={N} m={m}: {ms:.3f} ms  {ms*1e3/(m-1):.3f} us/round")
