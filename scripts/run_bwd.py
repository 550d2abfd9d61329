import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200 import chamfer
from imgadapointr_b200.utils import synthetic
dev = torch.device("cuda:0")
pred, gt = synthetic.pcn_pair(48, 16384, 202)
p, g = pred.to(dev), gt.to(dev)
d1, d2, i1, i2 = chamfer.forward(p, g)
g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
for _ in range(3):
    chamfer.backward(p, g, i1, i2, g1, g2)
torch.cuda.synchronize()
from imgadapointr_b200 import _lib
print("max fan-in:", int(torch.bincount(i1[0].long(), minlength=16384).max()), int(torch.bincount(i2[0].long(), minlength=16384).max()))
