"""ncu driver: one auction EMD call at the test loop's shape (B = 1, n = 16384, eps 0.005, 100 iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.extensions.emd import emd_module as emd
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(16384)
x, y = torch.rand((1, 16384, 3), generator=g).to(dev), torch.rand((1, 16384, 3), generator=g).to(dev)
for _ in range(2):
    d, a = emd.emdModule()(x, y, 0.005, 100)
torch.cuda.synchronize()
print(float(torch.sqrt(d).mean()))
