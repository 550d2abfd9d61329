"""ncu driver: one pruned-path kNN call (48 x 2048 x 2048, k = 16, uniform cloud, self-query)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
dev = torch.device("cuda:0")
u = (torch.rand(48, 2048, 3, device=dev) - 0.5).contiguous()
for _ in range(3):
    knn_point(16, u, u)
torch.cuda.synchronize()
