"""ncu driver: channels-last graph feature forward + backward at the transformer shape (B=48, N=M=512, C=384, k=8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import graph_feature_cl, knn_point
dev = torch.device("cuda:0")
B, N, C, k = 48, 512, 384, 8
pos = torch.rand((B, N, 3), device=dev)
idx = knn_point(k, pos, pos)
q = torch.randn((B, N, C), device=dev, requires_grad=True)
v = torch.randn((B, N, C), device=dev, requires_grad=True)
go = torch.randn((B, N, k, 2 * C), device=dev)
for _ in range(2):
    q.grad = v.grad = None
    graph_feature_cl(q, v, idx).backward(go)
torch.cuda.synchronize()
print("ok")
