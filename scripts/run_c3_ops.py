"""Small driver for ncu: FPS 2048->512, kNN k=16 (2048x2048), graph feature fwd+bwd at the config-3 shapes (B=48)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models import base_blocks as bb
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic

dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
gt = synthetic.pcn_gt(48, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
xk = torch.randn((48, 8, 2048), device=dev, requires_grad=True)
xq = torch.randn((48, 8, 2048), device=dev, requires_grad=True)
for _ in range(3):
    pointnet2_utils.furthest_point_sample(xyz, 512)
    idx = knn_point(16, xyz, xyz)
    out = bb.graph_feature(xq, xk, idx)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
print("ok")
