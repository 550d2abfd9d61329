import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.utils import synthetic, misc
dev = torch.device("cuda:0")
gt = synthetic.pcn_gt(48, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512); c256 = misc.fps(c512, 256)
for _ in range(2):
    for (k, R, Q) in [(16, xyz, xyz), (16, xyz, c512), (16, c512, c512), (8, c256, c256), (8, c256, c512)]:
        idx = knn_point(k, R, Q)
torch.cuda.synchronize()
import ctypes
flags_total = 0
print("ok")
