"""Timing: serving-size batches (B = 1..8) of the encoder geometry ops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic, misc
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
def t(fn, it=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for B in (1, 2, 4, 8):
    gt = synthetic.pcn_gt(B, 16384, 303)
    xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
    g = gt.to(dev)
    c512 = misc.fps(xyz, 512)
    l1 = ChamferDistanceL1()
    with torch.no_grad():
        print("B=%d  fps 2048->512 %.0f us | knn16 2048x2048 %.0f us | knn16 2048->512 %.0f us | knn8 512x512 %.0f us | CD-L1 fwd 16384^2 %.0f us" % (
            B, t(lambda: pointnet2_utils.furthest_point_sample(xyz, 512)), t(lambda: knn_point(16, xyz, xyz)),
            t(lambda: knn_point(16, xyz, c512)), t(lambda: knn_point(8, c512, c512)), t(lambda: l1(g, g.flip(1)))))
