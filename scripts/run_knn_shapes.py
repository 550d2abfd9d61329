"""ncu / timing driver: every kNN shape of the encoder (B=48) once per pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic, misc
dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
gt = synthetic.pcn_gt(48, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512)
c256 = misc.fps(c512, 256)
shapes = [(16, xyz, xyz), (16, xyz, c512), (16, c512, c512), (16, c512, c256), (8, c256, c256), (8, c256, c512), (8, c512, c512)]
for _ in range(2):
    for k, r, q in shapes:
        knn_point(k, r, q)
torch.cuda.synchronize()
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for k, r, q in shapes:
    print("k=%d R=%d Q=%d: %.4f ms" % (k, r.shape[1], q.shape[1], t(lambda: knn_point(k, r, q))))
print("sequence: %.4f ms" % t(lambda: [knn_point(k, r, q) for k, r, q in shapes]))
u = (torch.rand(48, 2048, 3, device=dev) - 0.5).contiguous()
g2 = synthetic.pcn_gt(48, 2048, 5).to(dev)
print("uniform self k=16 2048: %.4f ms" % t(lambda: knn_point(16, u, u)))
print("pcn surface (no padding) self k=16 2048: %.4f ms" % t(lambda: knn_point(16, g2, g2)))
g4 = synthetic.pcn_gt(48, 4096, 6).to(dev)
print("pcn surface self k=16 4096: %.4f ms" % t(lambda: knn_point(16, g4, g4)))
gc512 = misc.fps(g2, 512)
print("pcn surface (no padding) k=16 R=2048 Q=512: %.4f ms" % t(lambda: knn_point(16, g2, gc512)))
print("pcn surface (no padding) k=16 R=512 Q=512: %.4f ms" % t(lambda: knn_point(16, gc512, gc512)))
gq = g2[:, :256].contiguous()
print("pcn surface (no padding) k=16 R=2048 Q=256 (first 256 points): %.4f ms" % t(lambda: knn_point(16, g2, gq)))
