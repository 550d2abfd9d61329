"""ncu driver: the three kNN paths (select k<=16 large Q, list small Q, split+merge few queries) and FPS at B=48."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic, misc
dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
gt = synthetic.pcn_gt(48, 16384, 303)
g = gt.to(dev)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512)
c64 = xyz[:, :64].contiguous()
for _ in range(2):
    knn_point(16, xyz, xyz)      # knn_select_kernel
    knn_point(16, xyz, c512)     # knn_kernel<1,4>
    knn_point(32, g, c64)        # knn_kernel<1,1> split + knn_merge_kernel
torch.cuda.synchronize()
print("ok")
