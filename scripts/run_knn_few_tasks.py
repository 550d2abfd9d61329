"""kNN calls whose tiles of 32 queries do not fill the machine: select kernel against the culled path with one task per
CTA (IAPT_KNN_CULL_SOLO_BELOW = task count below which the 4 warps of a CTA share ONE task's tile list from the start)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic, misc

dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False


def t(fn, it=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


gt = synthetic.pcn_gt(48, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512)
g2 = synthetic.pcn_gt(48, 2048, 5).to(dev)
gc512 = misc.fps(g2, 512)
g1 = synthetic.pcn_gt(48, 1024, 7).to(dev)
g4 = synthetic.pcn_gt(8, 4096, 6).to(dev)
cases = [("B=48 2048 x 512 padded partial", xyz, c512), ("B=48 2048 x 512 surface", g2, gc512)]
for B in (1, 2, 4, 8, 12, 16, 20, 24, 32):
    cases.append(("B=%d 2048 x 2048 surface" % B, g2[:B].contiguous(), None))
for B in (8, 16, 24):
    cases.append(("B=%d 1024 x 1024 surface" % B, g1[:B].contiguous(), None))
for B in (1, 4):
    cases.append(("B=%d 4096 x 4096 surface" % B, g4[:B].contiguous(), None))
for name, r, q in cases:
    q = r if q is None else q
    os.environ["IAPT_KNN_PRUNED"] = "2"
    ref = knn_point(16, r, q)
    line = "%-34s select %.4f" % (name, t(lambda: knn_point(16, r, q)))
    os.environ["IAPT_KNN_PRUNED"] = "1"
    for solo in (1, 1 << 30):
        os.environ["IAPT_KNN_CULL_SOLO_BELOW"] = str(solo)
        same = bool((knn_point(16, r, q) == ref).all())
        line += "  culled %s %.4f%s" % ("solo" if solo > 1 else "4/CTA", t(lambda: knn_point(16, r, q)), "" if same else " MISMATCH")
    print(line, flush=True)
