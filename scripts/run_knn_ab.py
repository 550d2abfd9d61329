"""A/B timing of the kNN paths at B=48, k=16 (CUDA events): the select kernel against the Hilbert-binned culled kernel over
its knobs (IAPT_KNN_SEEDS, IAPT_KNN_CULL_GROUP, IAPT_KNN_PROBES - read per call, so one process measures all of them);
every variant is also compared with the select kernel's indices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic, misc
from imgadapointr_b200 import _lib

dev = torch.device("cuda:0")
pointnet2_utils.furthest_point_sample.reuse_prefix = False
NB = int(os.environ.get("AB_B", "48"))
gt = synthetic.pcn_gt(NB, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512)
u = (torch.rand(NB, 2048, 3, device=dev) - 0.5).contiguous()
g2 = synthetic.pcn_gt(NB, 2048, 5).to(dev)
g4 = synthetic.pcn_gt(NB, 4096, 6).to(dev)
g1 = synthetic.pcn_gt(NB, 1024, 7).to(dev)
gc512 = misc.fps(g2, 512)
z = xyz.clone()
z[:, 1200:] = 0  # heavy zero padding: 848 of 2048 points at the origin in every cloud
cases = [("padded partial 2048 self", xyz, xyz), ("padded partial 2048 x 512 fps", xyz, c512), ("41% zeros 2048 self", z, z),
         ("uniform 2048 self", u, u), ("surface 2048 self", g2, g2), ("surface 2048 x 512 fps", g2, gc512), ("surface 4096 self", g4, g4), ("surface 1024 self", g1, g1)]


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


def mode(pruned, seeds=None, warps=2, probes=1, near=12):
    os.environ["IAPT_KNN_NEAR"] = str(near)
    os.environ["IAPT_KNN_PROBES"] = str(probes)
    os.environ["IAPT_KNN_PRUNED"] = pruned
    os.environ["IAPT_KNN_CULL_GROUP"] = str(warps)
    if seeds is None:
        os.environ.pop("IAPT_KNN_SEEDS", None)
    else:
        os.environ["IAPT_KNN_SEEDS"] = str(seeds)


seed_list = [int(s) for s in os.environ.get("AB_SEEDS", "4").split(",")]
near_list = [int(s) for s in os.environ.get("AB_NEAR", "12").split(",")]
probe_list = [int(s) for s in os.environ.get("AB_PROBES", "1").split(",")]
warp_list = [int(s) for s in os.environ.get("AB_GROUP", "1,4").split(",")]
for name, r, q in cases:
    mode("2")
    ref_idx = knn_point(16, r, q)
    line = "%-32s select %.4f" % (name, t(lambda: knn_point(16, r, q)))
    for s in seed_list:
        for w in warp_list:
            for pr in probe_list:
                for nr in near_list:
                    mode("1", s, w, pr, nr)
                    same = bool((knn_point(16, r, q) == ref_idx).all())
                    line += "  s%d/n%d/g%d/p%d %.4f%s" % (s, nr, w, pr, t(lambda: knn_point(16, r, q)), "" if same else " MISMATCH")
    print(line, flush=True)
