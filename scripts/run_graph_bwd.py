"""Timing driver: get_graph_feature backward at the DGCNN grouper shapes (B=48), each gradient alone and both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200.models import base_blocks as bb
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.utils import synthetic, misc
dev = torch.device("cuda:0")
gt = synthetic.pcn_gt(48, 16384, 303)
xyz = synthetic.pcn_partial(gt, 2048, 304).to(dev)
c512 = misc.fps(xyz, 512); c256 = misc.fps(c512, 256)
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
shapes = [(8, xyz, xyz, 16), (32, xyz, c512, 16), (64, c512, c512, 16), (64, c512, c256, 16)]
for C, rk, rq, k in shapes:
    idx = knn_point(k, rk, rq)
    Nk, Nq = rk.shape[1], rq.shape[1]
    for gq_on, gk_on in [(True, False), (False, True), (True, True)]:
        xq = torch.randn((48, C, Nq), device=dev, requires_grad=gq_on)
        xk = torch.randn((48, C, Nk), device=dev, requires_grad=gk_on)
        out = bb.graph_feature(xq, xk, idx)
        go = torch.randn_like(out)
        ms = t(lambda: torch.autograd.grad(out, [x for x in (xq, xk) if x.requires_grad], go, retain_graph=True))
        print("C=%d Nk=%d Nq=%d k=%d gq=%d gk=%d: %.4f ms  (grad_out %.0f MB)" % (C, Nk, Nq, k, gq_on, gk_on, ms, go.numel() * 4 / 1e6))
    f = t(lambda: bb.graph_feature(xq, xk, idx))
    print("   forward %.4f ms" % f)
