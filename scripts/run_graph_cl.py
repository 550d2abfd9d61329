"""Timing: DynamicGraphAttention's gather / subtract / concat (channels last) at the AdaPoinTr token shapes, ours vs the
reference formulation on the same GPU, forward and forward+backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_formulation as tf
from imgadapointr_b200.models.Transformer_utils import graph_feature_cl, knn_point
dev = torch.device("cuda:0")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for (B, N, M, C, k) in [(48, 512, 512, 384, 8), (48, 256, 256, 384, 8), (48, 576, 576, 384, 8), (16, 512, 512, 384, 8)]:
    pos = torch.rand((B, M, 3), device=dev)
    idx = knn_point(k, pos, pos[:, :N].contiguous())
    q = torch.randn((B, N, C), device=dev, requires_grad=True)
    v = torch.randn((B, M, C), device=dev, requires_grad=True)
    go = torch.randn((B, N, k, 2 * C), device=dev)
    def ours_fb():
        q.grad = v.grad = None
        graph_feature_cl(q, v, idx).backward(go)
    def ref_fb():
        q.grad = v.grad = None
        tf.dynamic_graph_feature(q, v, idx).backward(go)
    with torch.no_grad():
        fo, fr = t(lambda: graph_feature_cl(q, v, idx)), t(lambda: tf.dynamic_graph_feature(q, v, idx))
    print("B=%d N=%d M=%d C=%d k=%d out %.0f MB: fwd ours %.3f ms ref %.3f ms | fwd+bwd ours %.3f ms ref %.3f ms" % (
        B, N, M, C, k, go.numel() * 4 / 1e6, fo, fr, t(ours_fb), t(ref_fb)))
