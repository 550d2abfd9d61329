"""Developer timing harness (not the bench contract): CUDA-event timings of each op at the BASELINE configs."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    out = {}
    B = int(os.environ.get("B", 48))
    pred, gt = synthetic.pcn_pair(B, 16384, 202)
    p, g = pred.to(dev), gt.to(dev)
    t = timeit(lambda: chamfer.forward(p, g))
    pairs = B * 16384 * 16384
    out["chamfer_fwd_ms"] = t
    out["chamfer_fwd_pairs_per_s"] = pairs / (t[1] * 1e-3)
    out["chamfer_fwd_frac_fp32_peak_1965"] = 16 * pairs / (t[1] * 1e-3) / 74.45e12
    d1, d2, i1, i2 = chamfer.forward(p, g)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    out["chamfer_bwd_ms"] = timeit(lambda: chamfer.backward(p, g, i1, i2, g1, g2))
    pr = p.clone().requires_grad_(True)
    loss = ChamferDistanceL1()

    def step():
        pr.grad = None
        loss(pr, g).backward()

    out["chamfer_l1_fwd_bwd_ms"] = timeit(step)
    pointnet2_utils.furthest_point_sample.reuse_prefix = False  # time real launches
    gt4 = synthetic.pcn_gt(B, 16384, 303)
    xyz = synthetic.pcn_partial(gt4, 2048, 304).to(dev)
    out["fps_2048_512_ms"] = timeit(lambda: pointnet2_utils.furthest_point_sample(xyz, 512))
    out["knn_2048_2048_k16_ms"] = timeit(lambda: knn_point(16, xyz, xyz))
    c = xyz[:, :512].contiguous()
    out["knn_2048_512_k16_ms"] = timeit(lambda: knn_point(16, xyz, c))
    c64 = xyz[:, :64].contiguous()
    out["knn_16384_64_k32_ms"] = timeit(lambda: knn_point(32, g, c64))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
