"""Randomised parity sweep on the GPU against the CPU oracle (tests infrastructure, not product code):
random batch / cloud sizes, duplicates, zero padding and exact-tie grids through chamfer forward / backward, FPS, kNN,
graph-feature backward and the auction EMD.  Usage: python scripts/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.emd import emd_module as emd
from imgadapointr_b200.models import base_blocks as bb
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils

dev = torch.device("cuda:0")
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
pointnet2_utils.furthest_point_sample.reuse_prefix = False


def cloud(B, N):
    kind = rng.integers(0, 5)
    if kind == 0:
        x = rng.random((B, N, 3), dtype=np.float32) - 0.5
    elif kind == 1:  # coarse grid: exact ties everywhere
        x = (np.floor((rng.random((B, N, 3)) - 0.5) * 8) / 8).astype(np.float32)
    elif kind == 2:  # zero-padded tail
        x = rng.random((B, N, 3), dtype=np.float32) - 0.5
        x[:, int(N * rng.uniform(0.3, 0.9)):] = 0.0
    elif kind == 4:  # collapsed: every point is one of a handful of positions (fan-ins of thousands, exact-sum path)
        centres = (rng.random((B, int(rng.integers(1, 6)), 3), dtype=np.float32) - 0.5)
        x = np.take_along_axis(centres, rng.integers(0, centres.shape[1], size=(B, N, 1)).repeat(3, axis=2), axis=1)
        x = np.ascontiguousarray(x.astype(np.float32))
    else:  # duplicates of a few points + wide dynamic range
        x = (rng.standard_normal((B, N, 3)) * 10 ** rng.uniform(-3, 2)).astype(np.float32)
        src = rng.integers(0, N, size=N // 3)
        x[:, : N // 3] = x[:, src]
    return np.ascontiguousarray(x)


t_end, n_cases, fails = time.time() + budget, 0, []
while time.time() < t_end:
    B = int(rng.integers(1, 7))
    N = int(rng.choice([1, 2, 31, 33, 64, 100, 255, 257, 700, 1024, 2048, 3000, 5000, 8193, 16384]))
    M = int(rng.choice([1, 3, 32, 65, 128, 500, 1000, 2048, 4097, 9000]))
    a, b = cloud(B, N), cloud(B, M)
    ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
    case = "B=%d N=%d M=%d" % (B, N, M)
    try:
        d1, d2, i1, i2 = chamfer.forward(ta, tb)
        od1, od2, oi1, oi2 = oracle.chamfer_forward(a, b)
        assert np.array_equal(d1.cpu().numpy(), od1) and np.array_equal(d2.cpu().numpy(), od2), "chamfer dist"
        assert np.array_equal(i1.cpu().numpy(), oi1) and np.array_equal(i2.cpu().numpy(), oi2), "chamfer idx"
        g1 = rng.standard_normal((B, N)).astype(np.float32)
        g2 = rng.standard_normal((B, M)).astype(np.float32)
        gx1, gx2 = chamfer.backward(ta, tb, i1, i2, torch.from_numpy(g1).to(dev), torch.from_numpy(g2).to(dev))
        o1, o2 = oracle.chamfer_backward(a, b, oi1, oi2, g1, g2)
        # 1e-5 of the gradient scale, plus the float32 accumulation bound of an ordered sum for points with a large
        # fan-in (hundreds of terms that may cancel; the reference's float atomicAdd has the same bound)
        ai1, ai2 = oi1.astype(np.int64), oi2.astype(np.int64)
        v1 = np.abs(2 * g1[..., None] * (a - np.take_along_axis(b, ai1[..., None], 1))).astype(np.float64)
        v2 = np.abs(2 * g2[..., None] * (b - np.take_along_axis(a, ai2[..., None], 1))).astype(np.float64)
        mag1, mag2 = v1.copy(), v2.copy()
        for bb_ in range(B):
            np.add.at(mag2[bb_], ai1[bb_], v1[bb_])
            np.add.at(mag1[bb_], ai2[bb_], v2[bb_])
        for got, want, mag in ((gx1, o1, mag1), (gx2, o2, mag2)):
            scale = max(float(np.abs(want).max()), 1e-30)
            err = np.abs(got.cpu().numpy() - want)
            assert bool((err <= 1e-5 * scale + 1e-6 * mag).all()), "chamfer grad"
        # the one-launch step must agree bit for bit with the separate entry points (same kernel, other phase set)
        up = torch.tensor([float(rng.uniform(0.1, 3.0))], device=dev)
        kind_l = int(rng.choice([chamfer.LOSS_L1, chamfer.LOSS_L2, chamfer.LOSS_L1_PM]))
        r = chamfer.step(ta, tb, kind_l, sums=True, grad1=True, grad2=(kind_l != chamfer.LOSS_L1_PM), grad_loss=up)
        assert torch.equal(r.dist1, d1) and torch.equal(r.idx1, i1) and torch.equal(r.dist2, d2) and torch.equal(r.idx2, i2), "step search"
        q1, q2 = chamfer.loss_backward(ta, tb, i1, i2, d1, d2, kind_l, up, True, kind_l != chamfer.LOSS_L1_PM)
        same = lambda x, y: (x is None and y is None) or torch.equal(torch.nan_to_num(x, nan=7.0), torch.nan_to_num(y, nan=7.0))
        assert same(r.grad1, q1) and same(r.grad2, q2), "step grad vs loss_backward"
        s64 = [d1.double().sum(1), d2.double().sum(1), d1.double().sqrt().sum(1), d2.double().sqrt().sum(1)]
        for c in range(4):
            assert torch.allclose(r.sums[:, c].double(), s64[c], rtol=2e-5, atol=1e-30), "step sums"
        if N >= 2:
            m = int(rng.integers(1, min(N, 600) + 1))
            skip = bool(rng.integers(0, 2))
            pointnet2_utils.furthest_point_sample.skip_near_origin = skip
            fi = pointnet2_utils.furthest_point_sample(ta, m)
            assert np.array_equal(fi.cpu().numpy(), oracle.fps(a, m, skip)), "fps m=%d skip=%d" % (m, skip)
            pointnet2_utils.furthest_point_sample.tie_order = "upstream"
            fu = pointnet2_utils.furthest_point_sample(ta, m)
            pointnet2_utils.furthest_point_sample.tie_order = "lowest"
            assert np.array_equal(fu.cpu().numpy(), oracle.fps_ex(a, m, skip, "upstream")[0]), "fps upstream m=%d" % m
        k = int(rng.integers(1, min(N, 64) + 1))
        os.environ["IAPT_KNN_PRUNED"] = str(int(rng.integers(0, 3)))  # default / forced pruned / forced select
        ki, kd = knn_point(k, ta, tb, return_dist=True)
        oi, od = oracle.knn(k, a, b)
        assert np.array_equal(ki.cpu().numpy(), oi) and np.array_equal(kd.cpu().numpy(), od), "knn k=%d" % k
        if N * k <= 60000:
            C = int(rng.integers(1, 9))
            xq = torch.randn((B, C, M), device=dev, requires_grad=True)
            xk = torch.randn((B, C, N), device=dev, requires_grad=True)
            go = torch.randn((B, 2 * C, M, k), device=dev)
            bb.graph_feature(xq, xk, ki).backward(go)
            aq, ak = xq.detach().double().requires_grad_(True), xk.detach().double().requires_grad_(True)
            feat = torch.gather(ak.unsqueeze(2).expand(B, C, M, N), 3, ki.unsqueeze(1).expand(B, C, M, k))
            torch.cat((feat - aq.unsqueeze(-1), aq.unsqueeze(-1).expand(B, C, M, k)), 1).backward(go.double())
            mag = float(go.abs().max()) * k * max(1, M // max(N, 1))  # size of the sums behind one gradient entry
            for got, want in ((xq.grad, aq.grad), (xk.grad, ak.grad)):
                assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-6 * mag, "graph grad"
        if n_cases % 5 == 0:  # auction EMD needs clouds of one size: reuse a, resample b
            n_e = int(rng.choice([1, 7, 64, 300, 700]))
            ea = np.ascontiguousarray(rng.random((B, n_e, 3), dtype=np.float32))
            eb = np.ascontiguousarray(rng.random((B, n_e, 3), dtype=np.float32))
            if rng.integers(0, 2):
                eb[:, n_e // 2:] = eb[:, : n_e - n_e // 2]  # duplicated targets: exact ties
            it_e = int(rng.integers(1, 40))
            ed, eas = emd.emdModule()(torch.from_numpy(ea).to(dev), torch.from_numpy(eb).to(dev), 0.005, it_e)
            od_, oa_ = oracle.emd_forward(ea, eb, 0.005, it_e)
            assert np.array_equal(eas.cpu().numpy(), oa_) and np.array_equal(ed.cpu().numpy(), od_), "emd n=%d it=%d" % (n_e, it_e)
    except AssertionError as e:
        fails.append((case, str(e)))
        print("FAIL", case, e, flush=True)
    n_cases += 1
pointnet2_utils.furthest_point_sample.skip_near_origin = True
print("cases", n_cases, "failures", len(fails))
sys.exit(1 if fails else 0)
