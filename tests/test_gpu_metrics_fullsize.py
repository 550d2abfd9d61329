"""GPU: metrics drop-in (utils/metrics.py) against the float64 oracle, and size-independent properties at the full
BASELINE.json sizes (B=48 x 16384 x 16384; encoder shapes at B=48) where the CPU oracle would take minutes."""
import numpy as np
import pytest
import torch

import oracle
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic
from imgadapointr_b200.utils.metrics import MetricAccumulator, Metrics, per_cloud_metrics

pytestmark = pytest.mark.gpu


def test_metrics_get_matches_oracle(dev):
    pred, gt = synthetic.pcn_pair(1, 4096, 404, noise=0.006)
    f, cdl1, cdl2, emd = Metrics.get(pred.to(dev), gt.to(dev))
    counts, boundary = oracle.fscore_counts(pred.numpy(), gt.numpy(), 0.01)
    # F-Score: FP32 sqrt(dist) < th vs the reference's float64 open3d distances: only points within 1e-6 of the
    # threshold may flip -> |dF| <= 2 * boundary / N
    want_f = oracle.fscore(pred.numpy(), gt.numpy(), 0.01)
    assert abs(float(f) - want_f) <= 2.0 * max(int(boundary[0]), 1) / 4096 + 1e-7
    tf = oracle.torch_formulation
    assert abs(float(cdl1) - 1000 * float(tf.chamfer_l1(pred.double(), gt.double()))) <= 1e-5 * float(cdl1)
    assert abs(float(cdl2) - 1000 * float(tf.chamfer_l2(pred.double(), gt.double()))) <= 1e-5 * float(cdl2)
    assert float(emd) == 0.0 and f.device == cdl1.device == dev
    # per-item entry points of the reference class
    assert abs(float(Metrics._get_f_score(pred.to(dev), gt.to(dev))) - float(f)) < 1e-7
    assert abs(float(Metrics._get_chamfer_distancel1(pred.to(dev), gt.to(dev))) - float(cdl1)) <= 1e-5 * float(cdl1)


def test_metrics_ignore_zeros_and_batch(dev):
    gt = synthetic.pcn_gt(1, 2048, 11)
    pred = synthetic.pcn_partial(gt, 1024, 12)  # zero padded prediction: CD filters the zero-sum points, F does not
    n_zero = int((pred[0].sum(1) == 0).sum())
    assert n_zero > 0
    vals = Metrics.get_dict(pred.to(dev), gt.to(dev))
    keep = pred[0].sum(1) != 0
    tf = oracle.torch_formulation
    want1 = 1000 * float(tf.chamfer_l1(pred[:, keep].double(), gt.double()))
    assert abs(float(vals['CDL1']) - want1) <= 1e-5 * want1
    want_f = tf.fscore(pred, gt, 0.01)  # F-Score uses ALL points (metrics.py:67-90 has no zero filter)
    assert abs(float(vals['F-Score']) - want_f) <= 4.0 / 1024
    # batch > 1: mean of per-cloud F (metrics.py:73-77), CD over all B*N elements, no filtering
    p3, g3 = synthetic.pcn_pair(3, 1500, 13, noise=0.008)
    f, c1, c2, _ = Metrics.get(p3.to(dev), g3.to(dev))
    assert abs(float(f) - tf.fscore(p3, g3, 0.01)) <= 4.0 / 1500
    assert abs(float(c1) - 1000 * float(tf.chamfer_l1(p3.double(), g3.double()))) <= 1e-5 * float(c1)
    fb, c1b, c2b = per_cloud_metrics(p3.to(dev), g3.to(dev))
    for i in range(3):
        fi, c1i, c2i, _ = Metrics.get(p3[i:i + 1].to(dev), g3[i:i + 1].to(dev))
        assert abs(float(fb[i]) - float(fi)) < 1e-6 and abs(float(c1b[i]) - float(c1i)) <= 1e-5 * float(c1i)
        assert abs(float(c2b[i]) - float(c2i)) <= 1e-5 * float(c2i)
    acc = MetricAccumulator(2, dev)
    acc.update_batch(torch.tensor([0, 1, 1], device=dev), fb, c1b, c2b)
    assert torch.allclose(acc.overall()[1], ((c1b[0] + (c1b[1] + c1b[2]) / 2) / 2).double())


def test_fullsize_chamfer_properties(dev):
    """BASELINE config 2 at full size: B=48, 16384 x 16384."""
    pred, gt = synthetic.pcn_pair(48, 16384, 202)
    p, g = pred.to(dev), gt.to(dev)
    d1, d2, i1, i2 = chamfer.forward(p, g)
    # (1) every reported distance is exactly the direct-form distance to the reported neighbour, every index in range
    nn1 = torch.gather(g, 1, i1.long().unsqueeze(-1).expand(-1, -1, 3))
    nn2 = torch.gather(p, 1, i2.long().unsqueeze(-1).expand(-1, -1, 3))
    for x, nn, d in ((p, nn1, d1), (g, nn2, d2)):
        dx, dy, dz = (nn - x).unbind(-1)
        ref = torch.addcmul(torch.addcmul(dy * dy, dx, dx), dz, dz)  # fma(dz,dz,fma(dx,dx,dy*dy)) up to fusion
        assert float((ref - d).abs().max()) <= 2e-7 * float(d.max()) + 1e-12
    assert int(i1.min()) >= 0 and int(i1.max()) < 16384 and int(i2.min()) >= 0 and int(i2.max()) < 16384
    # (2) swapping the arguments swaps the outputs bit for bit
    e1, e2, j1, j2 = chamfer.forward(g, p)
    assert torch.equal(e1, d2) and torch.equal(e2, d1) and torch.equal(j1, i2) and torch.equal(j2, i1)
    # (3) permuting gt leaves dist1 untouched and maps idx1 through the permutation (no exact ties in this data)
    perm = torch.randperm(16384, generator=torch.Generator().manual_seed(1)).to(dev)
    f1, f2, k1, k2 = chamfer.forward(p, g[:, perm])
    assert torch.equal(f1, d1) and torch.equal(perm[k1.long()], i1.long())
    assert torch.equal(f2, d2[:, perm])
    # (4) a subset of the batch equals the reference-checked small run (batch independence)
    s1, s2, t1, t2 = chamfer.forward(p[5:7].contiguous(), g[5:7].contiguous())
    assert torch.equal(s1, d1[5:7]) and torch.equal(t2, i2[5:7])
    # (5) spot-check 3 clouds against the CPU oracle (1 s each)
    for b in (0, 47):
        od1, od2, oi1, oi2 = oracle.chamfer_forward(pred[b:b + 1].numpy(), gt[b:b + 1].numpy())
        assert np.array_equal(od1[0], d1[b].cpu().numpy()) and np.array_equal(oi2[0], i2[b].cpu().numpy())
    # (6) fused L1 loss + gradient: run-to-run bit reproducible, and the gradient moves pred towards gt
    losses, grads = [], []
    for _ in range(2):
        x = p.detach().clone().requires_grad_(True)
        loss = ChamferDistanceL1()(x, g)
        loss.backward()
        losses.append(loss.detach().clone())
        grads.append(x.grad.clone())
    assert torch.equal(losses[0], losses[1]) and torch.equal(grads[0], grads[1])
    stepped = (p - 1e-3 * grads[0] / grads[0].abs().max()).contiguous()
    assert float(ChamferDistanceL1()(stepped, g)) < float(losses[0])


def test_fullsize_encoder_geometry(dev):
    """BASELINE config 3 at full size: B=48 zero-padded partial clouds, FPS 2048->512 and the encoder's kNN shapes."""
    gt = synthetic.pcn_gt(48, 16384, 303)
    xyz = synthetic.pcn_partial(gt, 2048, 304)
    x = xyz.to(dev)
    fi = pointnet2_utils.furthest_point_sample(x, 512)
    assert np.array_equal(fi.cpu().numpy(), oracle.fps(xyz.numpy(), 512, True))  # 48 clouds: fast on the CPU
    assert all(len(set(r.tolist())) == 512 or (xyz[b].abs().sum(1) > 0).sum() < 512 for b, r in enumerate(fi.cpu()))
    idx, dist = knn_point(16, x, x, return_dist=True)
    assert idx.shape == (48, 2048, 16) and int(idx.min()) >= 0 and int(idx.max()) < 2048
    assert bool((dist[..., 1:] >= dist[..., :-1]).all())  # ascending distances
    tie = dist[..., 1:] == dist[..., :-1]
    assert bool((idx[..., 1:][tie] > idx[..., :-1][tie]).all())  # ascending index inside exact ties
    assert float(dist[..., 0].max()) == 0.0  # every point is its own (or a duplicate's) nearest neighbour
    oi, od = oracle.knn(16, xyz.numpy(), xyz.numpy())  # all 48 clouds: the select path and its tie handling, bit for bit
    assert np.array_equal(idx.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    i8 = knn_point(8, x, x)
    assert np.array_equal(i8.cpu().numpy(), oi[..., :8])  # k = 8 is a prefix of k = 16 under the (distance, index) order
    # the reference formulation on the GPU picks the same neighbour sets wherever its FP32 expansion cannot flip a tie
    D = -2 * torch.matmul(x, x.transpose(1, 2)) + (x ** 2).sum(-1)[:, :, None] + (x ** 2).sum(-1)[:, None, :]
    ref = torch.topk(D, 16, dim=-1, largest=False, sorted=False).indices
    srt = torch.sort(D, dim=-1).values
    clear = (srt[..., 16] - srt[..., 15]) > 1e-6
    assert float(clear.float().mean()) > 0.5
    assert torch.equal(torch.sort(ref, -1).values[clear], torch.sort(idx, -1).values[clear])
