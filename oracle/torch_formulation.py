"""The reference's PyTorch CPU formulation of the hot path (BASELINE.json north_star).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  These run on the host cores; bench.py times them as
the ``cpu_baseline`` and as the ``--impl reference`` arm.

* Chamfer: per cloud ``torch.cdist(a, b) ** 2`` -> ``min`` both ways -> the reductions of
  extensions/chamfer_dist/__init__.py:35-44 (L2), :71-84 (L1), :93-103 (L1_PM); backward by autograd.
* kNN: ``square_distance`` + ``torch.topk(largest=False, sorted=False)`` restated verbatim from
  models/Transformer_utils.py:19-51 (that module itself cannot be imported here: it needs timm and pointnet2_ops).
* FPS: Python loop with the upstream pointnet2_ops semantics (start 0, temp=1e10, optional near-origin skip).
* F-Score: float64 ``torch.cdist`` NN both ways (utils/metrics.py:67-90; open3d is absent).
"""
from __future__ import annotations

import torch


def square_distance(src, dst):
    """models/Transformer_utils.py:32-51 (same op order, in place)."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def knn_point(nsample, xyz, new_xyz):
    """models/Transformer_utils.py:19-30."""
    sqrdists = square_distance(new_xyz, xyz)
    _, group_idx = torch.topk(sqrdists, nsample, dim=-1, largest=False, sorted=False)
    return group_idx


def chamfer_nn(xyz1, xyz2):
    """dist1, dist2, idx1, idx2 via cdist+min, one cloud at a time (a 16384^2 matrix is 1 GB)."""
    d1s, d2s, i1s, i2s = [], [], [], []
    for a, b in zip(xyz1, xyz2):
        d = torch.cdist(a[None], b[None])[0] ** 2
        d1, i1 = d.min(1)
        d2, i2 = d.min(0)
        d1s.append(d1), d2s.append(d2), i1s.append(i1), i2s.append(i2)
    return torch.stack(d1s), torch.stack(d2s), torch.stack(i1s), torch.stack(i2s)


def chamfer_l2(xyz1, xyz2):
    d1, d2, _, _ = chamfer_nn(xyz1, xyz2)
    return d1.mean() + d2.mean()


def chamfer_l1(xyz1, xyz2):
    d1, d2, _, _ = chamfer_nn(xyz1, xyz2)
    return (d1.sqrt().mean() + d2.sqrt().mean()) / 2


def chamfer_l1_fwd_bwd(xyz1, xyz2):
    """ChamferDistanceL1(pred, gt).backward() with autograd through cdist/min/sqrt/mean; returns (loss, grad_xyz1)."""
    xyz1 = xyz1.detach().clone().requires_grad_(True)
    loss = chamfer_l1(xyz1, xyz2)
    loss.backward()
    return loss.detach(), xyz1.grad


def fps_loop(xyz, npoint, skip_near_origin=True):
    """Iterative FPS, whole batch vectorised per round (temp=1e10, start 0, lowest-index argmax)."""
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, dtype=xyz.dtype)
    if skip_near_origin:
        mag = (xyz * xyz).sum(-1)
        skip = mag.double() <= 1e-3
    else:
        skip = torch.zeros(B, N, dtype=torch.bool)
    idx = torch.zeros(B, npoint, dtype=torch.int64)
    old = torch.zeros(B, dtype=torch.int64)
    ar = torch.arange(B)
    for j in range(1, npoint):
        d = ((xyz - xyz[ar, old][:, None, :]) ** 2).sum(-1)
        temp = torch.minimum(temp, d)
        cand = torch.where(skip, torch.full_like(temp, -1.0), temp)
        old = torch.argmax(cand, dim=1)
        idx[:, j] = old
    return idx.to(torch.int32)


def fscore(pred, gt, th=0.01):
    """utils/metrics.py:67-90 with the KD-tree replaced by a float64 cdist (same distances)."""
    fs = []
    for p, g in zip(pred, gt):
        d = torch.cdist(p[None].double(), g[None].double())[0]
        d1 = d.min(1).values
        d2 = d.min(0).values
        precision = float((d1 < th).sum()) / float(len(d1))
        recall = float((d2 < th).sum()) / float(len(d2))
        fs.append(2 * recall * precision / (recall + precision) if recall + precision else 0.0)
    return sum(fs) / len(fs)


def get_graph_feature(k, coor_q, x_q, coor_k, x_k, idx=None):
    """DGCNN_Grouper.get_graph_feature restated (models/base_blocks.py:565-588): kNN under no_grad, flat gather of the
    neighbour features, cat(feature - x_q, x_q) -> [B, 2C, Nq, k].  ``idx`` ([B,Nq,k]) overrides the kNN result."""
    batch_size, num_dims, num_points_k = x_k.shape
    num_points_q = x_q.size(2)
    with torch.no_grad():
        if idx is None:
            idx = knn_point(k, coor_k.transpose(-1, -2).contiguous(), coor_q.transpose(-1, -2).contiguous())
        idx = idx.transpose(-1, -2).contiguous()
        idx_base = torch.arange(0, batch_size, device=x_q.device).view(-1, 1, 1) * num_points_k
        idx = (idx + idx_base).view(-1)
    xk = x_k.transpose(2, 1).contiguous()
    feature = xk.view(batch_size * num_points_k, -1)[idx, :]
    feature = feature.view(batch_size, k, num_points_q, num_dims).permute(0, 3, 2, 1).contiguous()
    xq = x_q.view(batch_size, num_dims, num_points_q, 1).expand(-1, -1, -1, k)
    return torch.cat((feature - xq, xq), dim=1)


def three_nn(unknown, known):
    """pointnet2_ops three_nn restated: 3 nearest known points per unknown point, (sqrt distance, index), ascending by
    (distance, index) (upstream scans in ascending index with strict `<`)."""
    d2 = ((unknown[:, :, None, :] - known[:, None, :, :]) ** 2).sum(-1)
    srt, order = torch.sort(d2, dim=-1, stable=True)
    return srt[..., :3].sqrt(), order[..., :3].to(torch.int32)


def three_interpolate(features, idx, weight):
    """out[b,c,n] = sum_j weight[b,n,j] * features[b,c,idx[b,n,j]]."""
    B, C, M = features.shape
    g = torch.gather(features[:, :, None, :].expand(-1, -1, idx.size(1), -1), 3, idx.long()[:, None].expand(-1, C, -1, -1))
    return (g * weight[:, None]).sum(-1)


def grouping_operation(features, idx):
    """out[b,c,p,s] = features[b,c,idx[b,p,s]]."""
    B, C, N = features.shape
    return torch.gather(features[:, :, None, :].expand(-1, -1, idx.size(1), -1), 3, idx.long()[:, None].expand(-1, C, -1, -1))


def ball_query(radius, nsample, xyz, new_xyz):
    """pointnet2_ops ball_query restated (first nsample indices within radius, padded with the first hit, zeros if none)."""
    B, P, _ = new_xyz.shape
    out = torch.zeros((B, P, nsample), dtype=torch.int32)
    d2 = ((new_xyz[:, :, None, :] - xyz[:, None, :, :]) ** 2).sum(-1)
    for b in range(B):
        for p in range(P):
            hits = torch.nonzero(d2[b, p] < radius * radius).flatten()[:nsample]
            if len(hits):
                out[b, p, :] = int(hits[0])
                out[b, p, :len(hits)] = hits.to(torch.int32)
    return out


def dynamic_graph_feature(q, v, idx):
    """The gather / expand / subtract / concat of DynamicGraphAttention.forward (models/Transformer_utils.py:814-816):
    q [B,N,C], v [B,M,C], idx [B,N,k] -> [B,N,k,2C]."""
    B, N, C = q.shape
    k = idx.size(-1)
    local_v = torch.gather(v, 1, idx.reshape(B, -1, 1).expand(-1, -1, C)).reshape(B, N, k, C)
    qe = q.unsqueeze(-2).expand(-1, -1, k, -1)
    return torch.cat((local_v - qe, qe), dim=-1)


def dynamic_graph_attention(knn_map, k, q, q_pos, v=None, v_pos=None, denoise_length=None):
    """DynamicGraphAttention.forward (models/Transformer_utils.py:796-860) restated with the reference formulation of
    knn_point; ``knn_map`` is the module's learned map."""
    if denoise_length is None:
        v = q if v is None else v
        v_pos = q_pos if v_pos is None else v_pos
        feature = dynamic_graph_feature(q, v, knn_point(k, v_pos, q_pos))
    else:
        r = q.size(1) - denoise_length
        f_r = dynamic_graph_feature(q[:, :r], q[:, :r], knn_point(k, q_pos[:, :r], q_pos[:, :r]))
        f_n = dynamic_graph_feature(q[:, r:], q, knn_point(k, q_pos, q_pos[:, r:]))
        feature = torch.cat([f_r, f_n], dim=1)
    return knn_map(feature).max(-2)[0]
