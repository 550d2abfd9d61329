"""Deterministic synthetic PCN-shaped inputs.  CPU torch only; callers move to the GPU.

  U: iid uniform in [-0.5, 0.5)^3
  S: "PCN-shaped": gt on the surface of a random union of 3 axis-aligned boxes + 1 ellipsoid scaled into
     [-0.5,0.5]^3; pred = gt[perm] + N(0, 0.01^2); partial = the points facing a random view direction, subsampled
     to 2048 or zero-padded like datasets/data_transforms.py:55-67
  G: U rounded to a 2^-bits grid (every distance formula is exact in FP32 -> real ties), optionally with the
     tail duplicated / zeroed (PCN padding)
"""
from __future__ import annotations

import torch


def _gen(seed):
    return torch.Generator().manual_seed(int(seed))


def uniform(B, N, seed):
    return torch.rand((B, N, 3), generator=_gen(seed), dtype=torch.float32) - 0.5


def grid(B, N, seed, bits=6, dup_frac=0.0, zero_frac=0.0):
    g = _gen(seed)
    scale = float(1 << bits)
    x = torch.floor((torch.rand((B, N, 3), generator=g) - 0.5) * scale) / scale
    x = x.to(torch.float32)
    n_dup = int(N * dup_frac)
    if n_dup:
        src = torch.randint(0, N - n_dup, (B, n_dup), generator=g)
        x[:, N - n_dup:] = torch.gather(x, 1, src.unsqueeze(-1).expand(-1, -1, 3))
    n_zero = int(N * zero_frac)
    if n_zero:
        x[:, N - n_zero:] = 0.0
    return x


def _surface_cloud(N, g):
    """N points on a union of 3 boxes + 1 ellipsoid, inside [-0.5, 0.5]^3."""
    parts = []
    per = [N // 4] * 3 + [N - 3 * (N // 4)]
    for i in range(3):
        n = per[i]
        half = 0.08 + 0.17 * torch.rand(3, generator=g)
        center = (torch.rand(3, generator=g) - 0.5) * 0.4
        p = (torch.rand((n, 3), generator=g) * 2 - 1) * half
        face = torch.randint(0, 3, (n,), generator=g)
        sign = torch.randint(0, 2, (n,), generator=g).float() * 2 - 1
        p[torch.arange(n), face] = sign * half[face]
        parts.append(p + center)
    n = per[3]
    v = torch.randn((n, 3), generator=g)
    v = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
    axes = 0.1 + 0.2 * torch.rand(3, generator=g)
    parts.append(v * axes + (torch.rand(3, generator=g) - 0.5) * 0.3)
    pts = torch.cat(parts, 0)
    pts = pts[torch.randperm(N, generator=g)]
    return pts.clamp_(-0.5, 0.5).to(torch.float32)


def pcn_gt(B, N, seed):
    g = _gen(seed)
    return torch.stack([_surface_cloud(N, g) for _ in range(B)])


def pcn_pair(B, N, seed, noise=0.01):
    """(pred, gt) [B,N,3]: pred = permuted gt + Gaussian noise."""
    g = _gen(seed)
    gt = torch.stack([_surface_cloud(N, g) for _ in range(B)])
    perm = torch.stack([torch.randperm(N, generator=g) for _ in range(B)])
    pred = torch.gather(gt, 1, perm.unsqueeze(-1).expand(-1, -1, 3)) + noise * torch.randn((B, N, 3), generator=g)
    return pred.to(torch.float32).contiguous(), gt.contiguous()


def pcn_partial(gt, n_partial, seed):
    """Visible half of each gt cloud w.r.t. a random direction, subsampled / zero-padded to n_partial."""
    g = _gen(seed)
    out = []
    for cloud in gt:
        n = torch.randn(3, generator=g)
        vis = cloud[(cloud @ n) > 0]
        keep = 0.3 + 0.7 * float(torch.rand((), generator=g))
        vis = vis[torch.randperm(vis.size(0), generator=g)[: max(1, int(vis.size(0) * keep))]]
        if vis.size(0) >= n_partial:
            vis = vis[torch.randperm(vis.size(0), generator=g)[:n_partial]]
        else:
            vis = torch.cat([vis, torch.zeros((n_partial - vis.size(0), 3))], 0)
        out.append(vis)
    return torch.stack(out).to(torch.float32).contiguous()
