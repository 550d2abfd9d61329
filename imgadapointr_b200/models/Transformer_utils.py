"""Drop-in for the geometry helpers of the reference's ``models/Transformer_utils.py``:
``knn_point`` (:19-30), ``square_distance`` (:32-51), ``index_points`` (:53-69)."""
from __future__ import annotations

import torch

from .. import _lib


def knn_point(nsample, xyz, new_xyz, return_dist=False):
    """
    Input:
        nsample: max sample number in local region
        xyz: all points, [B, N, 3]
        new_xyz: query points, [B, S, 3]
    Return:
        group_idx: grouped points index, [B, S, nsample] int64

    The reference takes topk(sorted=False) of a materialised [B,S,N] matrix; this returns the same neighbour SET
    from one fused kernel, ordered ascending by (distance, index) (lowest index on exact ties).  Inputs may be
    non-contiguous slices (Transformer_utils.py:839); no autograd flows through the indices.
    """
    xyz = xyz.detach()
    new_xyz = new_xyz.detach()
    _lib.require_cuda_f32(xyz, "xyz", 3)
    _lib.require_cuda_f32(new_xyz, "new_xyz", 3)
    if xyz.size(2) != 3 or new_xyz.size(2) != 3:
        raise _lib.IaptError("knn_point: the B200 kernel handles 3-D coordinates ([B,N,3])")
    if xyz.size(0) != new_xyz.size(0):
        raise _lib.IaptError("knn_point: batch sizes differ")
    xyz, new_xyz = xyz.contiguous(), new_xyz.contiguous()
    B, R, _ = xyz.shape
    Q = new_xyz.size(1)
    k = int(nsample)
    if k > R:
        raise RuntimeError(f"knn_point: nsample={k} exceeds the number of points {R} (torch.topk raises here too)")
    dev = xyz.device
    with _lib.on_device(dev):
        idx = torch.empty((B, Q, k), dtype=torch.int64, device=dev)
        dist = torch.empty((B, Q, k), dtype=torch.float32, device=dev) if return_dist else None
        ws_bytes = int(_lib.lib().iapt_knn_workspace_bytes(B, R, Q, k))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        _lib.check(_lib.lib().iapt_knn(_lib.ptr(xyz), _lib.ptr(new_xyz), B, R, Q, k, _lib.ptr(idx), _lib.ptr(dist),
                                       _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)), "iapt_knn")
    return (idx, dist) if return_dist else idx


def square_distance(src, dst):
    """[B,N,C],[B,M,C] -> [B,N,M] squared distances (Transformer_utils.py:32-51).  Kept for API parity: nothing on the
    accelerated path materialises this matrix any more (knn_point is fused); plain torch ops, same op order."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def index_points(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,C] / [B,S,K,C] (Transformer_utils.py:53-69).

    One gather kernel (iapt_index_points) instead of the reference's arange/repeat + advanced indexing; when
    gradients w.r.t. ``points`` are needed the differentiable torch.gather formulation is used."""
    B = points.shape[0]
    C = points.shape[-1]
    flat = idx.reshape(B, -1).to(torch.int64)
    fast = (points.is_cuda and points.dtype == torch.float32 and points.dim() == 3 and flat.is_cuda
            and not (torch.is_grad_enabled() and points.requires_grad))
    if fast:
        pts = points.contiguous()
        flat = flat.contiguous()
        S = flat.size(1)
        dev = pts.device
        with _lib.on_device(dev):
            out = torch.empty((B, S, C), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_index_points(_lib.ptr(pts), _lib.ptr(flat), B, pts.size(1), C, S, _lib.ptr(out),
                                                    _lib.stream_ptr(dev)), "iapt_index_points")
        return out.reshape(*idx.shape, C)
    out = torch.gather(points, 1, flat.unsqueeze(-1).expand(-1, -1, C))
    return out.reshape(*idx.shape, C)
