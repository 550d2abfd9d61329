"""Drop-in for the GEOMETRY parts of the reference's ``models/base_blocks.py`` (the conv / attention layers around them
are out of scope and stay the reference's own):

* ``get_graph_feature`` - DGCNN_Grouper.get_graph_feature (base_blocks.py:565-588): kNN (no grad) + neighbour gather +
  ``cat(feature - x_q, x_q)`` -> ``[B, 2C, Nq, k]``, here one fused kernel after the fused kNN kernel (the reference
  makes five passes over that tensor), differentiable w.r.t. ``x_q`` and ``x_k`` with a deterministic backward.
* ``fps_downsample`` - DGCNN_Grouper.fps_downsample (base_blocks.py:547-563).

``dropin.patch_reference_modules()`` installs them on the reference's ``DGCNN_Grouper`` class.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..pointnet2_ops import pointnet2_utils
from .Transformer_utils import knn_point


class GraphFeature(torch.autograd.Function):
    """(x_q [B,C,Nq], x_k [B,C,Nk], idx [B,Nq,k] int64) -> [B,2C,Nq,k] = cat(x_k[..., idx] - x_q, x_q)."""

    @staticmethod
    def forward(ctx, x_q, x_k, idx):
        _lib.require_cuda_f32(x_q, "x_q", 3)
        _lib.require_cuda_f32(x_k, "x_k", 3)
        if idx.dtype != torch.int64 or idx.dim() != 3 or not idx.is_cuda:
            raise _lib.IaptError("idx must be a CUDA int64 [B, Nq, k] tensor (as knn_point returns)")
        x_q, x_k, idx = x_q.contiguous(), x_k.contiguous(), idx.contiguous()
        B, C, Nq = x_q.shape
        Nk = x_k.size(2)
        k = idx.size(2)
        if x_k.size(0) != B or x_k.size(1) != C or idx.size(0) != B or idx.size(1) != Nq:
            raise _lib.IaptError("graph_feature: shape mismatch")
        dev = x_q.device
        with _lib.on_device(dev):
            out = torch.empty((B, 2 * C, Nq, k), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_graph_feature(_lib.ptr(x_q), _lib.ptr(x_k), _lib.ptr(idx), B, C, Nq, Nk, k,
                                                     _lib.ptr(out), _lib.stream_ptr(dev)), "iapt_graph_feature")
        ctx.save_for_backward(idx)
        ctx.dims = (B, C, Nq, Nk, k)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, C, Nq, Nk, k = ctx.dims
        grad_out = grad_out.to(torch.float32).contiguous()
        dev = grad_out.device
        need_q, need_k = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        L = _lib.lib()
        with _lib.on_device(dev):
            gq = torch.empty((B, C, Nq), dtype=torch.float32, device=dev) if need_q else None
            gk = torch.empty((B, C, Nk), dtype=torch.float32, device=dev) if need_k else None
            ws_bytes = int(L.iapt_graph_feature_grad_workspace_bytes(B, Nq, Nk, k)) if need_k else 0
            ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
            _lib.check(L.iapt_graph_feature_grad(_lib.ptr(grad_out), _lib.ptr(idx), B, C, Nq, Nk, k, _lib.ptr(gq),
                                                 _lib.ptr(gk), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
                       "iapt_graph_feature_grad")
        return gq, gk, None


def graph_feature(x_q, x_k, idx):
    return GraphFeature.apply(x_q, x_k, idx)


def get_graph_feature(k, coor_q, x_q, coor_k, x_k):
    """coor: [B,3,n], x: [B,C,n] -> [B,2C,Nq,k]  (base_blocks.py:565-588; ``k`` is the grouper's ``self.k``)."""
    with torch.no_grad():
        idx = knn_point(k, coor_k.transpose(-1, -2).contiguous(), coor_q.transpose(-1, -2).contiguous())  # [B,Nq,k]
    return graph_feature(x_q, x_k, idx)


def fps_downsample(coor, x, num_group):
    """coor [B,3,N], x [B,C,N] -> (new_coor [B,3,m], new_x [B,C,m])  (base_blocks.py:547-563)."""
    xyz = coor.transpose(1, 2).contiguous()
    fps_idx = pointnet2_utils.furthest_point_sample(xyz, num_group)
    combined_x = torch.cat([coor, x], dim=1)
    new_combined_x = pointnet2_utils.gather_operation(combined_x, fps_idx)
    return new_combined_x[:, :3], new_combined_x[:, 3:]
