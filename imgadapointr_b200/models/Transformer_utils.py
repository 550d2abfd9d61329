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
    ext = _lib.torch_ext()
    if ext is not None:
        idx, dist = ext.knn(xyz, new_xyz, k, bool(return_dist))
        return (idx, dist) if return_dist else idx
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
    """[B,N,C],[B,M,C] -> [B,N,M] squared distances in the reference's expansion form and op order
    (Transformer_utils.py:32-51: ``-2 src.dst^T``, ``+= sum(src^2)``, ``+= sum(dst^2)``), one kernel writing the matrix
    once (iapt_square_distance).  Kept for API parity: ``knn_point`` is fused and never materialises it.  CUDA float32
    only, like every entry point here; not differentiable (the reference only uses it under ``knn_point``)."""
    src, dst = src.detach(), dst.detach()
    _lib.require_cuda_f32(src, "src", 3)
    _lib.require_cuda_f32(dst, "dst", 3)
    if src.size(0) != dst.size(0) or src.size(2) != dst.size(2):
        raise _lib.IaptError("square_distance: src [B,N,C] and dst [B,M,C] must share B and C")
    src, dst = src.contiguous(), dst.contiguous()
    ext = _lib.torch_ext()
    if ext is not None:
        return ext.square_distance(src, dst)
    B, N, C = src.shape
    M = dst.size(1)
    dev = src.device
    with _lib.on_device(dev):
        out = torch.empty((B, N, M), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().iapt_square_distance(_lib.ptr(src), _lib.ptr(dst), B, N, M, C, _lib.ptr(out),
                                                   _lib.stream_ptr(dev)), "iapt_square_distance")
    return out


class _IndexPoints(torch.autograd.Function):
    """points [B,N,C] float32 CUDA, flat idx [B,S] int64 -> [B,S,C]; deterministic row-sum backward."""

    @staticmethod
    def forward(ctx, pts, flat):
        B, N, C = pts.shape
        S = flat.size(1)
        dev = pts.device
        with _lib.on_device(dev):
            out = torch.empty((B, S, C), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_index_points(_lib.ptr(pts), _lib.ptr(flat), B, N, C, S, _lib.ptr(out),
                                                    _lib.stream_ptr(dev)), "iapt_index_points")
        ctx.save_for_backward(flat)
        ctx.dims = (B, N, C, S)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (flat,) = ctx.saved_tensors
        B, N, C, S = ctx.dims
        grad_out = grad_out.to(torch.float32).contiguous()
        dev = grad_out.device
        L = _lib.lib()
        with _lib.on_device(dev):
            gp = torch.empty((B, N, C), dtype=torch.float32, device=dev)
            ws_bytes = int(L.iapt_graph_feature_grad_workspace_bytes(B, S, N, 1))
            ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
            _lib.check(L.iapt_index_points_grad(_lib.ptr(grad_out), _lib.ptr(flat), B, N, C, S, _lib.ptr(gp), _lib.ptr(ws),
                                                ws_bytes, _lib.stream_ptr(dev)), "iapt_index_points_grad")
        return gp, None


def index_points(points, idx):
    """points [B,N,C], idx [B,S] or [B,S,K] -> [B,S,C] / [B,S,K,C] (Transformer_utils.py:53-69).

    One gather kernel (iapt_index_points) instead of the reference's arange/repeat + advanced indexing, and when
    ``points`` needs a gradient a deterministic backward (iapt_index_points_grad: every source row is the sum, in
    ascending position, of the gradient rows that gathered it) instead of torch's atomicAdd scatter.  CUDA float32
    ``points`` and CUDA integer ``idx`` only - like every other entry point there is no eager / CPU path."""
    _lib.require_cuda_f32(points, "points", 3)
    if not (isinstance(idx, torch.Tensor) and idx.is_cuda and not idx.is_floating_point() and idx.dim() >= 2):
        raise _lib.IaptError("index_points: idx must be a CUDA integer tensor [B, S] or [B, S, K]")
    B = points.shape[0]
    C = points.shape[-1]
    if idx.size(0) != B:
        raise _lib.IaptError("index_points: points and idx must share the batch size")
    flat = idx.reshape(B, -1).to(torch.int64)
    if C == 0 or flat.size(1) == 0 or B == 0:  # nothing to gather
        return points.new_empty((*idx.shape, C))
    return _IndexPoints.apply(points.contiguous(), flat.contiguous()).reshape(*idx.shape, C)


class _GraphFeatureCL(torch.autograd.Function):
    """(q [B,N,C], v [B,M,C], idx [B,N,k] int64) -> [B,N,k,2C] = cat(v[idx] - q, q), channels last."""

    @staticmethod
    def forward(ctx, q, v, idx):
        B, N, C = q.shape
        M = v.size(1)
        k = idx.size(2)
        dev = q.device
        with _lib.on_device(dev):
            out = torch.empty((B, N, k, 2 * C), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_graph_feature_cl(_lib.ptr(q), _lib.ptr(v), _lib.ptr(idx), B, N, M, C, k,
                                                        _lib.ptr(out), _lib.stream_ptr(dev)), "iapt_graph_feature_cl")
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, M, C, k)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        B, N, M, C, k = ctx.dims
        grad_out = grad_out.to(torch.float32).contiguous()
        dev = grad_out.device
        need_q, need_v = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        L = _lib.lib()
        with _lib.on_device(dev):
            gq = torch.empty((B, N, C), dtype=torch.float32, device=dev) if need_q else None
            gv = torch.empty((B, M, C), dtype=torch.float32, device=dev) if need_v else None
            ws_bytes = int(L.iapt_graph_feature_grad_workspace_bytes(B, N, M, k)) if need_v else 0
            ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
            _lib.check(L.iapt_graph_feature_cl_grad(_lib.ptr(grad_out), _lib.ptr(idx), B, N, M, C, k, _lib.ptr(gq),
                                                    _lib.ptr(gv), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
                       "iapt_graph_feature_cl_grad")
        return gq, gv, None


def graph_feature_cl(q, v, idx):
    """``cat((index_points(v, idx) - q.unsqueeze(-2), q.unsqueeze(-2).expand(..)), -1)`` of DynamicGraphAttention
    (Transformer_utils.py:814-816) in one kernel: q [B,N,C], v [B,M,C], idx [B,N,k] -> [B,N,k,2C]; bit-identical
    values, deterministic gradients for q and v."""
    _lib.require_cuda_f32(q, "q", 3)
    _lib.require_cuda_f32(v, "v", 3)
    if idx.dtype != torch.int64 or idx.dim() != 3 or not idx.is_cuda:
        raise _lib.IaptError("idx must be a CUDA int64 [B, N, k] tensor (as knn_point returns)")
    if v.size(0) != q.size(0) or v.size(2) != q.size(2) or idx.size(0) != q.size(0) or idx.size(1) != q.size(1):
        raise _lib.IaptError("graph_feature_cl: shape mismatch")
    return _GraphFeatureCL.apply(q.contiguous(), v.contiguous(), idx.contiguous())


def _neighbour_indices(k, q_pos, v_pos, idx, denoise_length):
    """[B, N, k] indices into v for every query token.  Plain mode: the given ``idx`` or one fused kNN.  Denoise mode
    (self-attention only, Transformer_utils.py:826-849): the last ``denoise_length`` tokens look among ALL tokens, the
    others only among themselves - both index v directly, so the two kNN results are simply stacked."""
    if denoise_length is None:
        return knn_point(k, v_pos, q_pos) if idx is None else idx
    head = knn_point(k, v_pos[:, :-denoise_length], q_pos[:, :-denoise_length])
    tail = knn_point(k, v_pos, q_pos[:, -denoise_length:])
    return torch.cat([head, tail], dim=1)


def dynamic_graph_attention_forward(self, q, q_pos, v=None, v_pos=None, idx=None, denoise_length=None):
    """Replacement for DynamicGraphAttention.forward (Transformer_utils.py:796-860), same signature and result: the
    neighbourhood query runs on the fused kNN kernel and the gather / subtract / concat on one fused kernel
    (``graph_feature_cl``: [B,N,k,2C] = cat(v[idx] - q, q)); ``self.knn_map`` (the learned Linear + LeakyReLU) and the
    max over k are the module's own.  ``dropin.patch_reference_modules()`` installs it on the reference class."""
    if denoise_length is not None and not (idx is None and v is None and v_pos is None):
        raise ValueError("denoise_length applies to self-attention with on-line neighbour search: "
                         "v, v_pos and idx must not be given")
    v = q if v is None else v
    v_pos = q_pos if v_pos is None else v_pos
    for name, pos in (("q_pos", q_pos), ("v_pos", v_pos)):
        if pos.dim() != 3 or pos.size(-1) != 3:
            raise ValueError(f"{name} must be [B, N, 3], got {tuple(pos.shape)}")
    if q.size(-1) != self.dim or v.size(-1) != self.dim:
        raise ValueError(f"token width {q.size(-1)} / {v.size(-1)} does not match the module's dim {self.dim}")
    idx = _neighbour_indices(self.k, q_pos, v_pos, idx, denoise_length)
    if idx.size(-1) != self.k or idx.size(1) != q.size(1):
        raise ValueError(f"neighbour indices must be [B, {q.size(1)}, {self.k}], got {tuple(idx.shape)}")
    return self.knn_map(graph_feature_cl(q, v, idx)).max(-2)[0]  # [B, N, C]
