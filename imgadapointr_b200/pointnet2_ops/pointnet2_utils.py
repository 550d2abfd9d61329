"""Drop-in for ``pointnet2_ops.pointnet2_utils`` - the two ops on the reference's hot path.

``furthest_point_sample(xyz, npoint) -> int32 [B, npoint]``  (call sites utils/misc.py:37, models/base_blocks.py:550)
``gather_operation(features, idx) -> [B, C, npoint]``        (call sites utils/misc.py:38, models/base_blocks.py:555-558)

Like upstream they are ``autograd.Function``s (``.apply`` aliases), require CUDA float32 contiguous inputs and have
no CPU path.  Semantics of FPS: start at index 0, running minimum initialised to 1e10, upstream's near-origin skip
(``|p|^2 <= 1e-3`` points are never selected; PCN partial clouds are zero padded) ON by default, and the LOWEST
index on exact ties of the running minimum (upstream's tie order follows its thread layout; the selected
coordinates are identical either way when the tie is between duplicated points).  Set
``furthest_point_sample.skip_near_origin = False`` for plain FPS.

The remaining upstream names (three_nn, three_interpolate, grouping_operation, ball_query, QueryAndGroup,
GroupAll) are not reached by any shipped config (SURVEY.md note A) and raise NotImplementedError.
"""
from __future__ import annotations

import torch

from .. import _lib

SKIP_NEAR_ORIGIN_DEFAULT = True


def _fps(xyz: torch.Tensor, npoint: int, skip_near_origin: bool) -> torch.Tensor:
    _lib.require_cuda_f32(xyz, "xyz", 3)
    if xyz.size(2) != 3:
        raise _lib.IaptError("xyz must be [B, N, 3]")
    if not xyz.is_contiguous():
        raise _lib.IaptError("xyz must be contiguous (upstream CHECK_CONTIGUOUS)")
    B, N, _ = xyz.shape
    npoint = int(npoint)
    dev = xyz.device
    L = _lib.lib()
    with _lib.on_device(dev):
        idx = torch.empty((B, npoint), dtype=torch.int32, device=dev)
        ws_bytes = int(L.iapt_fps_workspace_bytes(B, N, npoint))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
        _lib.check(
            L.iapt_fps(_lib.ptr(xyz), B, N, npoint, 1 if skip_near_origin else 0, _lib.ptr(idx), _lib.ptr(ws), ws_bytes,
                       _lib.stream_ptr(dev)), "iapt_fps")
    return idx


class FurthestPointSampling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = _fps(xyz.detach(), npoint, SKIP_NEAR_ORIGIN_DEFAULT if furthest_point_sample.skip_near_origin is None
                   else furthest_point_sample.skip_near_origin)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


# ---- prefix reuse --------------------------------------------------------------------
# Greedy FPS always starts at index 0, so the m-point sample of a cloud is a PREFIX of any longer sample of the same
# cloud.  The reference samples the same tensor several times per forward (AdaPoinTr.py:113 -> 256 points, :126 -> 64
# points of the same `xyz`); the second call is served from the first one's result.  A cached result is reused only
# for the very same live tensor object (weak reference) at the same autograd version, i.e. provably unmodified.
_prefix_cache = []  # [(weakref(xyz), version, skip, idx)], most recent first
_PREFIX_CACHE_SIZE = 4


def _prefix_lookup(xyz, npoint, skip):
    for ref, version, sk, idx in _prefix_cache:
        if ref() is xyz and version == xyz._version and sk == skip and idx.size(1) >= npoint and idx.device == xyz.device:
            return idx
    return None


def furthest_point_sample(xyz, npoint):
    skip = SKIP_NEAR_ORIGIN_DEFAULT if furthest_point_sample.skip_near_origin is None \
        else furthest_point_sample.skip_near_origin
    if furthest_point_sample.reuse_prefix and isinstance(xyz, torch.Tensor):
        hit = _prefix_lookup(xyz, int(npoint), skip)
        if hit is not None:
            furthest_point_sample.prefix_hits += 1
            return hit[:, :int(npoint)].contiguous()
    out = FurthestPointSampling.apply(xyz, npoint)
    if furthest_point_sample.reuse_prefix:
        import weakref

        try:
            _prefix_cache.insert(0, (weakref.ref(xyz), xyz._version, skip, out))
            del _prefix_cache[_PREFIX_CACHE_SIZE:]
        except TypeError:
            pass
    return out


furthest_point_sample.skip_near_origin = None  # None -> SKIP_NEAR_ORIGIN_DEFAULT
furthest_point_sample.reuse_prefix = True
furthest_point_sample.prefix_hits = 0


def _check_idx(idx, B):
    if not idx.is_cuda or idx.dtype != torch.int32 or idx.dim() != 2 or idx.size(0) != B:
        raise _lib.IaptError("idx must be a CUDA int32 [B, npoint] tensor")
    return idx.contiguous()


class GatherOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        _lib.require_cuda_f32(features, "features", 3)
        if not features.is_contiguous():
            raise _lib.IaptError("features must be contiguous (upstream CHECK_CONTIGUOUS)")
        B, C, N = features.shape
        idx = _check_idx(idx, B)
        m = idx.size(1)
        dev = features.device
        with _lib.on_device(dev):
            out = torch.empty((B, C, m), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_gather(_lib.ptr(features), _lib.ptr(idx), B, C, N, m, _lib.ptr(out),
                                              _lib.stream_ptr(dev)), "iapt_gather")
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.to(torch.float32).contiguous()
        B, C, m = grad_out.shape
        dev = grad_out.device
        with _lib.on_device(dev):
            gfeat = torch.empty((B, C, ctx.N), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().iapt_gather_grad(_lib.ptr(grad_out), _lib.ptr(idx), B, C, ctx.N, m, _lib.ptr(gfeat),
                                                   _lib.stream_ptr(dev)), "iapt_gather_grad")
        return gfeat, None


gather_operation = GatherOperation.apply


def gather_xyz(xyz: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[B,N,3], int32 [B,m] -> [B,m,3] in one launch (the fused core of utils.misc.fps); not differentiable."""
    _lib.require_cuda_f32(xyz, "xyz", 3)
    xyz = xyz.contiguous()
    B, N, _ = xyz.shape
    idx = _check_idx(idx, B)
    m = idx.size(1)
    dev = xyz.device
    with _lib.on_device(dev):
        out = torch.empty((B, m, 3), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().iapt_gather_xyz(_lib.ptr(xyz), _lib.ptr(idx), B, N, m, _lib.ptr(out),
                                              _lib.stream_ptr(dev)), "iapt_gather_xyz")
    return out


def _not_on_hot_path(name):
    def f(*a, **k):
        raise NotImplementedError(
            f"pointnet2_utils.{name} is outside the accelerated hot path (not reached by any shipped config; "
            "SURVEY.md section 2 note A)")
    f.__name__ = name
    return f


three_nn = _not_on_hot_path("three_nn")
three_interpolate = _not_on_hot_path("three_interpolate")
grouping_operation = _not_on_hot_path("grouping_operation")
ball_query = _not_on_hot_path("ball_query")
