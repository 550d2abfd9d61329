"""Drop-in for the reference's pybind module ``chamfer`` (extensions/chamfer_dist/chamfer_cuda.cpp:36-39).

``forward(xyz1, xyz2) -> [dist1, dist2, idx1, idx2]`` and
``backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2) -> [grad_xyz1, grad_xyz2]``
with the reference's shapes and dtypes (dist float32, idx int32), computed by libiapt_b200.so on the current
stream of the inputs' device.  Differences, all deliberate:
  * inputs are validated (CUDA, float32, [B,n,3]) and made contiguous instead of silently reading garbage;
  * CUDA errors raise instead of being printf()ed (chamfer.cu:166-169);
  * the current torch stream is used instead of the legacy default stream (chamfer.cu:159).
"""
from __future__ import annotations

import os

import torch

from . import _lib


def _prep(xyz1, xyz2):
    _lib.require_cuda_f32(xyz1, "xyz1", 3)
    _lib.require_cuda_f32(xyz2, "xyz2", 3)
    if xyz1.size(2) != 3 or xyz2.size(2) != 3:
        raise _lib.IaptError("xyz1 / xyz2 must be [B, n, 3]")
    if xyz1.size(0) != xyz2.size(0):
        raise _lib.IaptError("xyz1 and xyz2 must have the same batch size")
    if xyz1.device != xyz2.device:
        raise _lib.IaptError("xyz1 and xyz2 must be on the same device")
    return xyz1.contiguous(), xyz2.contiguous()


def forward(xyz1: torch.Tensor, xyz2: torch.Tensor):
    xyz1, xyz2 = _prep(xyz1.detach(), xyz2.detach())
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    L = _lib.lib()
    with _lib.on_device(dev):
        dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx1 = torch.empty((B, N), dtype=torch.int32, device=dev)
        idx2 = torch.empty((B, M), dtype=torch.int32, device=dev)
        ws_bytes = 8 * B * (N + M) + 256 if (B > 0 and N > 0 and M > 0) else 0  # == iapt_chamfer_forward_workspace_bytes
        ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
        _lib.check(
            L.iapt_chamfer_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), B, N, M, _lib.ptr(dist1), _lib.ptr(dist2),
                                   _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
            "iapt_chamfer_forward")
    return [dist1, dist2, idx1, idx2]


_USE_BWD_LISTS = os.environ.get("IAPT_BWD_LISTS", "1") != "0"


def _bwd_workspace(L, B, N, M, dev):
    """Workspace of the fan-in-list backward path (include/iapt_b200.h); none when IAPT_BWD_LISTS=0 (scan kernel only)."""
    if not _USE_BWD_LISTS or B <= 0 or N <= 0 or M <= 0:
        return None, 0
    n = int(L.iapt_chamfer_backward_workspace_bytes(B, N, M))
    return (torch.empty((n,), dtype=torch.uint8, device=dev) if n else None), n


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    xyz1, xyz2 = _prep(xyz1.detach(), xyz2.detach())
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    if idx1.dtype != torch.int32 or idx2.dtype != torch.int32:
        raise _lib.IaptError("idx1 / idx2 must be int32 (as returned by forward)")
    idx1, idx2 = idx1.contiguous(), idx2.contiguous()
    g1 = grad_dist1.to(torch.float32).contiguous()  # autograd may hand over expanded / strided grads
    g2 = grad_dist2.to(torch.float32).contiguous()
    if tuple(g1.shape) != (B, N) or tuple(g2.shape) != (B, M) or tuple(idx1.shape) != (B, N) or tuple(idx2.shape) != (B, M):
        raise _lib.IaptError("backward: shape mismatch between xyz, idx and grad_dist")
    L = _lib.lib()
    with _lib.on_device(dev):
        gx1 = torch.empty_like(xyz1)
        gx2 = torch.empty_like(xyz2)
        ws, ws_bytes = _bwd_workspace(L, B, N, M, dev)
        _lib.check(
            L.iapt_chamfer_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(g1),
                                    _lib.ptr(g2), B, N, M, _lib.ptr(gx1), _lib.ptr(gx2), _lib.ptr(ws), ws_bytes,
                                    _lib.stream_ptr(dev)),
            "iapt_chamfer_backward")
    return [gx1, gx2]


def reduce_sums(dist1: torch.Tensor, dist2: torch.Tensor) -> torch.Tensor:
    """[B,4] per-cloud {sum d1, sum d2, sum sqrt d1, sum sqrt d2} in one launch (iapt_chamfer_reduce)."""
    B, N = dist1.shape
    M = dist2.size(1)
    dev = dist1.device
    with _lib.on_device(dev):
        sums = torch.empty((B, 4), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().iapt_chamfer_reduce(_lib.ptr(dist1), _lib.ptr(dist2), B, N, M, _lib.ptr(sums),
                                                  _lib.stream_ptr(dev)), "iapt_chamfer_reduce")
    return sums


def fscore_counts(dist1: torch.Tensor, dist2: torch.Tensor, th: float) -> torch.Tensor:
    """[B,2] int32 {#sqrt(dist1)<th, #sqrt(dist2)<th} (iapt_fscore_counts)."""
    B, N = dist1.shape
    M = dist2.size(1)
    dev = dist1.device
    with _lib.on_device(dev):
        counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().iapt_fscore_counts(_lib.ptr(dist1), _lib.ptr(dist2), B, N, M, float(th),
                                                 _lib.ptr(counts), _lib.stream_ptr(dev)), "iapt_fscore_counts")
    return counts


LOSS_L1, LOSS_L2, LOSS_L2_SPLIT, LOSS_L1_PM = 1, 2, 3, 4


def loss_forward(dist1: torch.Tensor, dist2: torch.Tensor, kind: int):
    """(loss[2], sums[B,4]) - the reference modules' reductions in two small launches (iapt_chamfer_loss)."""
    B, N = dist1.shape
    M = dist2.size(1)
    dev = dist1.device
    with _lib.on_device(dev):
        sums = torch.empty((B, 4), dtype=torch.float32, device=dev)
        loss = torch.empty((2,), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().iapt_chamfer_loss(_lib.ptr(dist1), _lib.ptr(dist2), B, N, M, int(kind), _lib.ptr(sums),
                                                _lib.ptr(loss), _lib.stream_ptr(dev)), "iapt_chamfer_loss")
    return loss, sums


def loss_backward(xyz1, xyz2, idx1, idx2, dist1, dist2, kind: int, grad_loss):
    """Gradient of the fused loss w.r.t. (xyz1, xyz2) in one kernel (iapt_chamfer_backward_fused).
    grad_loss: CUDA float32 tensor with 1 element (2 for LOSS_L2_SPLIT) or None (= 1)."""
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    if grad_loss is not None:
        grad_loss = grad_loss.to(torch.float32).contiguous()
        if grad_loss.numel() != (2 if kind == LOSS_L2_SPLIT else 1):
            raise _lib.IaptError("loss_backward: grad_loss has the wrong number of elements")
    with _lib.on_device(dev):
        gx1 = torch.empty_like(xyz1)
        gx2 = torch.empty_like(xyz2)
        ws, ws_bytes = _bwd_workspace(_lib.lib(), B, N, M, dev)
        _lib.check(
            _lib.lib().iapt_chamfer_backward_fused(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx1), _lib.ptr(idx2),
                                                   _lib.ptr(dist1), _lib.ptr(dist2), int(kind), _lib.ptr(grad_loss), B, N,
                                                   M, _lib.ptr(gx1), _lib.ptr(gx2), _lib.ptr(ws), ws_bytes,
                                                   _lib.stream_ptr(dev)),
            "iapt_chamfer_backward_fused")
    return gx1, gx2
