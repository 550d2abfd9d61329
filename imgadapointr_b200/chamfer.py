"""Drop-in for the reference's pybind module ``chamfer`` (extensions/chamfer_dist/chamfer_cuda.cpp:36-39).

``forward(xyz1, xyz2) -> [dist1, dist2, idx1, idx2]`` and
``backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2) -> [grad_xyz1, grad_xyz2]``
with the reference's shapes and dtypes (dist float32, idx int32), computed by libiapt_b200.so on the current
stream of the inputs' device.  Differences, all deliberate:
  * inputs are validated (CUDA, float32, [B,n,3]) and made contiguous instead of silently reading garbage;
  * CUDA errors raise instead of being printf()ed (chamfer.cu:166-169);
  * the current torch stream is used instead of the legacy default stream (chamfer.cu:159).
``step`` is the whole loss / metric step (search + reductions + gradient) as one kernel launch
(``iapt_chamfer_step``); the loss modules and ``utils.metrics`` are built on it.
"""
from __future__ import annotations

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


def _workspace(nbytes, dev):
    return torch.empty((max(int(nbytes), 8),), dtype=torch.uint8, device=dev)


def forward(xyz1: torch.Tensor, xyz2: torch.Tensor):
    xyz1, xyz2 = _prep(xyz1.detach(), xyz2.detach())
    ext = _lib.torch_ext()
    if ext is not None:
        return ext.forward(xyz1, xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    L = _lib.lib()
    with _lib.on_device(dev):
        dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
        dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
        idx1 = torch.empty((B, N), dtype=torch.int32, device=dev)
        idx2 = torch.empty((B, M), dtype=torch.int32, device=dev)
        ws_bytes = int(L.iapt_chamfer_forward_workspace_bytes(B, N, M))
        ws = _workspace(ws_bytes, dev)
        _lib.check(
            L.iapt_chamfer_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), B, N, M, _lib.ptr(dist1), _lib.ptr(dist2),
                                   _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
            "iapt_chamfer_forward")
    return [dist1, dist2, idx1, idx2]


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2, need1: bool = True, need2: bool = True):
    """[grad_xyz1, grad_xyz2]; ``need1`` / ``need2`` = False skips that gradient (returned as None)."""
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
    if not (need1 or need2):
        return [None, None]
    ext = _lib.torch_ext()
    if ext is not None:
        return ext.backward(xyz1, xyz2, idx1, idx2, g1, g2, bool(need1), bool(need2))
    L = _lib.lib()
    with _lib.on_device(dev):
        gx1 = torch.empty_like(xyz1) if need1 else None
        gx2 = torch.empty_like(xyz2) if need2 else None
        ws_bytes = int(L.iapt_chamfer_backward_workspace_bytes(B, N, M))
        ws = _workspace(ws_bytes, dev)
        _lib.check(
            L.iapt_chamfer_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(g1),
                                    _lib.ptr(g2), B, N, M, _lib.ptr(gx1), _lib.ptr(gx2), _lib.ptr(ws), ws_bytes,
                                    _lib.stream_ptr(dev)),
            "iapt_chamfer_backward")
    return [gx1, gx2]


LOSS_NONE, LOSS_L1, LOSS_L2, LOSS_L2_SPLIT, LOSS_L1_PM = 0, 1, 2, 3, 4
STEP_IGNORE_ZERO_SUM, STEP_STATS, STEP_GATE_ON_VALID = 1, 2, 4


class StepResult:
    """Outputs of one ``iapt_chamfer_step`` launch (attributes are None when not requested)."""

    __slots__ = ("dist1", "dist2", "idx1", "idx2", "sums", "counts", "valid", "loss", "grad1", "grad2", "workspace")


def step(xyz1, xyz2, kind: int = LOSS_NONE, *, sums: bool = False, counts: bool = False, valid: bool = False,
         th: float = 0.0, grad1: bool = False, grad2: bool = False, grad_loss=None, ignore_zero_sum: bool = False,
         gate_valid=None, stats: bool = False) -> StepResult:
    """The whole Chamfer step in ONE kernel launch: nearest-neighbour search, per-cloud sums / F-Score counts, the loss
    scalar(s) of ``kind`` and (``grad1`` / ``grad2``) its gradient w.r.t. the clouds for upstream ``grad_loss`` (a CUDA
    tensor; None = 1, i.e. the unit gradient to be scaled later with ``scale``).  ``ignore_zero_sum`` applies the
    reference's ``ignore_zeros`` filter inside the kernel; ``gate_valid`` (the ``valid`` tensor of an earlier step on the
    same clouds) turns the launch into a no-op when there is nothing to ignore - decided on the device, no read-back."""
    xyz1, xyz2 = _prep(xyz1.detach(), xyz2.detach())
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    L = _lib.lib()
    r = StepResult()
    flags = (STEP_IGNORE_ZERO_SUM if ignore_zero_sum else 0) | (STEP_STATS if stats else 0)
    if gate_valid is not None:
        if gate_valid.dtype != torch.int32 or tuple(gate_valid.shape) != (B, 2) or not gate_valid.is_contiguous():
            raise _lib.IaptError("step: gate_valid must be the contiguous int32 [B,2] `valid` of an earlier step")
        flags |= STEP_GATE_ON_VALID
    if grad_loss is not None:
        grad_loss = grad_loss.to(torch.float32).contiguous()
        if grad_loss.numel() != (2 if kind == LOSS_L2_SPLIT else 1):
            raise _lib.IaptError("step: grad_loss has the wrong number of elements")
    ext = _lib.torch_ext()
    if ext is not None:
        (r.dist1, r.dist2, r.idx1, r.idx2, r.sums, r.counts, r.valid, r.loss, r.grad1, r.grad2,
         r.workspace) = ext.step(xyz1, xyz2, int(kind), flags & ~STEP_GATE_ON_VALID, float(th), bool(sums), bool(counts),
                                 bool(valid), gate_valid, bool(grad1), bool(grad2), grad_loss)
        return r
    with _lib.on_device(dev):
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        r.dist1, r.dist2 = torch.empty((B, N), **f32), torch.empty((B, M), **f32)
        r.idx1, r.idx2 = torch.empty((B, N), **i32), torch.empty((B, M), **i32)
        r.sums = torch.empty((B, 4), **f32) if sums else None
        r.counts = torch.empty((B, 2), **i32) if counts else None
        r.valid = gate_valid if gate_valid is not None else (torch.empty((B, 2), **i32) if valid else None)
        r.loss = torch.empty((2,), **f32) if kind != LOSS_NONE else None
        r.grad1 = torch.empty_like(xyz1) if grad1 else None
        r.grad2 = torch.empty_like(xyz2) if grad2 else None
        ws_bytes = int(L.iapt_chamfer_step_workspace_bytes(B, N, M))
        r.workspace = _workspace(ws_bytes, dev)
        _lib.check(
            L.iapt_chamfer_step(_lib.ptr(xyz1), _lib.ptr(xyz2), B, N, M, int(kind), flags, float(th), _lib.ptr(r.dist1),
                                _lib.ptr(r.dist2), _lib.ptr(r.idx1), _lib.ptr(r.idx2), _lib.ptr(r.sums),
                                _lib.ptr(r.counts), _lib.ptr(r.valid), _lib.ptr(r.loss), _lib.ptr(grad_loss),
                                _lib.ptr(r.grad1), _lib.ptr(r.grad2), _lib.ptr(r.workspace), ws_bytes,
                                _lib.stream_ptr(dev)),
            "iapt_chamfer_step")
    return r


def step_stats(workspace: torch.Tensor):
    """(search_ms, kernel_ms) of the last ``stats=True`` step on this workspace, from device timestamps (synchronises)."""
    import ctypes

    out = (ctypes.c_ulonglong * 9)()
    _lib.check(_lib.lib().iapt_chamfer_step_stats(_lib.ptr(workspace), out), "iapt_chamfer_step_stats")
    return (out[2] - out[0]) * 1e-6, (out[3] - out[1]) * 1e-6


def step_phase_ends(workspace: torch.Tensor):
    """ms after kernel start at which the last item of each phase (search, resolve, prefix, fill, gather) ended; None
    for a phase that did not run."""
    import ctypes

    out = (ctypes.c_ulonglong * 9)()
    _lib.check(_lib.lib().iapt_chamfer_step_stats(_lib.ptr(workspace), out), "iapt_chamfer_step_stats")
    return [((out[4 + h] - out[1]) * 1e-6 if out[4 + h] else None) for h in range(5)]


def scale(src: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    """src * factor[0] with the factor read on the device (no host sync): applies an upstream gradient (iapt_scale)."""
    src = src.contiguous()
    factor = factor.to(torch.float32).reshape(-1).contiguous()
    ext = _lib.torch_ext()
    if ext is not None:
        return ext.scale(src, factor)
    dev = src.device
    with _lib.on_device(dev):
        dst = torch.empty_like(src)
        _lib.check(_lib.lib().iapt_scale(_lib.ptr(src), _lib.ptr(factor), src.numel(), _lib.ptr(dst),
                                         _lib.stream_ptr(dev)), "iapt_scale")
    return dst


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


def loss_forward(dist1: torch.Tensor, dist2: torch.Tensor, kind: int):
    """(loss[2], sums[B,4]) from given distances - the reference modules' reductions in two small launches
    (iapt_chamfer_loss).  The loss modules use ``step`` instead; this serves callers that already hold distances."""
    B, N = dist1.shape
    M = dist2.size(1)
    dev = dist1.device
    with _lib.on_device(dev):
        sums = torch.empty((B, 4), dtype=torch.float32, device=dev)
        loss = torch.empty((2,), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().iapt_chamfer_loss(_lib.ptr(dist1), _lib.ptr(dist2), B, N, M, int(kind), _lib.ptr(sums),
                                                _lib.ptr(loss), _lib.stream_ptr(dev)), "iapt_chamfer_loss")
    return loss, sums


def loss_backward(xyz1, xyz2, idx1, idx2, dist1, dist2, kind: int, grad_loss, need1: bool = True, need2: bool = True):
    """Gradient of the fused loss w.r.t. (xyz1, xyz2) from saved indices and distances in one launch
    (iapt_chamfer_backward_fused).  grad_loss: CUDA float32 tensor with 1 element (2 for LOSS_L2_SPLIT) or None (= 1)."""
    B, N, _ = xyz1.shape
    M = xyz2.size(1)
    dev = xyz1.device
    if grad_loss is not None:
        grad_loss = grad_loss.to(torch.float32).contiguous()
        if grad_loss.numel() != (2 if kind == LOSS_L2_SPLIT else 1):
            raise _lib.IaptError("loss_backward: grad_loss has the wrong number of elements")
    if not (need1 or need2):
        return None, None
    ext = _lib.torch_ext()
    if ext is not None:
        gx1, gx2 = ext.loss_backward(xyz1, xyz2, idx1, idx2, dist1, dist2, int(kind), grad_loss, bool(need1), bool(need2))
        return gx1, gx2
    L = _lib.lib()
    with _lib.on_device(dev):
        gx1 = torch.empty_like(xyz1) if need1 else None
        gx2 = torch.empty_like(xyz2) if need2 else None
        ws_bytes = int(L.iapt_chamfer_backward_workspace_bytes(B, N, M))
        ws = _workspace(ws_bytes, dev)
        _lib.check(
            L.iapt_chamfer_backward_fused(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx1), _lib.ptr(idx2),
                                          _lib.ptr(dist1), _lib.ptr(dist2), int(kind), _lib.ptr(grad_loss), B, N, M,
                                          _lib.ptr(gx1), _lib.ptr(gx2), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
            "iapt_chamfer_backward_fused")
    return gx1, gx2
