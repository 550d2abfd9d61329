"""EMD approximation module (auction algorithm) - same interface as the reference's extensions/emd/emd_module.py:31-79.

    dist, assignment = emdModule()(xyz1, xyz2, eps, iters)
        xyz1 [B,n,3] predicted cloud, xyz2 [B,n,3] ground truth, same size, coordinates normalised to [0, 1]
        dist [B,n]  squared distance of every predicted point to its assigned target (sqrt -> L2 distance)
        assignment [B,n] int32 index of the matched ground-truth point (not guaranteed to be a bijection)

The gradient flows to xyz1 only (as upstream).  Differences, all deliberate: inputs must already be CUDA float32
(upstream silently moves and casts them), n need not be a multiple of 1024, B is not limited to 512, the work buffers are
one workspace allocation, and the result is deterministic (see include/iapt_b200.h for the two tie rules).
"""
import torch
from torch import nn
from torch.autograd import Function

from ... import _lib


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        _lib.require_cuda_f32(xyz1, "xyz1", 3)
        _lib.require_cuda_f32(xyz2, "xyz2", 3)
        if xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
            raise _lib.IaptError("emd: expected [B, n, 3] clouds with the same batch size")
        if xyz1.size(1) != xyz2.size(1):
            raise _lib.IaptError("emd: the two point clouds should have the same size (emd_cuda.cu:238-241)")
        xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
        B, n, _ = xyz1.shape
        dev = xyz1.device
        L = _lib.lib()
        with _lib.on_device(dev):
            dist = torch.empty((B, n), dtype=torch.float32, device=dev)
            assignment = torch.empty((B, n), dtype=torch.int32, device=dev)
            ws_bytes = int(L.iapt_emd_workspace_bytes(B, n))
            ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
            _lib.check(L.iapt_emd_forward(_lib.ptr(xyz1), _lib.ptr(xyz2), B, n, float(eps), int(iters), _lib.ptr(dist),
                                          _lib.ptr(assignment), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)),
                       "iapt_emd_forward")
        ctx.save_for_backward(xyz1, xyz2, assignment)
        ctx.mark_non_differentiable(assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.to(torch.float32).contiguous()
        B, n, _ = xyz1.shape
        dev = xyz1.device
        with _lib.on_device(dev):
            gradxyz1 = torch.empty_like(xyz1)
            _lib.check(_lib.lib().iapt_emd_backward(_lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(graddist), _lib.ptr(assignment),
                                                    B, n, _lib.ptr(gradxyz1), _lib.stream_ptr(dev)), "iapt_emd_backward")
        return gradxyz1, torch.zeros_like(xyz2), None, None


class emdModule(nn.Module):
    def __init__(self):
        super(emdModule, self).__init__()

    def forward(self, input1, input2, eps, iters):
        return emdFunction.apply(input1, input2, eps, iters)
