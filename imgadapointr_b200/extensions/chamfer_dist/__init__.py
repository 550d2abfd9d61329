"""Drop-in for the reference package ``extensions.chamfer_dist`` (extensions/chamfer_dist/__init__.py:13-103).

Same class names, constructor (``ignore_zeros=False``), argument order (prediction first, ground truth second),
reductions and the ``ignore_zeros`` quirk (only when the batch size is 1; drops points whose coordinate SUM is 0,
__init__.py:37-41).  ``ChamferFunction`` calls the B200 library through ``imgadapointr_b200.chamfer`` exactly
where the reference calls its pybind module ``chamfer`` (__init__.py:16,24).
"""
import torch

from ... import _lib, chamfer


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_xyz1, grad_xyz2 = chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2,
                                                ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return grad_xyz1, grad_xyz2


class FusedChamferLoss(torch.autograd.Function):
    """Search + loss reduction + analytic gradient as ONE kernel launch (``iapt_chamfer_step``), no sqrt/mean autograd
    chain.  When an input needs a gradient the launch also forms d loss / d xyz for upstream
    1 and ``backward`` only applies the upstream scalar (``iapt_scale``); gradients nobody asked for (the ground truth
    in training) are never computed.  ``L2_split`` has two independent upstream scalars, so its gradient is a second
    launch from the saved indices (``iapt_chamfer_backward_fused``).  Same values as composing ChamferFunction with
    torch.sqrt / torch.mean (tests compare both paths); returns a 0-dim tensor (a pair for L2_split)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, kind):
        need1, need2 = bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1])
        one_launch = (need1 or need2) and kind != chamfer.LOSS_L2_SPLIT
        r = chamfer.step(xyz1, xyz2, kind, grad1=one_launch and need1, grad2=one_launch and need2)
        ctx.kind, ctx.one_launch, ctx.need = kind, one_launch, (need1, need2)
        if one_launch:
            ctx.save_for_backward(*[g for g in (r.grad1, r.grad2) if g is not None])
        elif need1 or need2:
            ctx.save_for_backward(xyz1, xyz2, r.idx1, r.idx2, r.dist1, r.dist2)
        if kind == chamfer.LOSS_L2_SPLIT:
            return r.loss[0], r.loss[1]
        return r.loss[0]

    @staticmethod
    def backward(ctx, *grads):
        need1, need2 = ctx.need
        if ctx.one_launch:
            saved = list(ctx.saved_tensors)
            g = grads[0].reshape(1)
            gx1 = chamfer.scale(saved.pop(0), g) if need1 else None
            gx2 = chamfer.scale(saved.pop(0), g) if need2 else None
            return gx1, gx2, None
        xyz1, xyz2, idx1, idx2, dist1, dist2 = ctx.saved_tensors
        if ctx.kind == chamfer.LOSS_L2_SPLIT:
            g = torch.stack([grads[0] if grads[0] is not None else torch.zeros_like(dist1[0, 0]),
                             grads[1] if grads[1] is not None else torch.zeros_like(dist1[0, 0])])
        else:
            g = grads[0].reshape(1)
        gx1, gx2 = chamfer.loss_backward(xyz1.detach().contiguous(), xyz2.detach().contiguous(), idx1, idx2, dist1,
                                         dist2, ctx.kind, g, need1, need2)
        return gx1, gx2, None


def _filter_zeros(xyz1, xyz2, ignore_zeros):
    batch_size = xyz1.size(0)
    if batch_size == 1 and ignore_zeros:
        non_zeros1 = torch.sum(xyz1, dim=2).ne(0)
        non_zeros2 = torch.sum(xyz2, dim=2).ne(0)
        xyz1 = xyz1[non_zeros1].unsqueeze(dim=0)
        xyz2 = xyz2[non_zeros2].unsqueeze(dim=0)
    return xyz1, xyz2


class _ChamferBase(torch.nn.Module):
    #: use the fused search+loss+gradient path (set False to compose ChamferFunction with torch.sqrt/mean exactly
    #: like the reference's module source does; both paths are parity-tested against each other)
    fused = True

    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def _dists(self, xyz1, xyz2):
        xyz1, xyz2 = _filter_zeros(xyz1, xyz2, self.ignore_zeros)
        return ChamferFunction.apply(xyz1, xyz2)

    def _fused(self, xyz1, xyz2, kind):
        xyz1, xyz2 = _filter_zeros(xyz1, xyz2, self.ignore_zeros)
        if xyz1.size(0) == 0 or xyz1.size(1) == 0 or xyz2.size(1) == 0:
            return None  # degenerate: let the composed path reproduce the reference's NaN-from-empty-mean behaviour
        ext = _lib.torch_ext()
        if ext is not None and xyz1.is_cuda and xyz1.dtype == torch.float32 and xyz2.dtype == torch.float32:
            # the same op as FusedChamferLoss below, as a C++ autograd node (csrc/torch_binding.cpp): no Python between
            # the module call and the launch
            out = ext.fused_chamfer_loss(xyz1, xyz2, kind)
            return (out[0], out[1]) if kind == chamfer.LOSS_L2_SPLIT else out[0]
        return FusedChamferLoss.apply(xyz1, xyz2, kind)


class ChamferDistanceL2(_ChamferBase):
    """mean(dist1) + mean(dist2)   (__init__.py:28-44)"""

    def forward(self, xyz1, xyz2):
        if self.fused and (out := self._fused(xyz1, xyz2, chamfer.LOSS_L2)) is not None:
            return out
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(_ChamferBase):
    """(mean(dist1), mean(dist2))   (__init__.py:46-62)"""

    def forward(self, xyz1, xyz2):
        if self.fused and (out := self._fused(xyz1, xyz2, chamfer.LOSS_L2_SPLIT)) is not None:
            return out
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(_ChamferBase):
    """(mean(sqrt(dist1)) + mean(sqrt(dist2))) / 2   (__init__.py:64-84)"""

    def forward(self, xyz1, xyz2):
        if self.fused and (out := self._fused(xyz1, xyz2, chamfer.LOSS_L1)) is not None:
            return out
        dist1, dist2 = self._dists(xyz1, xyz2)
        dist1 = torch.sqrt(dist1)
        dist2 = torch.sqrt(dist2)
        return (torch.mean(dist1) + torch.mean(dist2)) / 2


class ChamferDistanceL1_PM(_ChamferBase):
    """mean(sqrt(dist1))   (__init__.py:86-103)"""

    def forward(self, xyz1, xyz2):
        if self.fused and (out := self._fused(xyz1, xyz2, chamfer.LOSS_L1_PM)) is not None:
            return out
        dist1, _ = self._dists(xyz1, xyz2)
        dist1 = torch.sqrt(dist1)
        return torch.mean(dist1)
