"""Drop-in for the reference package ``extensions.chamfer_dist`` (extensions/chamfer_dist/__init__.py:13-103).

Same class names, constructor (``ignore_zeros=False``), argument order (prediction first, ground truth second),
reductions and the ``ignore_zeros`` quirk (only when the batch size is 1; drops points whose coordinate SUM is 0,
__init__.py:37-41).  ``ChamferFunction`` calls the B200 library through ``imgadapointr_b200.chamfer`` exactly
where the reference calls its pybind module ``chamfer`` (__init__.py:16,24).
"""
import torch

from ... import chamfer


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_xyz1, grad_xyz2 = chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
        return grad_xyz1, grad_xyz2


def _filter_zeros(xyz1, xyz2, ignore_zeros):
    batch_size = xyz1.size(0)
    if batch_size == 1 and ignore_zeros:
        non_zeros1 = torch.sum(xyz1, dim=2).ne(0)
        non_zeros2 = torch.sum(xyz2, dim=2).ne(0)
        xyz1 = xyz1[non_zeros1].unsqueeze(dim=0)
        xyz2 = xyz2[non_zeros2].unsqueeze(dim=0)
    return xyz1, xyz2


class _ChamferBase(torch.nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def _dists(self, xyz1, xyz2):
        xyz1, xyz2 = _filter_zeros(xyz1, xyz2, self.ignore_zeros)
        return ChamferFunction.apply(xyz1, xyz2)


class ChamferDistanceL2(_ChamferBase):
    """mean(dist1) + mean(dist2)   (__init__.py:28-44)"""

    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(_ChamferBase):
    """(mean(dist1), mean(dist2))   (__init__.py:46-62)"""

    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(_ChamferBase):
    """(mean(sqrt(dist1)) + mean(sqrt(dist2))) / 2   (__init__.py:64-84)"""

    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        dist1 = torch.sqrt(dist1)
        dist2 = torch.sqrt(dist2)
        return (torch.mean(dist1) + torch.mean(dist2)) / 2


class ChamferDistanceL1_PM(_ChamferBase):
    """mean(sqrt(dist1))   (__init__.py:86-103)"""

    def forward(self, xyz1, xyz2):
        dist1, _ = self._dists(xyz1, xyz2)
        dist1 = torch.sqrt(dist1)
        return torch.mean(dist1)
