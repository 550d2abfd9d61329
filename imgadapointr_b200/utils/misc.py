"""Drop-in for the hot-path part of the reference's ``utils/misc.py``: ``fps`` (utils/misc.py:32-40)."""
import torch

from ..pointnet2_ops import pointnet2_utils


def fps(data, number):
    """data [B,N,3], number int -> [B,number,3] (contiguous).

    Reference: furthest_point_sample + gather_operation on a transposed copy + transpose back (4 launches, 2 copies).
    Here: the FPS kernel + one xyz gather.  When ``data`` requires grad the differentiable gather_operation path is
    used so gradients flow exactly as in the reference.
    """
    data = data.contiguous()
    fps_idx = pointnet2_utils.furthest_point_sample(data, number)
    if torch.is_grad_enabled() and data.requires_grad:
        return pointnet2_utils.gather_operation(data.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
    return pointnet2_utils.gather_xyz(data, fps_idx)
