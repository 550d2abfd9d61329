"""GPU (needs >= 2 devices, skipped otherwise): several GPUs driven from ONE process, as the reference does under
nn.DataParallel (tools/runner.py:96) - kernel attributes must be configured per device and work must land on the
tensors' device, not the current one."""
import numpy as np
import pytest
import torch

import oracle
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import synthetic

pytestmark = pytest.mark.gpu


def test_second_device_in_same_process(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    x1, x2 = synthetic.uniform(2, 3000, 1), synthetic.uniform(2, 2500, 2)
    want = oracle.chamfer_forward(x1.numpy(), x2.numpy())
    for d in (0, 1, 0):
        device = torch.device("cuda", d)
        with torch.cuda.device(0):  # current device deliberately differs from the tensors' device for d == 1
            a = x1.to(device).requires_grad_(True)
            b = x2.to(device)
            got = chamfer.forward(a, b)
            assert all(t.device == device for t in got)
            assert np.array_equal(got[0].cpu().numpy(), want[0]) and np.array_equal(got[3].cpu().numpy(), want[3])
            loss = ChamferDistanceL1()(a, b)
            loss.backward()
            assert a.grad.device == device and torch.isfinite(a.grad).all()
            fi = pointnet2_utils.furthest_point_sample(b, 64)
            assert fi.device == device and np.array_equal(fi.cpu().numpy(), oracle.fps(x2.numpy(), 64, True))
            ki = knn_point(16, b, b[:, :100].contiguous())
            assert ki.device == device and np.array_equal(ki.cpu().numpy(), oracle.knn(16, x2.numpy(), x2.numpy()[:, :100])[0])
    torch.cuda.synchronize()
