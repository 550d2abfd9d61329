"""GPU: the drop-in scenario of INTEGRATION.md section 2(a) - reference-style consumer code that does `import chamfer`,
`from pointnet2_ops import pointnet2_utils`, `from extensions.chamfer_dist import ...` and star-imports knn_point runs
unchanged after dropin.install() / patch_reference_modules()."""
import subprocess
import sys
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONSUMER = r'''
import sys, types, torch
sys.path.insert(0, %r)
from imgadapointr_b200 import dropin
dropin.install()

# ---- what the reference's extensions/chamfer_dist/__init__.py does (:10,:13-25) -----------------------------------
import chamfer
class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2
    @staticmethod
    def backward(ctx, g1, g2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        a, b = chamfer.backward(xyz1, xyz2, idx1, idx2, g1, g2)
        return a, b
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = (torch.rand(2, 500, 3, generator=g) - 0.5).to(dev).requires_grad_(True)
y = (torch.rand(2, 700, 3, generator=g) - 0.5).to(dev)
d1, d2 = ChamferFunction.apply(x, y)
loss = (torch.mean(torch.sqrt(d1)) + torch.mean(torch.sqrt(d2))) / 2      # ChamferDistanceL1.forward (:82-84)
loss.backward()
from extensions.chamfer_dist import ChamferDistanceL1                      # models/*.py import this
x2 = x.detach().clone().requires_grad_(True)
l2 = ChamferDistanceL1()(x2, y); l2.backward()
assert abs(float(loss) - float(l2)) <= 1e-6 * float(loss)
assert float((x.grad - x2.grad).abs().max()) <= 1e-5 * float(x.grad.abs().max())

# ---- what utils/misc.py does (:10,:32-40) ---------------------------------------------------------------------------
from pointnet2_ops import pointnet2_utils
def fps(data, number):
    fps_idx = pointnet2_utils.furthest_point_sample(data, number)
    return pointnet2_utils.gather_operation(data.transpose(1, 2).contiguous(), fps_idx).transpose(1, 2).contiguous()
c = fps(y, 64)
assert c.shape == (2, 64, 3) and torch.equal(c[:, 0], y[:, 0])

# ---- a stand-in for the reference's models.base_blocks with its star-imported knn_point (base_blocks.py:9) ----------
models = types.ModuleType("models"); models.__path__ = []
bb = types.ModuleType("models.base_blocks")
def knn_point(nsample, xyz, new_xyz):   # the reference formulation, to be replaced in place
    d = torch.cdist(new_xyz, xyz) ** 2
    return torch.topk(d, nsample, dim=-1, largest=False, sorted=False)[1]
class DGCNN_Grouper:
    k = 16
    @staticmethod
    def fps_downsample(coor, x, num_group): raise RuntimeError("reference path")
    def get_graph_feature(self, coor_q, x_q, coor_k, x_k): raise RuntimeError("reference path")
bb.knn_point = knn_point; bb.DGCNN_Grouper = DGCNN_Grouper
sys.modules["models"] = models; sys.modules["models.base_blocks"] = bb
tu = types.ModuleType("models.Transformer_utils")
class DynamicGraphAttention(torch.nn.Module):   # attributes as in Transformer_utils.py:784-794, forward to be replaced
    def __init__(self, dim, k=10):
        super().__init__()
        self.dim, self.k = dim, k
        self.knn_map = torch.nn.Sequential(torch.nn.Linear(dim * 2, dim), torch.nn.LeakyReLU(negative_slope=0.2))
    def forward(self, q, q_pos, v=None, v_pos=None, idx=None, denoise_length=None): raise RuntimeError("reference path")
tu.DynamicGraphAttention = DynamicGraphAttention; tu.knn_point = knn_point
sys.modules["models.Transformer_utils"] = tu
patched = dropin.patch_reference_modules()
assert "models.base_blocks.knn_point" in patched and "models.base_blocks.DGCNN_Grouper.get_graph_feature" in patched
assert "models.Transformer_utils.DynamicGraphAttention.forward" in patched
att = tu.DynamicGraphAttention(12, k=4).to(dev)
tok = torch.randn(2, 700, 12, device=dev, requires_grad=True)
o = att(tok, y); o.sum().backward()
assert o.shape == (2, 700, 12) and tok.grad is not None and bool(torch.isfinite(tok.grad).all())
idx = bb.knn_point(8, y, y[:, :50].contiguous())
assert idx.dtype == torch.int64 and idx.shape == (2, 50, 8) and bool((idx[:, :, 0] == torch.arange(50, device=dev)).all())
coor = y.transpose(1, 2).contiguous(); feat = torch.randn(2, 6, 700, device=dev)
f = bb.DGCNN_Grouper().get_graph_feature(coor, feat, coor, feat)
assert f.shape == (2, 12, 700, 16)
nc, nx = bb.DGCNN_Grouper.fps_downsample(coor, feat, 100)
assert nc.shape == (2, 3, 100) and nx.shape == (2, 6, 100)
print("dropin ok", patched)
'''


def test_reference_style_consumer_runs_on_the_b200_path(dev):
    out = subprocess.run([sys.executable, "-c", CONSUMER % ROOT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "dropin ok" in out.stdout
