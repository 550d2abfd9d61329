"""Dev: host-side call overhead of the public API at a tiny size (pure launch/Python cost)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import ChamferDistanceL1
from imgadapointr_b200.models.Transformer_utils import knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
dev = torch.device("cuda:0")
x = torch.rand(1, 64, 3, device=dev); y = torch.rand(1, 64, 3, device=dev)
def t(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
from imgadapointr_b200 import _lib
for kind in ("torch", "ctypes"):  # compiled PyTorch binding vs the ctypes table, same C ABI underneath
    _lib.set_binding(kind)
    d1, d2, i1, i2 = chamfer.forward(x, y)
    print(f"[{kind:6s} binding] host us/call:  chamfer.forward {t(lambda: chamfer.forward(x, y)):6.1f}   "
          f"chamfer.backward {t(lambda: chamfer.backward(x, y, i1, i2, d1, d2)):6.1f}   "
          f"chamfer.step(L1, grad1) {t(lambda: chamfer.step(x, y, chamfer.LOSS_L1, grad1=True)):6.1f}   "
          f"knn_point {t(lambda: knn_point(8, x, y)):6.1f}   fps {t(lambda: pointnet2_utils.furthest_point_sample(x, 16)):6.1f}",
          flush=True)
    l1m = ChamferDistanceL1()
    def fbm():
        p = x.detach().requires_grad_(True); l1m(p, y).backward()
    print(f"[{kind:6s} binding] ChamferDistanceL1 module fwd+bwd us/call {t(fbm):6.1f}", flush=True)
_lib.set_binding("torch")
print("chamfer.forward us/call", t(lambda: chamfer.forward(x, y)))
d1, d2, i1, i2 = chamfer.forward(x, y)
print("chamfer.backward us/call", t(lambda: chamfer.backward(x, y, i1, i2, d1, d2)))
l1 = ChamferDistanceL1()
def fb():
    p = x.detach().requires_grad_(True); l1(p, y).backward()
print("L1 fused fwd+bwd us/call", t(fb))
l1c = ChamferDistanceL1(); l1c.fused = False
def fb2():
    p = x.detach().requires_grad_(True); l1c(p, y).backward()
print("L1 composed fwd+bwd us/call", t(fb2))
from imgadapointr_b200.graphs import GraphedChamferLoss
gl = GraphedChamferLoss("L1")
def fb3():
    p = x.detach().requires_grad_(True); gl(p, y).backward()
print("L1 graphed fwd+bwd us/call", t(fb3))
gt16 = torch.rand(1, 16384, 3, device=dev); part = torch.rand(1, 2048, 3, device=dev)
def c1(mod):
    p = part.detach().requires_grad_(True); mod(p, gt16).backward()
print("C1 (2048 x 16384) fused us/call", t(lambda: c1(l1), 100), " graphed us/call", t(lambda: c1(gl), 100))
print("knn us/call", t(lambda: knn_point(8, x, y)))
print("fps us/call", t(lambda: pointnet2_utils.furthest_point_sample(x, 16)))
print("torch.empty us/call", t(lambda: torch.empty((1, 64), device=dev)))
X = torch.rand(1, 16384, 3, device=dev); Y = torch.rand(1, 16384, 3, device=dev)
for _ in range(5): chamfer.forward(X, Y)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): chamfer.forward(X, Y)
b.record(); torch.cuda.synchronize()
print("chamfer.forward B=1 16384^2 gpu us/call", a.elapsed_time(b) / 20 * 1e3)
