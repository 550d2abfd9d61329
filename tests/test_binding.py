"""The compiled PyTorch binding (csrc/torch_binding.cpp -> iapt_torch_ext*.so) over the C ABI: the module the reference
builds as pybind `chamfer` (extensions/chamfer_dist/chamfer_cuda.cpp:22-39) plus the pointnet2_ops._ext / knn entry
points.  CPU: it imports, reports the ABI version, exposes the reference's surface and rejects what the reference would
silently misread.  GPU: bit-identical to the ctypes binding of the same library on every entry point."""
import numpy as np
import pytest
import torch

from imgadapointr_b200 import _lib, chamfer
from imgadapointr_b200.utils import synthetic


def _ext():
    from imgadapointr_b200 import iapt_torch_ext

    return iapt_torch_ext


def test_extension_loads_and_mirrors_the_reference_module_surface():
    e = _ext()
    assert e.abi_version() == _lib.ABI_VERSION
    for name in ("forward", "backward",  # chamfer_cuda.cpp:36-39
                 "furthest_point_sampling", "gather_points", "gather_points_grad",  # pointnet2_ops._ext
                 "knn", "square_distance", "step", "loss_backward", "scale", "gather_xyz"):
        assert callable(getattr(e, name)), name


def test_extension_rejects_cpu_wrong_dtype_and_rank_before_touching_the_gpu():
    e = _ext()
    a = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        e.forward(a, a)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        e.furthest_point_sampling(a, 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        e.knn(a, a, 2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        e.gather_points(torch.zeros(1, 3, 4), torch.zeros(1, 2, dtype=torch.int32))


def test_python_mirror_has_no_eager_path():
    """index_points / square_distance / knn_point raise on CPU tensors instead of computing with torch ops."""
    from imgadapointr_b200.models.Transformer_utils import index_points, knn_point, square_distance

    pts = torch.zeros(2, 5, 3)
    with pytest.raises(_lib.IaptError):
        index_points(pts, torch.zeros(2, 4, dtype=torch.int64))
    with pytest.raises(_lib.IaptError):
        square_distance(pts, pts)
    with pytest.raises(_lib.IaptError):
        knn_point(2, pts, pts)


@pytest.fixture
def both_bindings():
    """Runs a callable under the compiled binding and under ctypes, restoring the default afterwards."""
    def run(fn):
        out = {}
        try:
            for kind in ("torch", "ctypes"):
                _lib.set_binding(kind)
                out[kind] = fn()
        finally:
            _lib.set_binding("torch")
        return out["torch"], out["ctypes"]
    return run


def _same(a, b):
    if a is None or b is None:
        return a is None and b is None
    return a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)


@pytest.mark.gpu
def test_compiled_binding_equals_ctypes_chamfer(dev, both_bindings):
    p, g = synthetic.pcn_pair(3, 1500, seed=11)
    p, g = p.to(dev), g[:, :1100].contiguous().to(dev)
    t, c = both_bindings(lambda: chamfer.forward(p, g))
    assert all(_same(x, y) for x, y in zip(t, c))
    d1, d2, i1, i2 = t
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    t, c = both_bindings(lambda: chamfer.backward(p, g, i1, i2, g1, g2))
    assert all(_same(x, y) for x, y in zip(t, c))
    t, c = both_bindings(lambda: chamfer.backward(p, g, i1, i2, g1, g2, True, False))
    assert _same(t[0], c[0]) and t[1] is None and c[1] is None
    up = torch.tensor([0.37], device=dev)
    for kind in (chamfer.LOSS_L1, chamfer.LOSS_L2, chamfer.LOSS_L1_PM):
        t, c = both_bindings(lambda: chamfer.step(p, g, kind, sums=True, counts=True, valid=True, th=0.01, grad1=True,
                                                  grad2=(kind != chamfer.LOSS_L1_PM), grad_loss=up))
        for name in ("dist1", "dist2", "idx1", "idx2", "sums", "counts", "valid", "loss", "grad1", "grad2"):
            assert _same(getattr(t, name), getattr(c, name)), (kind, name)
    t, c = both_bindings(lambda: chamfer.loss_backward(p, g, i1, i2, d1, d2, chamfer.LOSS_L1, up))
    assert all(_same(x, y) for x, y in zip(t, c))
    t, c = both_bindings(lambda: chamfer.scale(p, up))
    assert _same(t, c)


@pytest.mark.gpu
def test_compiled_binding_equals_ctypes_fps_gather_knn(dev, both_bindings):
    from imgadapointr_b200.models.Transformer_utils import knn_point, square_distance
    from imgadapointr_b200.pointnet2_ops import pointnet2_utils as pu

    xyz = synthetic.pcn_partial(synthetic.pcn_gt(4, 4096, seed=5), 2048, seed=6).to(dev)
    t, c = both_bindings(lambda: pu.furthest_point_sample(xyz, 300))
    assert _same(t, c)
    idx = t
    feat = torch.rand(4, 7, 2048, device=dev)

    def gather():
        f = feat.clone().requires_grad_(True)
        out = pu.gather_operation(f, idx)
        out.backward(torch.ones_like(out) * 0.5)
        return out.detach(), f.grad

    t, c = both_bindings(gather)
    assert all(_same(x, y) for x, y in zip(t, c))
    t, c = both_bindings(lambda: pu.gather_xyz(xyz, idx))
    assert _same(t, c)
    for (k, R, Q) in ((16, 2048, 2048), (8, 300, 77), (32, 2048, 64)):
        t, c = both_bindings(lambda: knn_point(k, xyz[:, :R].contiguous(), xyz[:, :Q].contiguous(), return_dist=True))
        assert all(_same(x, y) for x, y in zip(t, c))
    t, c = both_bindings(lambda: square_distance(xyz[:, :333].contiguous(), xyz[:, :1001].contiguous()))
    assert _same(t, c)


@pytest.mark.gpu
def test_compiled_binding_checks_contiguity_dtype_and_device(dev):
    e = _ext()
    a = torch.rand(2, 64, 3, device=dev)
    with pytest.raises(RuntimeError, match="contiguous"):
        e.forward(a.transpose(0, 1).contiguous().transpose(0, 1), a)
    with pytest.raises(RuntimeError, match="Float|float"):
        e.forward(a.double(), a.double())
    with pytest.raises(RuntimeError, match="dims"):
        e.forward(a[0], a[0])
    with pytest.raises(RuntimeError, match=r"\[B, n, 3\]"):
        e.forward(torch.rand(2, 64, 4, device=dev), a)
    d1, d2, i1, i2 = e.forward(a, a)
    with pytest.raises(RuntimeError, match="Int|int"):
        e.backward(a, a, i1.long(), i2, d1, d2)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        e.backward(a, a, i1[:, :10].contiguous(), i2, d1, d2)
    with pytest.raises(RuntimeError, match="failed"):  # the C ABI's own argument check (k > R) surfaces as an exception
        e.knn(a, a, 65)


@pytest.mark.gpu
def test_square_distance_matches_the_reference_formulation(dev):
    """models/Transformer_utils.py:32-51 in its own op order (expansion form): agrees with the torch formulation to FP32
    rounding of that form, including C != 3 and ragged sizes."""
    from imgadapointr_b200.models.Transformer_utils import square_distance
    from oracle import torch_formulation as tf

    for (B, N, M, C) in ((2, 257, 1023, 3), (1, 8, 4, 3), (3, 100, 1024, 3), (2, 33, 65, 6), (1, 1, 1, 1)):
        g = torch.Generator().manual_seed(B * 1000 + N + M + C)
        src, dst = torch.rand((B, N, C), generator=g) - 0.5, torch.rand((B, M, C), generator=g) - 0.5
        want = tf.square_distance(src, dst)
        got = square_distance(src.to(dev), dst.to(dev)).cpu()
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-5, atol=2e-6), (B, N, M, C, (got - want).abs().max())
        want64 = tf.square_distance(src.double(), dst.double())
        assert (got.double() - want64).abs().max() < 2e-6


@pytest.mark.gpu
def test_cpp_autograd_loss_node_equals_python_function(dev, both_bindings):
    """The loss modules run through a C++ autograd node under the compiled binding (csrc/torch_binding.cpp:
    FusedChamferLossFn) and through the Python autograd.Function under ctypes: same loss bits, same gradient bits, for
    every module, with either or both inputs requiring a gradient, and nothing differentiated under no_grad."""
    from imgadapointr_b200.extensions.chamfer_dist import (ChamferDistanceL1, ChamferDistanceL1_PM, ChamferDistanceL2,
                                                            ChamferDistanceL2_split)

    p, g = synthetic.pcn_pair(2, 1300, seed=21)
    p, g = p.to(dev), g[:, :900].contiguous().to(dev)
    for cls in (ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL1_PM, ChamferDistanceL2_split):
        for req in ((True, False), (False, True), (True, True)):
            def run():
                a = p.clone().requires_grad_(req[0])
                b = g.clone().requires_grad_(req[1])
                out = cls()(a, b)
                loss = out[0] * 0.7 + out[1] * 1.3 if isinstance(out, tuple) else out * 0.9
                loss.backward()
                vals = torch.stack(list(out)) if isinstance(out, tuple) else out
                return vals.detach(), a.grad, b.grad
            t, c = both_bindings(run)
            assert all(_same(x, y) for x, y in zip(t, c)), (cls.__name__, req)
            if cls is ChamferDistanceL1_PM and req == (False, True):
                assert t[2] is not None and float(t[2].abs().max()) > 0  # gt also moves the one-sided loss
    with torch.no_grad():
        out = ChamferDistanceL1()(p.clone().requires_grad_(True), g)
    assert not out.requires_grad
    # non-contiguous input (a slice) and a double backward-free second call reuse nothing stale
    a = torch.cat([p, p], dim=1)[:, ::2].requires_grad_(True)
    ChamferDistanceL1()(a, g).backward()
    assert a.grad.shape == a.shape and torch.isfinite(a.grad).all()
