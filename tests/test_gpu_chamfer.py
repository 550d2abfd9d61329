"""GPU parity: Chamfer forward / backward through the C ABI vs the CPU oracle (bit-exact dist + idx, grads 1e-5)."""
import numpy as np
import pytest
import torch

import oracle
from imgadapointr_b200 import chamfer
from imgadapointr_b200.extensions.chamfer_dist import (ChamferDistanceL1, ChamferDistanceL1_PM, ChamferDistanceL2,
                                                       ChamferDistanceL2_split, ChamferFunction)
from imgadapointr_b200.utils import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["auto", "8", "16"], autouse=True)
def tile_variant(request, monkeypatch):
    """Every test of this file runs three times: with the tile the heuristic picks, and with the 2048-point (QT=8) and
    4096-point (QT=16) register tiles forced (IAPT_FWD_VARIANT is read per launch) - so the kernel variant bench.py
    times also goes through the ragged, tie-heavy, extreme-magnitude and random-shape suites."""
    if request.param != "auto":
        monkeypatch.setenv("IAPT_FWD_VARIANT", request.param)
    return request.param


GRAD_RTOL = 1e-5  # north_star: Chamfer values and gradients within 1e-5 relative in FP32


def _check_forward(x1, x2, dev):
    d1, d2, i1, i2 = chamfer.forward(x1.to(dev), x2.to(dev))
    od1, od2, oi1, oi2 = oracle.chamfer_forward(x1.numpy(), x2.numpy())
    assert d1.dtype == torch.float32 and i1.dtype == torch.int32
    assert np.array_equal(d1.cpu().numpy(), od1)
    assert np.array_equal(d2.cpu().numpy(), od2)
    assert np.array_equal(i1.cpu().numpy(), oi1)
    assert np.array_equal(i2.cpu().numpy(), oi2)
    return d1, d2, i1, i2


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 5, 3), (3, 33, 257), (2, 513, 31), (1, 2048, 2049), (2, 2047, 4100),
                                   (4, 64, 128), (1, 3000, 700)])
def test_forward_ragged_uniform(dev, B, N, M):
    _check_forward(synthetic.uniform(B, N, 11 + N), synthetic.uniform(B, M, 13 + M), dev)


def test_forward_config1_pcn_shape(dev):
    """BASELINE config 1: partial 2048 (zero padded) vs gt 16384."""
    gt = synthetic.pcn_gt(1, 16384, 101)
    partial = synthetic.pcn_partial(gt, 2048, 102)
    _check_forward(partial, gt, dev)


@pytest.mark.parametrize("bits,dup,zero", [(6, 0.0, 0.0), (5, 0.25, 0.0), (6, 0.0, 0.25), (3, 0.1, 0.1)])
def test_forward_grid_ties_lowest_index(dev, bits, dup, zero):
    """Grid-quantised inputs: all arithmetic exact, many real ties -> the lowest index must win (chamfer.cu:47-50)."""
    x1 = synthetic.grid(2, 1500, 21, bits, dup, zero)
    x2 = synthetic.grid(2, 2300, 22, bits, dup, zero)
    d1, d2, i1, i2 = _check_forward(x1, x2, dev)
    # independent brute-force check of the tie rule in float64 (exact on the grid)
    D = ((x1[:, :, None, :].double() - x2[:, None, :, :].double()) ** 2).sum(-1)
    assert torch.equal(D.argmin(2).int(), i1.cpu()) or torch.equal(
        (D == D.min(2, keepdim=True).values).int().argmax(2).int(), i1.cpu())
    first2 = (D == D.min(1, keepdim=True).values).int().argmax(1).int()
    assert torch.equal(first2, i2.cpu())


def test_forward_identical_clouds_zero_distance(dev):
    x = synthetic.uniform(2, 1000, 5)
    d1, d2, i1, i2 = _check_forward(x, x.clone(), dev)
    assert float(d1.max()) == 0.0 and torch.equal(i1.cpu(), torch.arange(1000, dtype=torch.int32).expand(2, -1))


def test_forward_empty(dev):
    x = synthetic.uniform(2, 7, 1).to(dev)
    e = torch.zeros((2, 0, 3), device=dev)
    d1, d2, i1, i2 = chamfer.forward(x, e)
    assert d1.shape == (2, 7) and d2.shape == (2, 0) and float(d1.abs().sum()) == 0 and int(i1.abs().sum()) == 0
    gx1, gx2 = chamfer.backward(x, e, i1, i2, torch.ones_like(d1), torch.ones_like(d2))
    assert float(gx1.abs().sum()) == 0 and gx2.shape == (2, 0, 3)


def test_forward_noncontiguous_and_errors(dev):
    x = synthetic.uniform(2, 300, 3).to(dev)
    y = synthetic.uniform(2, 600, 4).to(dev)
    a = chamfer.forward(x, y[:, ::2])
    b = chamfer.forward(x, y[:, ::2].contiguous())
    assert all(torch.equal(u, v) for u, v in zip(a, b))
    with pytest.raises(Exception):
        chamfer.forward(x.cpu(), y.cpu())
    with pytest.raises(Exception):
        chamfer.forward(x.double(), y.double())
    with pytest.raises(Exception):
        chamfer.forward(x, y[:1])


def _oracle_grads(x1, x2, g1, g2):
    _, _, oi1, oi2 = oracle.chamfer_forward(x1, x2)
    return oracle.chamfer_backward(x1, x2, oi1, oi2, g1, g2)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("B,N,M", [(2, 64, 128), (2, 700, 300), (1, 2048, 16384), (3, 1000, 1000)])
def test_backward_matches_oracle(dev, B, N, M):
    x1, x2 = synthetic.uniform(B, N, 31), synthetic.uniform(B, M, 32)
    g1 = torch.randn((B, N), generator=torch.Generator().manual_seed(1))
    g2 = torch.randn((B, M), generator=torch.Generator().manual_seed(2))
    _, _, i1, i2 = chamfer.forward(x1.to(dev), x2.to(dev))
    gx1, gx2 = chamfer.backward(x1.to(dev), x2.to(dev), i1, i2, g1.to(dev), g2.to(dev))
    o1, o2 = _oracle_grads(x1.numpy(), x2.numpy(), g1.numpy(), g2.numpy())
    assert _rel(gx1.cpu().numpy(), o1) < GRAD_RTOL and _rel(gx2.cpu().numpy(), o2) < GRAD_RTOL


def test_backward_heavy_fan_in_and_determinism(dev):
    """All of xyz2 maps onto a collapsed xyz1 (an untrained network's output): worst-case fan-in, and the result must
    be bit-identical run to run (the reference's float atomics are not)."""
    B, N, M = 2, 512, 5000
    x1 = torch.zeros((B, N, 3))
    x1[:, 0] = 0.25
    x1[:, 1:] = 10.0 + torch.arange(N - 1).float()[None, :, None]
    x2 = synthetic.uniform(B, M, 41)
    g1 = torch.randn((B, N), generator=torch.Generator().manual_seed(3))
    g2 = torch.randn((B, M), generator=torch.Generator().manual_seed(4))
    a1, a2 = x1.to(dev), x2.to(dev)
    _, _, i1, i2 = chamfer.forward(a1, a2)
    assert int(i2.max()) == 0
    r = [chamfer.backward(a1, a2, i1, i2, g1.to(dev), g2.to(dev)) for _ in range(3)]
    assert all(torch.equal(r[0][0], q[0]) and torch.equal(r[0][1], q[1]) for q in r[1:])
    o1, o2 = _oracle_grads(x1.numpy(), x2.numpy(), g1.numpy(), g2.numpy())
    assert _rel(r[0][0].cpu().numpy(), o1) < GRAD_RTOL and _rel(r[0][1].cpu().numpy(), o2) < GRAD_RTOL


@pytest.mark.parametrize("B,N,M,distinct", [(48, 512, 16384, 4), (2, 3000, 16384, 1), (3, 700, 70000, 2)])
def test_backward_collapsed_prediction(dev, tile_variant, B, N, M, distinct):
    """An untrained network emits a collapsed prediction: all of xyz2 chooses one of a few prediction points, so a few
    bins have fan-ins of thousands (up to 70000: several bitmap windows).  Still deterministic and within 1e-5 of the
    oracle; VERDICT r01 measured 1.5 ms at the first shape with the old scan kernel."""
    if tile_variant != "auto":
        pytest.skip("backward only: independent of the search tile")
    x1 = torch.zeros((B, N, 3))
    x1[:] = 5.0 + torch.arange(N).float()[None, :, None]  # far away: never a nearest neighbour
    g = torch.Generator().manual_seed(7)
    x1[:, :distinct] = torch.rand((B, distinct, 3), generator=g) - 0.5
    x2 = synthetic.uniform(B, M, 43)
    g1 = torch.randn((B, N), generator=torch.Generator().manual_seed(3))
    g2 = torch.randn((B, M), generator=torch.Generator().manual_seed(4))
    a1, a2 = x1.to(dev), x2.to(dev)
    _, _, i1, i2 = chamfer.forward(a1, a2)
    assert int(i2.max()) < distinct
    r = [chamfer.backward(a1, a2, i1, i2, g1.to(dev), g2.to(dev)) for _ in range(2)]
    assert torch.equal(r[0][0], r[1][0]) and torch.equal(r[0][1], r[1][1])
    nb = min(B, 2)  # the oracle is sequential: two clouds suffice, all clouds are checked for determinism above
    o1, o2 = _oracle_grads(x1[:nb].numpy(), x2[:nb].numpy(), g1[:nb].numpy(), g2[:nb].numpy())
    # a float32 sum of thousands of terms: the bar is 1e-5 of the gradient scale plus 1e-6 of the summed magnitude of
    # the terms (the oracle accumulates in float64)
    oi2 = i2[:nb].cpu().numpy().astype(np.int64)
    mag = np.zeros((nb, N), np.float64)
    for c in range(nb):
        t = np.abs(2.0 * g2[c].numpy()[:, None] * (x2[c].numpy() - x1[c].numpy()[oi2[c]])).max(axis=1)
        np.add.at(mag[c], oi2[c], t)
    err1 = np.abs(r[0][0][:nb].cpu().numpy() - o1).max(axis=2)
    assert (err1 <= GRAD_RTOL * np.abs(o1).max() + 1e-6 * mag).all()
    assert _rel(r[0][1][:nb].cpu().numpy(), o2) < GRAD_RTOL
    # the fused loss step on the same collapsed input agrees with the two-step path
    st = chamfer.step(a1, a2, chamfer.LOSS_L2, grad1=True, grad2=True)
    w1 = torch.full((B, N), 1.0 / (B * N), device=dev)
    w2 = torch.full((B, M), 1.0 / (B * M), device=dev)
    p1, p2 = chamfer.backward(a1, a2, st.idx1, st.idx2, w1, w2)
    assert float((st.grad1 - p1).abs().max()) <= 1e-5 * float(p1.abs().max())
    assert float((st.grad2 - p2).abs().max()) <= 1e-5 * float(p2.abs().max())


def test_reference_gradcheck_restated(dev):
    """The reference's only test (extensions/chamfer_dist/test.py:23-29) is a float64 gradcheck on (4,64,3)/(4,128,3).
    Our kernels are FP32-only, so the same property is checked against central differences of a float64 loss built
    from the oracle's nearest neighbours."""
    x = synthetic.uniform(4, 64, 51)
    y = synthetic.uniform(4, 128, 52)
    xg = x.to(dev).requires_grad_(True)
    yg = y.to(dev).requires_grad_(True)
    d1, d2 = ChamferFunction.apply(xg, yg)
    (d1.sum() + d2.sum()).backward()

    def f(a, b):
        D = ((a[:, :, None, :] - b[:, None, :, :]) ** 2).sum(-1)
        return D.min(2).values.sum() + D.min(1).values.sum()

    a = x.double().requires_grad_(True)
    b = y.double().requires_grad_(True)
    f(a, b).backward()
    assert _rel(xg.grad.cpu().numpy(), a.grad.numpy()) < 1e-5
    assert _rel(yg.grad.cpu().numpy(), b.grad.numpy()) < 1e-5


@pytest.mark.parametrize("cls", [ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferDistanceL1_PM])
def test_loss_modules_match_cpu_formulation(dev, cls):
    pred, gt = synthetic.pcn_pair(3, 1024, 61)
    p = pred.to(dev).requires_grad_(True)
    out = cls()(p, gt.to(dev))
    tf = oracle.torch_formulation
    d1, d2, _, _ = tf.chamfer_nn(pred.double(), gt.double())
    want = {
        ChamferDistanceL1: lambda: (d1.sqrt().mean() + d2.sqrt().mean()) / 2,
        ChamferDistanceL2: lambda: d1.mean() + d2.mean(),
        ChamferDistanceL2_split: lambda: torch.stack([d1.mean(), d2.mean()]),
        ChamferDistanceL1_PM: lambda: d1.sqrt().mean(),
    }[cls]()
    got = torch.stack(list(out)) if isinstance(out, tuple) else out
    assert torch.allclose(got.cpu().double(), want, rtol=1e-5, atol=0)
    (got.sum()).backward()
    assert torch.isfinite(p.grad).all()


def test_ignore_zeros_quirk(dev):
    """ignore_zeros only acts for batch size 1 and drops points whose coordinate SUM is zero (__init__.py:37-41)."""
    gt = synthetic.pcn_gt(1, 2048, 71)
    partial = synthetic.pcn_partial(gt, 1024, 72)
    partial[0, 5] = torch.tensor([0.25, -0.5, 0.25])  # sums to zero without being the origin
    got = ChamferDistanceL1(ignore_zeros=True)(partial.to(dev), gt.to(dev))
    keep = partial[0].sum(1) != 0
    d1, d2, _, _ = oracle.torch_formulation.chamfer_nn(partial[:, keep].double(), gt.double())
    want = (d1.sqrt().mean() + d2.sqrt().mean()) / 2
    assert abs(float(got) - float(want)) <= 1e-5 * float(want)
    both = torch.cat([partial, partial]).to(dev)
    gt2 = torch.cat([gt, gt]).to(dev)
    assert float(ChamferDistanceL1(ignore_zeros=True)(both, gt2)) == float(ChamferDistanceL1()(both, gt2))


@pytest.mark.parametrize("cls", [ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferDistanceL1_PM])
@pytest.mark.parametrize("B,N,M", [(3, 1024, 1024), (2, 700, 2500), (1, 16384, 2048)])
def test_fused_loss_path_equals_composed_path(dev, cls, B, N, M):
    """The fused search+loss+gradient path (default) against the reference-style composition ChamferFunction +
    torch.sqrt/mean + autograd, values and gradients w.r.t. BOTH clouds, with non-trivial upstream weights."""
    pred, gt = synthetic.pcn_pair(B, max(N, M), 81)
    pred, gt = pred[:, :N].contiguous(), gt[:, :M].contiguous()
    res = {}
    for fused in (True, False):
        mod = cls()
        mod.fused = fused
        p = pred.to(dev).requires_grad_(True)
        g = gt.to(dev).requires_grad_(True)
        out = mod(p, g)
        out = torch.stack(list(out)) if isinstance(out, tuple) else out.reshape(1)
        w = torch.tensor([0.7, -1.3], device=dev)[: out.numel()]
        (out * w).sum().backward()
        res[fused] = (out.detach(), p.grad.clone(), g.grad.clone() if g.grad is not None else torch.zeros_like(g))
    a, b = res[True], res[False]
    assert torch.allclose(a[0], b[0], rtol=1e-6, atol=0)
    for x, y in ((a[1], b[1]), (a[2], b[2])):
        assert float((x - y).abs().max()) <= GRAD_RTOL * max(float(y.abs().max()), 1e-30)


def test_fused_l1_zero_distance_nan_quirk(dev):
    """SURVEY hard part 5: a predicted point that coincides exactly with a target makes the reference's L1 backward
    produce inf * 0 = NaN for that point; the fused path must reproduce it, not hide it."""
    pred, gt = synthetic.pcn_pair(1, 512, 91)
    pred[0, 7] = gt[0, 100]
    nan_sets = []
    for fused in (True, False):
        mod = ChamferDistanceL1()
        mod.fused = fused
        p = pred.to(dev).requires_grad_(True)
        mod(p, gt.to(dev)).backward()
        nan_sets.append(torch.isnan(p.grad).any(dim=-1).cpu())
    assert bool(nan_sets[0][0, 7]) and torch.equal(nan_sets[0], nan_sets[1])


@pytest.mark.parametrize("scale,offset", [(1e4, 3e4), (1e-3, 0.0), (1e-19, 0.0), (1.0, -7.5)])
def test_forward_extreme_magnitudes(dev, scale, offset):
    """Large, tiny (products underflow to denormals / zero) and offset coordinates: still bit-identical to the oracle."""
    x1 = synthetic.uniform(2, 600, 3) * scale + offset
    x2 = synthetic.uniform(2, 900, 4) * scale + offset
    _check_forward(x1.float().contiguous(), x2.float().contiguous(), dev)


def test_forward_many_tiny_clouds(dev):
    _check_forward(synthetic.uniform(300, 33, 5), synthetic.uniform(300, 17, 6), dev)


def test_nan_inputs_are_memory_safe(dev):
    """NaN / Inf coordinates are outside the parity contract but must not fault or return wild indices."""
    from imgadapointr_b200.models.Transformer_utils import knn_point
    from imgadapointr_b200.pointnet2_ops import pointnet2_utils

    x1 = synthetic.uniform(2, 700, 7)
    x2 = synthetic.uniform(2, 500, 8)
    x1[0, 5] = float("nan")
    x1[1] = float("nan")  # a whole cloud of NaNs
    x2[0, 9, 1] = float("inf")
    d1, d2, i1, i2 = chamfer.forward(x1.to(dev), x2.to(dev))
    torch.cuda.synchronize()
    assert int(i1.min()) >= 0 and int(i1.max()) < 500 and int(i2.min()) >= 0 and int(i2.max()) < 700
    gx1, gx2 = chamfer.backward(x1.to(dev), x2.to(dev), i1, i2, torch.ones_like(d1), torch.ones_like(d2))
    torch.cuda.synchronize()
    for k in (8, 32):
        ki = knn_point(k, x1.to(dev), x1.to(dev)[:, :64].contiguous())
        assert int(ki.min()) >= 0 and int(ki.max()) < 700
    fi = pointnet2_utils.furthest_point_sample(x1.to(dev), 32)
    torch.cuda.synchronize()
    assert int(fi.min()) >= 0 and int(fi.max()) < 700
    # finite clouds in the same batch are unaffected
    clean = chamfer.forward(x1[:1, 6:].contiguous().to(dev), x2[:1, 10:].contiguous().to(dev))
    od = oracle.chamfer_forward(x1[:1, 6:].numpy(), x2[:1, 10:].numpy())
    assert np.array_equal(clean[0].cpu().numpy(), od[0]) and np.array_equal(clean[2].cpu().numpy(), od[2])


@pytest.mark.parametrize("kind,cls", [("L1", ChamferDistanceL1), ("L2", ChamferDistanceL2),
                                      ("L2_split", ChamferDistanceL2_split), ("L1_PM", ChamferDistanceL1_PM)])
def test_graphed_loss_replay_matches_fused_path(dev, kind, cls):
    """CUDA-graph replay of the fused loss step (imgadapointr_b200.graphs): same kernels -> same bits, for several calls
    with different data through the same captured graph."""
    from imgadapointr_b200.graphs import GraphedChamferLoss

    graphed = GraphedChamferLoss(kind)
    for seed in (1, 2, 3):
        pred, gt = synthetic.pcn_pair(2, 2048, 100 + seed)
        gt = gt[:, :1500].contiguous()
        outs = []
        for mod in (graphed, cls()):
            p = pred.to(dev).requires_grad_(True)
            g = gt.to(dev).requires_grad_(True)
            out = mod(p, g)
            out = torch.stack(list(out)) if isinstance(out, tuple) else out.reshape(1)
            w = torch.tensor([0.7, -1.3], device=dev)[: out.numel()]
            (out * w).sum().backward()
            outs.append((out.detach().clone(), p.grad.clone(), g.grad.clone() if g.grad is not None else None))
        assert torch.equal(outs[0][0], outs[1][0])  # same kernels, same inputs: the loss is bit-identical
        # gradients: the graph stores d loss / d xyz for upstream 1 and scales afterwards (one extra rounding)
        for a, b in ((outs[0][1], outs[1][1]), (outs[0][2], outs[1][2])):
            if b is not None:
                assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max())


def test_forward_backward_random_shapes(dev):
    """40 seeded random (B, N, M) shapes - ragged sizes around the tile (2048/4096), stage (512) and chunk (32)
    boundaries - forward bit-exact, backward within 1e-5 of the oracle."""
    rng = np.random.default_rng(12345)
    edges = [1, 2, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 1025, 2047, 2048, 2049, 4095, 4097]
    for trial in range(40):
        B = int(rng.integers(1, 5))
        N = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(1, 5000))
        M = int(rng.choice(edges)) if rng.random() < 0.6 else int(rng.integers(1, 5000))
        x1 = synthetic.uniform(B, N, 1000 + trial)
        x2 = synthetic.uniform(B, M, 2000 + trial)
        if trial % 4 == 0:  # duplicates -> exact ties
            x2[:, M // 2:] = x2[:, : M - M // 2].clone()
        d1, d2, i1, i2 = _check_forward(x1, x2, dev)
        g1 = torch.randn((B, N), generator=torch.Generator().manual_seed(trial))
        g2 = torch.randn((B, M), generator=torch.Generator().manual_seed(trial + 1))
        gx1, gx2 = chamfer.backward(x1.to(dev), x2.to(dev), i1, i2, g1.to(dev), g2.to(dev))
        o1, o2 = _oracle_grads(x1.numpy(), x2.numpy(), g1.numpy(), g2.numpy())
        assert _rel(gx1.cpu().numpy(), o1) < GRAD_RTOL and _rel(gx2.cpu().numpy(), o2) < GRAD_RTOL, (B, N, M)
