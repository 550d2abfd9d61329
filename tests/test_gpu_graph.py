"""GPU parity: kNN consumers - graph feature gather/concat fwd+bwd, index_points, fps_downsample -
against golden outputs of the REFERENCE'S OWN DGCNN_Grouper methods (tests/golden/reference_grouper.npz) and the
torch restatement."""
import os

import numpy as np
import pytest
import torch

import oracle
from imgadapointr_b200.models import base_blocks as bb
from imgadapointr_b200.models.Transformer_utils import index_points, knn_point

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_grouper.npz")
tf = oracle.torch_formulation


def _t(z, name, dev):
    return torch.from_numpy(z[name]).to(dev)


def test_graph_feature_matches_reference_golden(dev):
    z = np.load(GOLD)
    x_q = _t(z, "x_q", dev).requires_grad_(True)
    x_k = _t(z, "x_k", dev).requires_grad_(True)
    idx = _t(z, "ref_idx", dev)
    out = bb.graph_feature(x_q, x_k, idx)
    assert torch.equal(out.detach().cpu(), torch.from_numpy(z["feature"]))  # gather + subtract: bit exact
    out.backward(_t(z, "grad_out", dev))
    for got, want in ((x_q.grad, z["grad_x_q"]), (x_k.grad, z["grad_x_k"])):
        want = torch.from_numpy(want)
        assert float((got.cpu() - want).abs().max()) <= 1e-5 * float(want.abs().max())


def test_get_graph_feature_end_to_end_and_determinism(dev):
    z = np.load(GOLD)
    k = z["ref_idx"].shape[-1]
    got = bb.get_graph_feature(k, _t(z, "coor_q", dev), _t(z, "x_q", dev), _t(z, "coor_k", dev), _t(z, "x_k", dev))
    want = torch.from_numpy(z["feature"])
    assert got.shape == want.shape
    # our kNN returns the same neighbour SET in (distance, index) order; every consumer reduces over k
    # (conv1x1 + GroupNorm + max, base_blocks.py:607-608), so compare order-independently
    assert torch.equal(got.cpu().sort(dim=-1).values, want.sort(dim=-1).values)
    assert torch.equal(got.cpu().max(dim=-1).values, want.max(dim=-1).values)
    # larger, duplicate-heavy case: forward equals the torch restatement for OUR idx, backward is bit-reproducible
    B, C, Nk, Nq, kk = 4, 35, 2048, 512, 16
    g = torch.Generator().manual_seed(3)
    coor_k = (torch.rand((B, 3, Nk), generator=g) - 0.5).to(dev)
    coor_q = coor_k[:, :, :Nq].contiguous()
    x_k = torch.randn((B, C, Nk), generator=g).to(dev)
    x_q = torch.randn((B, C, Nq), generator=g).to(dev)
    idx = knn_point(kk, coor_k.transpose(1, 2).contiguous(), coor_q.transpose(1, 2).contiguous())
    idx[:, :50] = 7  # heavy fan-in onto one source point
    ref = tf.get_graph_feature(kk, coor_q, x_q, coor_k, x_k, idx=idx)
    go = torch.randn(ref.shape, generator=g).to(dev)
    grads = []
    for _ in range(2):
        a, b = x_q.clone().requires_grad_(True), x_k.clone().requires_grad_(True)
        out = bb.graph_feature(a, b, idx)
        assert torch.equal(out, ref)
        out.backward(go)
        grads.append((a.grad.clone(), b.grad.clone()))
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    a, b = x_q.clone().requires_grad_(True), x_k.clone().requires_grad_(True)
    tf.get_graph_feature(kk, coor_q, a, coor_k, b, idx=idx).backward(go)
    assert float((grads[0][0] - a.grad).abs().max()) <= 1e-5 * float(a.grad.abs().max())
    assert float((grads[0][1] - b.grad).abs().max()) <= 1e-5 * float(b.grad.abs().max())


@pytest.mark.parametrize("B,C,Nk,Nq,k", [
    (3, 5, 300, 200, 16),     # whole-plane kernel, 256-thread CTAs
    (3, 6, 1000, 700, 8),     # 512-thread CTAs
    (2, 4, 2048, 2048, 16),   # 1024-thread CTAs, 128 KB plane
    (2, 3, 2048, 3000, 16),   # plane of 48000 floats: the largest that still fits
    (2, 3, 1500, 3200, 16),   # plane too large: separate grad_q / grad_k kernels, one-launch inverse index
    (2, 3, 700, 300, 5),      # k not a multiple of 4: separate kernels
    (2, 3, 3000, 500, 12),    # more than 2048 bins: count / scan / fill inverse index
])
def test_graph_feature_backward_every_kernel_path(dev, B, C, Nk, Nq, k):
    """grad_x_q / grad_x_k against float64 autograd of the torch restatement, and bit-reproducible, on every dispatch
    branch of iapt_graph_feature_grad (graph.cu: graph_feature_grad_launch / inverse_index_launch)."""
    g = torch.Generator().manual_seed(100 + Nk + Nq)
    x_k = torch.randn((B, C, Nk), generator=g).to(dev)
    x_q = torch.randn((B, C, Nq), generator=g).to(dev)
    idx = torch.randint(0, Nk, (B, Nq, k), generator=g).to(dev)
    idx[:, : Nq // 4] = idx[:, :1]  # heavy fan-in: a quarter of the queries share one neighbour list
    go = torch.randn((B, 2 * C, Nq, k), generator=g).to(dev)
    runs = []
    for _ in range(2):
        a, b = x_q.clone().requires_grad_(True), x_k.clone().requires_grad_(True)
        bb.graph_feature(a, b, idx).backward(go)
        runs.append((a.grad.clone(), b.grad.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    a, b = x_q.double().requires_grad_(True), x_k.double().requires_grad_(True)
    feat = torch.gather(b.unsqueeze(2).expand(B, C, Nq, Nk), 3, idx.unsqueeze(1).expand(B, C, Nq, k))
    torch.cat((feat - a.unsqueeze(-1), a.unsqueeze(-1).expand(B, C, Nq, k)), 1).backward(go.double())
    for got, want in ((runs[0][0], a.grad), (runs[0][1], b.grad)):
        assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # only one of the two gradients requested
    a = x_q.clone().requires_grad_(True)
    bb.graph_feature(a, x_k, idx).backward(go)
    assert torch.equal(a.grad, runs[0][0])
    b = x_k.clone().requires_grad_(True)
    bb.graph_feature(x_q, b, idx).backward(go)
    assert torch.equal(b.grad, runs[0][1])


def test_fps_downsample_matches_reference_golden(dev):
    z = np.load(GOLD)
    nc, nx = bb.fps_downsample(_t(z, "fd_coor", dev), _t(z, "fd_x", dev), z["fd_new_coor"].shape[-1])
    assert torch.equal(nc.cpu(), torch.from_numpy(z["fd_new_coor"])) and torch.equal(nx.cpu(), torch.from_numpy(z["fd_new_x"]))


def test_index_points_kernel_and_grad_path(dev):
    g = torch.Generator().manual_seed(5)
    pts = torch.randn((3, 400, 11), generator=g).to(dev)
    idx2 = torch.randint(0, 400, (3, 77), generator=g).to(dev)
    idx3 = torch.randint(0, 400, (3, 50, 8), generator=g).to(dev)
    bi = torch.arange(3, device=dev)
    assert torch.equal(index_points(pts, idx2), pts[bi.view(3, 1), idx2])
    assert torch.equal(index_points(pts, idx3), pts[bi.view(3, 1, 1), idx3])
    p = pts.clone().requires_grad_(True)
    index_points(p, idx3).sum().backward()  # differentiable path
    want = torch.zeros(3, 400, device=dev).scatter_add_(1, idx3.reshape(3, -1), torch.ones(3, 400, device=dev))
    assert torch.allclose(p.grad[..., 0], want)


def test_three_nn_interpolate_grouping_ball_query(dev):
    """The colder pointnet2_ops entry points (8f rank 4) against the torch restatement of upstream's semantics."""
    from imgadapointr_b200.pointnet2_ops import pointnet2_utils as pu

    g = torch.Generator().manual_seed(9)
    known = torch.rand((3, 300, 3), generator=g) - 0.5
    known[:, 10] = known[:, 4]  # duplicate -> tie, lowest index first
    unknown = torch.rand((3, 120, 3), generator=g) - 0.5
    unknown[:, 0] = known[:, 4]
    dist, idx = pu.three_nn(unknown.to(dev), known.to(dev))
    wd, wi = tf.three_nn(unknown, known)
    assert idx.dtype == torch.int32 and torch.equal(idx.cpu(), wi) and torch.allclose(dist.cpu(), wd, atol=1e-6)
    assert int(idx[0, 0, 0]) == 4 and int(idx[0, 0, 1]) == 10
    # three_interpolate fwd + bwd (weights as the reference forms them, Transformer_utils.py:229-231)
    feats = torch.randn((3, 7, 300), generator=g)
    recip = 1.0 / (wd + 1e-8)
    weight = recip / recip.sum(dim=2, keepdim=True)
    f = feats.to(dev).requires_grad_(True)
    out = pu.three_interpolate(f, idx, weight.to(dev))
    fr = feats.clone().requires_grad_(True)
    want = tf.three_interpolate(fr, wi, weight)
    assert torch.allclose(out.detach().cpu(), want.detach(), rtol=1e-5, atol=1e-6)
    go = torch.randn(want.shape, generator=g)
    out.backward(go.to(dev))
    want.backward(go)
    assert torch.allclose(f.grad.cpu(), fr.grad, rtol=1e-4, atol=1e-5)
    # grouping_operation fwd + bwd with repeated indices
    gi = torch.randint(0, 300, (3, 50, 9), generator=g, dtype=torch.int32)
    gi[:, :10] = 3
    f2 = feats.to(dev).requires_grad_(True)
    grouped = pu.grouping_operation(f2, gi.to(dev))
    fr2 = feats.clone().requires_grad_(True)
    wgrouped = tf.grouping_operation(fr2, gi)
    assert torch.equal(grouped.detach().cpu(), wgrouped.detach())
    go2 = torch.randn(wgrouped.shape, generator=g)
    grouped.backward(go2.to(dev))
    wgrouped.backward(go2)
    assert torch.allclose(f2.grad.cpu(), fr2.grad, rtol=1e-4, atol=1e-5)
    r1 = [pu.grouping_operation(f2.detach().requires_grad_(True), gi.to(dev)) for _ in range(2)]
    # ball_query
    bq = pu.ball_query(0.15, 8, known.to(dev), unknown.to(dev))
    assert bq.dtype == torch.int32 and torch.equal(bq.cpu(), tf.ball_query(0.15, 8, known, unknown))
    far = pu.ball_query(1e-4, 4, known.to(dev), (unknown + 5.0).to(dev))
    assert int(far.abs().sum()) == 0


@pytest.mark.parametrize("B,N,M,C,k", [(3, 100, 150, 384, 8), (2, 257, 64, 10, 5), (2, 64, 64, 516, 16), (1, 1, 3, 4, 2)])
def test_channels_last_graph_feature_and_index_points_grad(dev, B, N, M, C, k):
    """iapt_graph_feature_cl / _grad and iapt_index_points_grad against the torch restatement of DynamicGraphAttention's
    gather / subtract / concat (Transformer_utils.py:814-816): forward bit-exact, gradients against float64 autograd
    and bit-reproducible."""
    from imgadapointr_b200.models.Transformer_utils import graph_feature_cl
    g = torch.Generator().manual_seed(5 + N)
    q = torch.randn((B, N, C), generator=g).to(dev)
    v = torch.randn((B, M, C), generator=g).to(dev)
    idx = torch.randint(0, M, (B, N, k), generator=g).to(dev)
    idx[:, : max(1, N // 3)] = idx[:, :1]  # heavy fan-in
    want = tf.dynamic_graph_feature(q, v, idx)
    assert torch.equal(graph_feature_cl(q, v, idx), want)
    assert torch.equal(index_points(v, idx), torch.gather(v, 1, idx.reshape(B, -1, 1).expand(-1, -1, C)).reshape(B, N, k, C))
    go = torch.randn(want.shape, generator=g).to(dev)
    runs = []
    for _ in range(2):
        a, b = q.clone().requires_grad_(True), v.clone().requires_grad_(True)
        graph_feature_cl(a, b, idx).backward(go)
        runs.append((a.grad.clone(), b.grad.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    a, b = q.double().requires_grad_(True), v.double().requires_grad_(True)
    tf.dynamic_graph_feature(a, b, idx).backward(go.double())
    for got, ref in ((runs[0][0], a.grad), (runs[0][1], b.grad)):
        assert float((got.double() - ref).abs().max()) <= 1e-5 * max(float(ref.abs().max()), 1e-30)
    b1 = v.clone().requires_grad_(True)   # only grad_v requested
    graph_feature_cl(q, b1, idx).backward(go)
    assert torch.equal(b1.grad, runs[0][1])
    # index_points with a gradient: deterministic row sums instead of torch's atomicAdd scatter
    gi = torch.randn((B, N, k, C), generator=g).to(dev)
    outs = []
    for _ in range(2):
        b2 = v.clone().requires_grad_(True)
        index_points(b2, idx).backward(gi)
        outs.append(b2.grad.clone())
    assert torch.equal(outs[0], outs[1])
    b3 = v.double().requires_grad_(True)
    torch.gather(b3, 1, idx.reshape(B, -1, 1).expand(-1, -1, C)).reshape(B, N, k, C).backward(gi.double())
    assert float((outs[0].double() - b3.grad).abs().max()) <= 1e-5 * max(float(b3.grad.abs().max()), 1e-30)


def test_dynamic_graph_attention_forward_mirror(dev):
    """dynamic_graph_attention_forward (installed on the reference's DynamicGraphAttention by the drop-in) against the
    restated reference forward, self- and cross-attention and the denoise-token branch, values and gradients."""
    from imgadapointr_b200.models.Transformer_utils import dynamic_graph_attention_forward

    class Stub(torch.nn.Module):  # the attributes the reference module has (Transformer_utils.py:784-794)
        def __init__(self, dim, k):
            super().__init__()
            self.dim, self.k = dim, k
            self.knn_map = torch.nn.Sequential(torch.nn.Linear(dim * 2, dim), torch.nn.LeakyReLU(negative_slope=0.2))

    torch.manual_seed(11)
    m = Stub(48, 8).to(dev)
    g = torch.Generator().manual_seed(12)
    q = torch.randn((3, 200, 48), generator=g).to(dev)
    qp = (torch.rand((3, 200, 3), generator=g) - 0.5).to(dev)
    v = torch.randn((3, 300, 48), generator=g).to(dev)
    vp = (torch.rand((3, 300, 3), generator=g) - 0.5).to(dev)
    for kwargs, ref_kwargs in (({}, {}), ({"v": v, "v_pos": vp}, {"v": v, "v_pos": vp}), ({"denoise_length": 40}, {"denoise_length": 40})):
        a = q.clone().requires_grad_(True)
        out = dynamic_graph_attention_forward(m, a, qp, **kwargs)
        b = q.clone().requires_grad_(True)
        ref = tf.dynamic_graph_attention(m.knn_map, m.k, b, qp, **ref_kwargs)
        assert out.shape == ref.shape == (3, 200, 48)
        assert float((out - ref).detach().abs().max()) <= 1e-5 * float(ref.detach().abs().max())
        w = torch.randn(out.shape, generator=g).to(dev)
        (out * w).sum().backward()
        (ref * w).sum().backward()
        assert float((a.grad - b.grad).abs().max()) <= 1e-4 * float(b.grad.abs().max())
