"""GPU: out-of-bounds WRITE detection without compute-sanitizer (closed on this pool): every output and workspace of
each C-ABI entry point is carved out of a larger guard-filled allocation, the kernel is run on ragged sizes, and the
guard bytes on both sides must be untouched."""
import ctypes

import pytest
import torch

from imgadapointr_b200 import _lib
from imgadapointr_b200.utils import synthetic

pytestmark = pytest.mark.gpu
GUARD = 4096  # bytes on each side
PATTERN = 0xA5


class Guarded:
    def __init__(self, nbytes, dev):
        self.n = int(nbytes)
        self.buf = torch.full((self.n + 2 * GUARD,), PATTERN, dtype=torch.uint8, device=dev)

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD

    def view(self, dtype, shape):
        return self.buf[GUARD:GUARD + self.n].view(dtype).view(shape)

    def intact(self):
        return bool((self.buf[:GUARD] == PATTERN).all()) and bool((self.buf[GUARD + self.n:] == PATTERN).all())


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 33, 7), (3, 513, 2049), (2, 4097, 300), (1, 9000, 8191)])
def test_chamfer_entry_points_stay_in_bounds(dev, B, N, M):
    L = _lib.lib()
    st = _lib.stream_ptr(dev)
    x1 = synthetic.uniform(B, N, 1).to(dev)
    x2 = synthetic.uniform(B, M, 2).to(dev)
    d1, d2 = Guarded(4 * B * N, dev), Guarded(4 * B * M, dev)
    i1, i2 = Guarded(4 * B * N, dev), Guarded(4 * B * M, dev)
    wsn = L.iapt_chamfer_forward_workspace_bytes(B, N, M)
    ws = Guarded(wsn, dev)
    _lib.check(L.iapt_chamfer_forward(x1.data_ptr(), x2.data_ptr(), B, N, M, d1.ptr, d2.ptr, i1.ptr, i2.ptr, ws.ptr, wsn,
                                      st), "fwd")
    torch.cuda.synchronize()
    assert all(g.intact() for g in (d1, d2, i1, i2, ws))
    gx1, gx2 = Guarded(12 * B * N, dev), Guarded(12 * B * M, dev)
    bwn = L.iapt_chamfer_backward_workspace_bytes(B, N, M)
    bws = Guarded(bwn, dev)
    _lib.check(L.iapt_chamfer_backward(x1.data_ptr(), x2.data_ptr(), i1.ptr, i2.ptr, d1.ptr, d2.ptr, B, N, M, gx1.ptr,
                                       gx2.ptr, bws.ptr, bwn, st), "bwd")
    assert L.iapt_chamfer_backward(x1.data_ptr(), x2.data_ptr(), i1.ptr, i2.ptr, d1.ptr, d2.ptr, B, N, M, gx1.ptr,
                                   gx2.ptr, None, 0, st) == -2  # workspace required
    gz1 = Guarded(12 * B * N, dev)  # one side only: identical bits for the side that is computed
    _lib.check(L.iapt_chamfer_backward(x1.data_ptr(), x2.data_ptr(), i1.ptr, i2.ptr, d1.ptr, d2.ptr, B, N, M, gz1.ptr,
                                       None, bws.ptr, bwn, st), "bwd-one-side")
    sums, loss, cnt = Guarded(16 * B, dev), Guarded(8, dev), Guarded(8 * B, dev)
    _lib.check(L.iapt_chamfer_loss(d1.ptr, d2.ptr, B, N, M, 1, sums.ptr, loss.ptr, st), "loss")
    _lib.check(L.iapt_fscore_counts(d1.ptr, d2.ptr, B, N, M, ctypes.c_float(0.1), cnt.ptr, st), "fs")
    gy1, gy2 = Guarded(12 * B * N, dev), Guarded(12 * B * M, dev)
    _lib.check(L.iapt_chamfer_backward_fused(x1.data_ptr(), x2.data_ptr(), i1.ptr, i2.ptr, d1.ptr, d2.ptr, 1, None, B, N,
                                             M, gy1.ptr, gy2.ptr, bws.ptr, bwn, st), "bwdf")
    # the whole step in one launch: every output and the workspace guarded, all optional outputs requested
    e1, e2 = Guarded(4 * B * N, dev), Guarded(4 * B * M, dev)
    j1, j2 = Guarded(4 * B * N, dev), Guarded(4 * B * M, dev)
    s2, c2, v2, l2 = Guarded(16 * B, dev), Guarded(8 * B, dev), Guarded(8 * B, dev), Guarded(8, dev)
    h1, h2 = Guarded(12 * B * N, dev), Guarded(12 * B * M, dev)
    swn = L.iapt_chamfer_step_workspace_bytes(B, N, M)
    sws = Guarded(swn, dev)
    _lib.check(L.iapt_chamfer_step(x1.data_ptr(), x2.data_ptr(), B, N, M, 1, 2, ctypes.c_float(0.1), e1.ptr, e2.ptr,
                                   j1.ptr, j2.ptr, s2.ptr, c2.ptr, v2.ptr, l2.ptr, None, h1.ptr, h2.ptr, sws.ptr, swn,
                                   st), "step")
    k1 = Guarded(12 * B * N, dev)
    one = torch.full((1,), 0.5, device=dev)
    _lib.check(L.iapt_scale(h1.ptr, one.data_ptr(), 3 * B * N, k1.ptr, st), "scale")
    torch.cuda.synchronize()
    assert all(g.intact() for g in (gx1, gx2, gz1, bws, sums, loss, cnt, gy1, gy2, d1, d2, i1, i2))
    assert all(g.intact() for g in (e1, e2, j1, j2, s2, c2, v2, l2, h1, h2, sws, k1))
    f3 = lambda g, n: g.view(torch.float32, (B, n, 3))
    assert torch.equal(f3(gx1, N), f3(gz1, N))
    assert torch.isfinite(f3(gx1, N)).all()
    # the one-launch step reproduces the separate entry points bit for bit (same search, same gradient order)
    assert torch.equal(e1.view(torch.float32, (B, N)), d1.view(torch.float32, (B, N)))
    assert torch.equal(j2.view(torch.int32, (B, M)), i2.view(torch.int32, (B, M)))
    assert torch.equal(f3(h1, N), f3(gy1, N)) and torch.equal(f3(h2, M), f3(gy2, M))
    assert torch.equal(f3(k1, N), f3(h1, N) * 0.5)
    assert int(v2.view(torch.int32, (B, 2))[:, 0].min()) == N


@pytest.mark.parametrize("B,N,m,k", [(2, 7, 7, 3), (3, 129, 64, 16), (2, 3000, 100, 33), (1, 20000, 50, 8)])
def test_fps_gather_knn_graph_stay_in_bounds(dev, B, N, m, k):
    L = _lib.lib()
    st = _lib.stream_ptr(dev)
    xyz = synthetic.uniform(B, N, 3).to(dev)
    idx = Guarded(4 * B * m, dev)
    wsn = L.iapt_fps_workspace_bytes(B, N, m)
    ws = Guarded(max(wsn, 4), dev)
    _lib.check(L.iapt_fps(xyz.data_ptr(), B, N, m, 1, idx.ptr, ws.ptr, wsn, st), "fps")
    out_xyz = Guarded(12 * B * m, dev)
    _lib.check(L.iapt_gather_xyz(xyz.data_ptr(), idx.ptr, B, N, m, out_xyz.ptr, st), "gxyz")
    C = 5
    feat = torch.randn((B, C, N), device=dev)
    out = Guarded(4 * B * C * m, dev)
    _lib.check(L.iapt_gather(feat.data_ptr(), idx.ptr, B, C, N, m, out.ptr, st), "gather")
    gfeat = Guarded(4 * B * C * N, dev)
    _lib.check(L.iapt_gather_grad(out.ptr, idx.ptr, B, C, N, m, gfeat.ptr, st), "ggrad")
    Q = min(N, 37)
    qry = xyz[:, :Q].contiguous()
    kidx, kdist = Guarded(8 * B * Q * k, dev), Guarded(4 * B * Q * k, dev)
    kwn = L.iapt_knn_workspace_bytes(B, N, Q, min(k, N))
    kws = Guarded(max(kwn, 4), dev)
    _lib.check(L.iapt_knn(xyz.data_ptr(), qry.data_ptr(), B, N, Q, min(k, N), kidx.ptr, kdist.ptr, kws.ptr, kwn, st), "knn")
    torch.cuda.synchronize()
    assert all(g.intact() for g in (idx, ws, out_xyz, out, gfeat, kidx, kdist, kws))
    kk = min(k, N)
    xq = torch.randn((B, C, Q), device=dev)
    gf = Guarded(4 * B * 2 * C * Q * kk, dev)
    _lib.check(L.iapt_graph_feature(xq.data_ptr(), feat.data_ptr(), kidx.ptr, B, C, Q, N, kk, gf.ptr, st), "gf")
    gq, gk = Guarded(4 * B * C * Q, dev), Guarded(4 * B * C * N, dev)
    wn = L.iapt_graph_feature_grad_workspace_bytes(B, Q, N, kk)
    w2 = Guarded(wn, dev)
    _lib.check(L.iapt_graph_feature_grad(gf.ptr, kidx.ptr, B, C, Q, N, kk, gq.ptr, gk.ptr, w2.ptr, wn, st), "gfg")
    ip = Guarded(4 * B * Q * kk * 3, dev)
    _lib.check(L.iapt_index_points(xyz.data_ptr(), kidx.ptr, B, N, 3, Q * kk, ip.ptr, st), "ip")
    # channels-last consumers: index_points gradient, DynamicGraphAttention feature + gradients (C = 5 scalar path,
    # C = 8 vector path)
    ipg = Guarded(4 * B * N * 3, dev)
    wi = L.iapt_graph_feature_grad_workspace_bytes(B, Q * kk, N, 1)
    w3 = Guarded(wi, dev)
    _lib.check(L.iapt_index_points_grad(ip.ptr, kidx.ptr, B, N, 3, Q * kk, ipg.ptr, w3.ptr, wi, st), "ipg")
    guards = [gf, gq, gk, w2, ip, kidx, ipg, w3]
    for Cc in (5, 8):
        tq = torch.randn((B, Q, Cc), device=dev)
        tv = torch.randn((B, N, Cc), device=dev)
        fcl = Guarded(4 * B * Q * kk * 2 * Cc, dev)
        _lib.check(L.iapt_graph_feature_cl(tq.data_ptr(), tv.data_ptr(), kidx.ptr, B, Q, N, Cc, kk, fcl.ptr, st), "gfcl")
        cq, cv = Guarded(4 * B * Q * Cc, dev), Guarded(4 * B * N * Cc, dev)
        w4 = Guarded(wn, dev)
        _lib.check(L.iapt_graph_feature_cl_grad(fcl.ptr, kidx.ptr, B, Q, N, Cc, kk, cq.ptr, cv.ptr, w4.ptr, wn, st), "gfclg")
        guards += [fcl, cq, cv, w4]
    torch.cuda.synchronize()
    assert all(g.intact() for g in guards)


@pytest.mark.parametrize("B,n,iters", [(1, 1, 3), (2, 33, 5), (3, 1025, 20), (1, 5000, 10)])
def test_emd_entry_points_stay_in_bounds(dev, B, n, iters):
    L = _lib.lib()
    st = _lib.stream_ptr(dev)
    g = torch.Generator().manual_seed(n)
    x1 = torch.rand((B, n, 3), generator=g).to(dev)
    x2 = torch.rand((B, n, 3), generator=g).to(dev)
    dist, asg = Guarded(4 * B * n, dev), Guarded(4 * B * n, dev)
    wn = L.iapt_emd_workspace_bytes(B, n)
    ws = Guarded(wn, dev)
    _lib.check(L.iapt_emd_forward(x1.data_ptr(), x2.data_ptr(), B, n, ctypes.c_float(0.005), iters, dist.ptr, asg.ptr,
                                  ws.ptr, wn, st), "emd")
    grad = Guarded(12 * B * n, dev)
    _lib.check(L.iapt_emd_backward(x1.data_ptr(), x2.data_ptr(), dist.ptr, asg.ptr, B, n, grad.ptr, st), "emdb")
    torch.cuda.synchronize()
    assert all(gd.intact() for gd in (dist, asg, ws, grad))
    a = asg.view(torch.int32, (B, n))
    assert int(a.min()) >= 0 and int(a.max()) < n
    assert L.iapt_emd_forward(x1.data_ptr(), x2.data_ptr(), B, n, ctypes.c_float(0.005), 0, dist.ptr, asg.ptr, ws.ptr, wn,
                              st) == -1  # at least one iteration
    assert L.iapt_emd_forward(x1.data_ptr(), x2.data_ptr(), B, n, ctypes.c_float(0.005), 5, dist.ptr, asg.ptr, None, 0,
                              st) == -2  # workspace required
