"""GPU parity: FPS, gather and kNN through the C ABI vs the CPU oracle (indices bit-exact)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from imgadapointr_b200.models.Transformer_utils import index_points, knn_point
from imgadapointr_b200.pointnet2_ops import pointnet2_utils
from imgadapointr_b200.utils import misc, synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("B,N,m", [(2, 7, 7), (3, 100, 33), (2, 129, 64), (2, 300, 300), (4, 1000, 256), (2, 2048, 512),
                                   (2, 3000, 100), (1, 5000, 64), (1, 16384, 2048), (1, 20000, 50), (30, 16384, 40), (48, 16384, 24), (24, 16390, 24),
                                   (80, 9000, 20), (76, 15000, 12)])
def test_fps_matches_oracle(dev, B, N, m):
    """Covers every kernel tier: registers (N <= 8192), thread-block clusters of 8/4/2 CTAs (N > 8192 while all clusters
    fit on the GPU), shared-memory tier (B too large for clusters, N <= 14000) and the global-memory fallback."""
    xyz = synthetic.uniform(B, N, 100 + N)
    for skip in (True, False):
        pointnet2_utils.furthest_point_sample.skip_near_origin = skip
        try:
            got = pointnet2_utils.furthest_point_sample(xyz.to(dev), m)
        finally:
            pointnet2_utils.furthest_point_sample.skip_near_origin = None
        assert got.dtype == torch.int32 and got.shape == (B, m)
        assert np.array_equal(got.cpu().numpy(), oracle.fps(xyz.numpy(), m, skip))


def test_fps_zero_padded_partials_and_prefix_property(dev):
    gt = synthetic.pcn_gt(4, 16384, 303)
    xyz = synthetic.pcn_partial(gt, 2048, 304)  # zero padded like datasets/data_transforms.py:63-65
    i512 = pointnet2_utils.furthest_point_sample(xyz.to(dev), 512).cpu()
    assert np.array_equal(i512.numpy(), oracle.fps(xyz.numpy(), 512, True))
    # padded (origin) points are never selected, except index 0 if it happens to be one
    sel = torch.gather(xyz, 1, i512.long().unsqueeze(-1).expand(-1, -1, 3))
    assert int(((sel ** 2).sum(-1)[:, 1:] <= 1e-3).sum()) == 0
    # the 256- and 64-point samples of the same cloud are prefixes of the 512-point one
    i64 = pointnet2_utils.furthest_point_sample(xyz.to(dev), 64).cpu()
    assert torch.equal(i64, i512[:, :64])


def test_fps_ties_grid_and_degenerate(dev):
    xyz = synthetic.grid(3, 1024, 7, bits=3, dup_frac=0.3)
    got = pointnet2_utils.furthest_point_sample(xyz.to(dev), 200).cpu().numpy()
    assert np.array_equal(got, oracle.fps(xyz.numpy(), 200, True))
    zeros = torch.zeros((2, 64, 3))
    assert int(pointnet2_utils.furthest_point_sample(zeros.to(dev), 16).abs().sum()) == 0  # nothing eligible -> 0


@pytest.mark.parametrize("N,m", [(100, 40), (512, 256), (2048, 512), (4000, 300), (9000, 200), (16384, 256), (20000, 64)])
def test_fps_upstream_tie_order_every_tier(dev, N, m):
    """IAPT_FPS_TIE_UPSTREAM (the published kernel's (k mod block, k) order among equal maxima) against the oracle's
    restatement of it, on tie-heavy grids, through every kernel tier (registers, shared memory, cluster, global); the
    default order is unaffected.  On these inputs the two orders select different indices."""
    fps = pointnet2_utils.furthest_point_sample
    g = np.random.default_rng(N + m)
    grid = (np.round(g.uniform(-0.5, 0.5, (2, N, 3)) * 8) / 8 + 2.0).astype(np.float32)
    x = torch.from_numpy(grid).to(dev)
    low = fps(x, m).cpu().numpy()
    assert np.array_equal(low, oracle.fps(grid, m, True))
    fps.tie_order = "upstream"
    try:
        up = fps(x, m).cpu().numpy()
    finally:
        fps.tie_order = "lowest"
    want, stats = oracle.fps_ex(grid, m, True, "upstream")
    assert np.array_equal(up, want)
    assert stats[:, 0].sum() > 0
    # a block at least as large as the cloud gives every thread one point: the tree then keeps the lowest index too
    assert np.array_equal(up, low) == (N == 512)


def test_fps_upstream_tie_order_on_the_demo_clouds(dev):
    """The reference's demo clouds (zero-padded to 2048) at the grouper's sizes: both tie orders equal their oracles; on
    the lamp (03636649) they differ in 41 of 512 indices (exact ties between different points of a symmetric mesh)."""
    import glob
    import os

    fps = pointnet2_utils.furthest_point_sample
    differing = 0
    for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sample_*.npz"))):
        xyz = np.load(f)["xyz"][None]
        x = torch.from_numpy(xyz).to(dev)
        low = fps(x, 512).cpu().numpy()
        fps.tie_order = "upstream"
        try:
            up = fps(x, 512).cpu().numpy()
        finally:
            fps.tie_order = "lowest"
        assert np.array_equal(low, oracle.fps_ex(xyz, 512, True, "lowest")[0])
        assert np.array_equal(up, oracle.fps_ex(xyz, 512, True, "upstream")[0])
        differing += int((low != up).any())
    assert differing >= 1


def test_fps_input_checks(dev):
    xyz = synthetic.uniform(2, 64, 1)
    with pytest.raises(Exception):
        pointnet2_utils.furthest_point_sample(xyz, 8)  # CPU tensor: no CPU path
    with pytest.raises(Exception):
        pointnet2_utils.furthest_point_sample(xyz.to(dev).transpose(1, 2), 8)  # non-contiguous / wrong shape


def test_gather_forward_backward_and_misc_fps(dev):
    B, C, N, m = 3, 35, 500, 128
    g = torch.Generator().manual_seed(5)
    feat = torch.randn((B, C, N), generator=g)
    idx = torch.randint(0, N, (B, m), generator=g, dtype=torch.int32)
    idx[:, 5] = idx[:, 4]  # duplicates -> fan-in in the backward
    idx[:, 6] = idx[:, 4]
    f = feat.to(dev).requires_grad_(True)
    out = pointnet2_utils.gather_operation(f, idx.to(dev))
    assert np.array_equal(out.detach().cpu().numpy(), oracle.gather(feat.numpy(), idx.numpy()))
    go = torch.randn((B, C, m), generator=g)
    out.backward(go.to(dev))
    want = oracle.gather_grad(go.numpy(), idx.numpy(), N)
    assert np.allclose(f.grad.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    xyz = synthetic.uniform(2, 700, 9)
    got = misc.fps(xyz.to(dev), 50)
    oi = oracle.fps(xyz.numpy(), 50, True)
    want_xyz = np.take_along_axis(xyz.numpy(), oi[:, :, None].astype(np.int64), axis=1)
    assert got.is_contiguous() and np.array_equal(got.cpu().numpy(), want_xyz)
    xg = xyz.to(dev).requires_grad_(True)
    misc.fps(xg, 50).sum().backward()
    assert float(xg.grad.sum()) == 2 * 50 * 3


def _knn_check(xyz, new_xyz, k, dev, exact_formulation=False):
    got, gd = knn_point(k, xyz.to(dev), new_xyz.to(dev), return_dist=True)
    assert got.dtype == torch.int64 and got.shape == (xyz.size(0), new_xyz.size(1), k)
    want, wd = oracle.knn(k, xyz.numpy(), new_xyz.numpy())
    assert np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(gd.cpu().numpy(), wd)
    return got


@pytest.mark.parametrize("R,Q,k", [(2048, 2048, 16), (2048, 512, 16), (512, 512, 16), (512, 256, 16), (256, 256, 8),
                                   (512, 512, 8), (256, 512, 8), (576, 64, 8), (16384, 64, 32), (100, 37, 5), (33, 1, 33),
                                   (3000, 70, 64), (64, 10, 1)])
def test_knn_matches_oracle_config3_shapes(dev, R, Q, k):
    xyz = synthetic.uniform(2, R, 200 + R)
    new_xyz = synthetic.uniform(2, Q, 300 + Q) if Q != R else xyz
    _knn_check(xyz, new_xyz, k, dev)


@pytest.mark.parametrize("B,R,Q,k,kind", [
    (16, 2048, 512, 16, "uniform"),   # select path, 4 warps per 32 queries, each a 512-reference slice
    (16, 512, 512, 16, "grid"),       # 4 slices of one tile each, heavy exact ties
    (32, 512, 256, 16, "zeros"),      # zero-padded tail: hundreds of references tie at the threshold
    (32, 256, 256, 8, "uniform"),     # k = 8 instantiation, 2 slices
    (16, 256, 512, 8, "grid"),
    (16, 1001, 600, 16, "zeros"),     # R not a multiple of 4 (scalar staging), ragged last tile and query block
    (9, 1500, 1000, 13, "grid"),      # k between the template sizes, last slice shorter than the others
    (64, 2048, 1024, 16, "zeros"),    # one warp per 32 queries, sampled first sweep
    (300, 64, 32, 16, "uniform"),     # smallest reference count the select path takes
])
def test_knn_select_path_slices_and_ties(dev, knn_path, B, R, Q, k, kind):
    """Batches large enough (B*ceil(Q/32) >= 256 warps) for the thread-per-query select kernel, bit for bit against
    the oracle: every slice count, both k instantiations, unaligned clouds, zero padding and exact-tie grids."""
    if kind == "uniform":
        xyz = synthetic.uniform(B, R, 900 + R)
    elif kind == "grid":
        xyz = synthetic.grid(B, R, 901 + R, bits=3, dup_frac=0.2)
    else:
        xyz = synthetic.grid(B, R, 902 + R, bits=10, zero_frac=0.4)
    new_xyz = xyz[:, :Q].contiguous() if Q <= R and kind != "uniform" else synthetic.uniform(B, Q, 903 + Q)
    _knn_check(xyz, new_xyz, k, dev)


@pytest.mark.parametrize("B,R,Q,near", [
    (64, 2048, 1024, "odd_groups"),   # one warp per 32 queries, sampled sweep: dense cluster hidden in the unsampled groups
    (16, 1000, 512, "few_classes"),   # full sweep: the cluster sits in 4 of the 32 classes, the other minima are far
    (8, 2048, 1024, "few_classes8"),  # 4 slices: 8 classes per slice hold 128 near references each
])
def test_knn_select_overflow_falls_back_exactly(dev, knn_path, B, R, Q, near):
    """Adversarial index layouts that leave the class-minima bound loose, so queries near the cluster collect more
    than the 112-entry candidate list holds and take the in-kernel exact fallback; the other queries of the same
    warps stay on the fast path.  Bit-exact against the oracle either way."""
    g = torch.Generator().manual_seed(77 + R)
    xyz = torch.rand((B, R, 3), generator=g) - 0.5
    r = torch.arange(R)
    if near == "odd_groups":
        sel = ((r // 4) % 2 == 1) & (r < 1600)          # 800 references of the groups sweep 1 skips
        sel &= torch.rand(R, generator=g) < 0.3          # ~240 of them
    elif near == "few_classes":
        sel = (r % 32) < 4                               # 125 references in classes 0..3
    else:
        sel = (r % 32) < 8                               # 128 per 512-reference slice
    n = int(sel.sum())
    centre = torch.tensor([0.31, -0.27, 0.11])
    xyz[:, sel] = centre + (torch.rand((B, n, 3), generator=g) - 0.5) * 2e-3
    new_xyz = torch.rand((B, Q, 3), generator=g) - 0.5
    hot = torch.rand(Q, generator=g) < 0.1               # a few queries per warp sit inside the cluster
    new_xyz[:, hot] = centre + (torch.rand((B, int(hot.sum()), 3), generator=g) - 0.5) * 2e-3
    _knn_check(xyz, new_xyz, 16, dev)
    _knn_check(xyz, new_xyz, 7, dev)


@pytest.fixture(params=["culled-4", "culled-2", "culled-1", "select"])
def knn_path(request, monkeypatch):
    """Shapes with 1024 <= R <= 4096 and k <= 16 can take the Hilbert-binned, box-culled kernel (IAPT_KNN_PRUNED=1 forces
    it; by default for dense queries in enough warps; IAPT_KNN_CULL_GROUP = tasks per CTA: 4 warps that help each other through
    shared-memory deques, or 1 = no helping) or the thread-per-query select kernel (IAPT_KNN_PRUNED=2) - the variables
    are read per call and every variant must match the oracle bit for bit."""
    if request.param == "select":
        monkeypatch.setenv("IAPT_KNN_PRUNED", "2")
    else:
        monkeypatch.setenv("IAPT_KNN_PRUNED", "1")
        monkeypatch.setenv("IAPT_KNN_CULL_GROUP", request.param.split("-")[1])
    return request.param


@pytest.mark.parametrize("B,R,Q,k,kind", [
    (3, 1024, 1024, 16, "self"),        # smallest reference count of the pruned path, self-query (one sort)
    (2, 1025, 33, 9, "uniform"),        # ragged last tile and query block
    (2, 2047, 700, 16, "pcn"),
    (2, 4096, 4096, 16, "self"),        # largest cloud one CTA sorts
    (3, 3000, 1, 1, "uniform"),         # a single query, k = 1
    (2, 2048, 31, 5, "far"),            # queries outside the references' bounding box: no tile overlaps their box
    (4, 2048, 2048, 16, "zeros_self"),  # zero-padded partial: hundreds of duplicates at the origin tie at tau = 0
    (2, 1500, 400, 16, "grid"),         # quantised coordinates: exact ties between DIFFERENT positions
    (2, 1200, 64, 8, "same"),           # every reference identical: all distances tie
    (2, 2048, 512, 16, "fps"),          # queries = FPS centres of the cloud (the grouper's 2048 -> 512 call)
    (48, 2048, 2048, 16, "self"),       # config 3's largest call at full batch
    (2, 4096, 4096, 16, "zeros_self"),  # 2048 duplicates in ONE cell: the bin kernel's re-ordering of a multi-tile cell
    (2, 2048, 300, 16, "dups"),         # 12 distinct points, each ~170 times: more multi-tile cells than get re-ordered
    (2, 2048, 2048, 16, "scaled"),      # coordinates ~3e4 +- 1e4: the prefilter's margin against FP32 rounding
    (2, 2048, 2048, 12, "tiny"),        # coordinates ~1e-19: squared distances in the subnormal range
])
def test_knn_pruned_path_matches_oracle(dev, knn_path, B, R, Q, k, kind):
    g = torch.Generator().manual_seed(1000 + R + Q + k)
    if kind in ("self", "uniform", "far", "fps"):
        xyz = torch.rand((B, R, 3), generator=g) - 0.5
    elif kind == "scaled":
        xyz = (torch.rand((B, R, 3), generator=g) - 0.5) * 2e4 + 3e4
    elif kind == "tiny":
        xyz = (torch.rand((B, R, 3), generator=g) - 0.5) * 2e-19
    elif kind == "dups":
        base = torch.rand((B, 12, 3), generator=g) - 0.5
        xyz = torch.gather(base, 1, torch.randint(0, 12, (B, R, 1), generator=g).expand(-1, -1, 3)).contiguous()
    elif kind == "pcn":
        xyz = synthetic.pcn_gt(B, R, 11)
    elif kind == "zeros_self":
        xyz = synthetic.pcn_partial(synthetic.pcn_gt(B, 8192, 12), R, 13)
        xyz[:, R // 2:] = 0.0
    elif kind == "grid":
        xyz = synthetic.grid(B, R, 14, bits=3, dup_frac=0.2)
    else:
        xyz = torch.full((B, R, 3), 0.25)
    if kind in ("self", "zeros_self", "scaled", "tiny"):
        new_xyz = xyz
    elif kind == "far":
        new_xyz = torch.rand((B, Q, 3), generator=g) * 0.1 + 3.0
    elif kind == "fps":
        idx = torch.from_numpy(oracle.fps(xyz.numpy(), Q, True)).long()
        new_xyz = torch.gather(xyz, 1, idx[..., None].expand(-1, -1, 3)).contiguous()
    elif kind == "grid":
        new_xyz = xyz[:, :Q].contiguous()
    else:
        new_xyz = torch.rand((B, Q, 3), generator=g) - 0.5
    if B > 8:  # the oracle is the slow side: check 4 clouds of the full batch
        a, q = xyz.to(dev), (xyz.to(dev) if new_xyz is xyz else new_xyz.to(dev))
        got, gd = knn_point(k, a, q, return_dist=True)
        oi, od = oracle.knn(k, xyz[:4].numpy(), new_xyz[:4].numpy())
        assert np.array_equal(got[:4].cpu().numpy(), oi) and np.array_equal(gd[:4].cpu().numpy(), od)
        again, _ = knn_point(k, a, q, return_dist=True)
        assert torch.equal(got, again)
        return
    if new_xyz is xyz:  # the same tensor object on the device: the self-query shortcut (one sort for both sides)
        a = xyz.to(dev)
        got, gd = knn_point(k, a, a, return_dist=True)
        oi, od = oracle.knn(k, xyz.numpy(), xyz.numpy())
        assert np.array_equal(got.cpu().numpy(), oi) and np.array_equal(gd.cpu().numpy(), od)
    else:
        _knn_check(xyz, new_xyz, k, dev)


@pytest.mark.timeout(120)
def test_knn_nan_inf_inputs_neither_hang_nor_leave_the_index_range(dev, knn_path):
    """NaN / Inf coordinates are outside the contract (the reference's topk gives arbitrary neighbours for them), but
    every path must finish and return indices a consumer can gather with."""
    g = torch.Generator().manual_seed(77)
    xyz = torch.rand((3, 2048, 3), generator=g) - 0.5
    xyz[0, 5] = float("nan")
    xyz[1, 7, 0] = float("inf")
    xyz[1, 100:140] = float("nan")
    xyz[2] = float("nan")
    a = xyz.to(dev)
    idx = knn_point(16, a, a)
    torch.cuda.synchronize()
    assert int(idx.min()) >= 0 and int(idx.max()) < 2048


def test_knn_zero_padded_and_grid_ties(dev, knn_path):
    gt = synthetic.pcn_gt(2, 4096, 303)
    xyz = synthetic.pcn_partial(gt, 2048, 305)
    _knn_check(xyz, xyz, 16, dev)
    g = synthetic.grid(2, 1500, 3, bits=4, dup_frac=0.2)
    got = _knn_check(g, g[:, :400].contiguous(), 16, dev)
    # on the exact grid the reference formulation's neighbour SET must agree wherever the k-th distance is not tied
    ref = oracle.torch_formulation.knn_point(16, g, g[:, :400].contiguous())
    D = oracle.torch_formulation.square_distance(g[:, :400].double(), g.double())
    kth = torch.sort(D, dim=-1).values
    untied = kth[..., 15] < kth[..., 16]
    a = torch.sort(got.cpu(), dim=-1).values[untied]
    b = torch.sort(ref, dim=-1).values[untied]
    assert untied.any() and torch.equal(a, b)


def test_knn_noncontiguous_slices_and_index_points(dev):
    pos = synthetic.uniform(2, 576, 17).to(dev)
    a = knn_point(8, pos[:, :512], pos[:, :512])  # Transformer_utils.py:839 passes slices
    b = knn_point(8, pos[:, :512].contiguous(), pos[:, :512].contiguous())
    assert torch.equal(a, b)
    pts = torch.randn((2, 512, 7), device=dev)
    got = index_points(pts, a)
    bi = torch.arange(2, device=dev).view(2, 1, 1).expand_as(a)
    assert torch.equal(got, pts[bi, a, :])
    with pytest.raises(Exception):
        knn_point(600, pos[:, :512], pos[:, :512])


def test_demo_sample_fixtures(dev):
    """Real PCN-style partial clouds (the reference's demo/samples/*/00.pcd, zero padded to 2048 like the dataset
    transform) with golden oracle outputs generated by tests/golden/make_golden.py."""
    files = sorted(glob.glob(os.path.join(GOLD, "sample_*.npz")))
    if not files:
        pytest.skip("no golden fixtures")
    for f in files:
        z = np.load(f)
        xyz = torch.from_numpy(z["xyz"])[None]
        fi = pointnet2_utils.furthest_point_sample(xyz.to(dev), 512).cpu().numpy()
        assert np.array_equal(fi[0], z["fps512"])
        ki = knn_point(16, xyz.to(dev), xyz.to(dev)).cpu().numpy()
        assert np.array_equal(ki[0], z["knn16"])


def test_fps_prefix_reuse_is_exact_and_safe(dev):
    """inside a ``prefix_reuse()`` scope repeated sampling of the SAME unmodified tensor is served
    from the longer sample's prefix; any in-place modification (autograd version bump) or a different tensor object
    invalidates it; outside a scope nothing is cached; inference tensors are never cached (they have no version)."""
    fps = pointnet2_utils.furthest_point_sample
    xyz = synthetic.uniform(3, 1500, 77).to(dev)
    h0 = fps.prefix_hits
    assert fps.reuse_prefix is False
    fps(xyz, 256)
    fps(xyz, 64)
    assert fps.prefix_hits == h0  # opt-in: no reuse outside a scope
    with pointnet2_utils.prefix_reuse():
        a256 = fps(xyz, 256)
        a64 = fps(xyz, 64)
        assert fps.prefix_hits == h0 + 1 and torch.equal(a64, a256[:, :64]) and a64.is_contiguous()
        a64.zero_()  # a caller scribbling over a returned result must not corrupt later hits
        a256.zero_()
        again = fps(xyz, 64)
        assert fps.prefix_hits == h0 + 2
        fps.reuse_prefix = False
        fresh = fps(xyz, 64)  # a fresh launch
        fps.reuse_prefix = True
        assert torch.equal(again, fresh)
        b300 = fps(xyz, 300)  # longer than anything cached -> computed, and consistent with the prefix property
        assert fps.prefix_hits == h0 + 2 and torch.equal(b300[:, :64], fresh)
        xyz[0, 0, 0] += 1.0  # in-place edit bumps the version: the cache must not be used
        c64 = fps(xyz, 64)
        assert fps.prefix_hits == h0 + 2
        assert np.array_equal(c64.cpu().numpy(), oracle.fps(xyz.cpu().numpy(), 64, True))
        clone = xyz.clone()
        fps(clone, 32)
        assert fps.prefix_hits == h0 + 2  # a different tensor object never hits
    assert fps.reuse_prefix is False
    with pointnet2_utils.prefix_reuse():
        with torch.inference_mode():
            t = synthetic.uniform(2, 700, 78).to(dev)
            i1 = fps(t, 128)
            i2 = fps(t, 32)  # no crash (inference tensors do not track versions) and no stale reuse
            assert fps.prefix_hits == h0 + 2 and torch.equal(i2, i1[:, :32])


def test_snowflake_query_knn(dev):
    """models/external/SnowFlakeNet_utils.py:284-289 (argsort of the distance matrix) vs the fused kernel."""
    from imgadapointr_b200.models.SnowFlakeNet_utils import query_knn

    xyz = synthetic.uniform(2, 700, 91).to(dev)
    q = xyz[:, :128].contiguous()
    for include_self in (True, False):
        got = query_knn(16, xyz, q, include_self=include_self)
        pad = 0 if include_self else 1
        D = oracle.torch_formulation.square_distance(q.double(), xyz.double())
        want = torch.argsort(D, dim=-1)[:, :, pad:16 + pad].int()
        assert got.dtype == torch.int32 and torch.equal(got, want)


def test_pytorch3d_knn_points_shim(dev):
    """K = 30 neighbourhoods as the reference's normal estimation asks for them (pctransformer.py:176)."""
    from imgadapointr_b200.utils.from_pytorch3d import knn_points
    xyz = synthetic.uniform(3, 700, 41)
    x = xyz.to(dev)
    out = knn_points(x, x, lengths1=torch.full((3,), 700), lengths2=torch.full((3,), 700), K=30, return_nn=True)
    oi, od = oracle.knn(30, xyz.numpy(), xyz.numpy())
    assert np.array_equal(out.idx.cpu().numpy(), oi) and np.array_equal(out.dists.cpu().numpy(), od)
    want_nn = np.take_along_axis(xyz.numpy()[:, None], oi[..., None], axis=2)
    assert out.knn.shape == (3, 700, 30, 3) and np.array_equal(out.knn.cpu().numpy(), want_nn)
    D = torch.cdist(xyz.double(), xyz.double()) ** 2
    assert torch.allclose(out.dists.cpu().double(), torch.sort(D, dim=-1).values[..., :30], atol=1e-6)
    assert knn_points(x, x[:, :100].contiguous(), K=1).knn is None
    with pytest.raises(NotImplementedError):
        knn_points(x, x, lengths1=torch.tensor([700, 650, 700]), K=3)
