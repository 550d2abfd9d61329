"""Generates the committed golden fixtures in tests/golden/ by running the REFERENCE'S OWN Python code in the build
container (it cannot travel to the GPU box, the fixtures can).  Run once: `python tests/golden/make_golden.py`.

What runs from /root/reference (imported, never copied):
  * models/Transformer_utils.py: knn_point / square_distance / index_points (CPU torch).  Its unrelated imports
    (timm, pointnet2_ops) are absent in this image and are stubbed with empty modules - they are not called.
  * extensions/chamfer_dist/__init__.py: ChamferFunction + the four loss modules.  Its `import chamfer` (the CUDA
    pybind module, needs a GPU) is satisfied by a CPU stand-in backed by oracle.chamfer_forward/backward, so the
    fixtures pin the reference's REDUCTIONS, autograd wiring and ignore_zeros filter on top of the oracle's search.
Inputs: the reference's demo/samples/*/00.pcd partial clouds (ASCII PCD, FIELDS x y z), zero padded to 2048 points the
way datasets/data_transforms.py:55-67 does, plus seeded synthetic pairs.
"""
import glob
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from imgadapointr_b200.utils import synthetic  # noqa: E402


def load_pcd(path):
    with open(path) as f:
        lines = f.read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith("DATA")) + 1
    return np.array([[float(v) for v in l.split()[:3]] for l in lines[start:] if l.strip()], dtype=np.float32)


def import_reference():
    # stand-in for the CUDA pybind module `chamfer` (chamfer_cuda.cpp:36-39), CPU, oracle-backed
    ch = types.ModuleType("chamfer")

    def forward(xyz1, xyz2):
        d1, d2, i1, i2 = oracle.chamfer_forward(xyz1.detach().numpy(), xyz2.detach().numpy())
        return [torch.from_numpy(d1), torch.from_numpy(d2), torch.from_numpy(i1), torch.from_numpy(i2)]

    def backward(xyz1, xyz2, idx1, idx2, g1, g2):
        a, b = oracle.chamfer_backward(xyz1.detach().numpy(), xyz2.detach().numpy(), idx1.numpy(), idx2.numpy(),
                                       g1.contiguous().numpy(), g2.contiguous().numpy())
        return [torch.from_numpy(a), torch.from_numpy(b)]

    ch.forward, ch.backward = forward, backward
    sys.modules["chamfer"] = ch
    timm = types.ModuleType("timm")
    timm_models = types.ModuleType("timm.models")
    timm_layers = types.ModuleType("timm.models.layers")
    timm_layers.DropPath = torch.nn.Identity
    timm_layers.trunc_normal_ = torch.nn.init.trunc_normal_
    sys.modules.update({"timm": timm, "timm.models": timm_models, "timm.models.layers": timm_layers})
    p2 = types.ModuleType("pointnet2_ops")
    p2u = types.ModuleType("pointnet2_ops.pointnet2_utils")
    p2.pointnet2_utils = p2u
    sys.modules.update({"pointnet2_ops": p2, "pointnet2_ops.pointnet2_utils": p2u})
    sys.path.insert(0, REF)
    import importlib.util

    def load(name, rel):  # load ONE reference file as a module (its package __init__ pulls in the whole model zoo)
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ref_cd = load("ref_chamfer_dist", "extensions/chamfer_dist/__init__.py")
    ref_tu = load("ref_transformer_utils", "models/Transformer_utils.py")
    # models/base_blocks.py (DGCNN_Grouper geometry): `from models.Transformer_utils import *`, `from utils import misc`
    # and `pointnet2_utils` are satisfied by the module loaded above, an empty stand-in and oracle-backed FPS/gather.
    p2u.furthest_point_sample = lambda xyz, m: torch.from_numpy(oracle.fps(xyz.detach().numpy(), int(m), True))
    p2u.gather_operation = lambda feat, idx: torch.gather(feat, 2, idx.long()[:, None, :].expand(-1, feat.size(1), -1))
    models_pkg = types.ModuleType("models")
    models_pkg.__path__ = []
    utils_pkg = types.ModuleType("utils")
    utils_pkg.__path__ = []
    utils_pkg.misc = types.ModuleType("utils.misc")
    sys.modules.update({"models": models_pkg, "models.Transformer_utils": ref_tu, "utils": utils_pkg,
                        "utils.misc": utils_pkg.misc, "extensions": types.ModuleType("extensions"),
                        "extensions.chamfer_dist": ref_cd})
    ref_bb = load("ref_base_blocks", "models/base_blocks.py")
    return ref_cd, ref_tu, ref_bb


def main():
    ref_cd, ref_tu, ref_bb = import_reference()
    torch.manual_seed(0)
    # ---- demo samples: FPS (oracle, unpinned), kNN k=16 from the reference's own knn_point -------------------
    for path in sorted(glob.glob(os.path.join(REF, "demo/samples/*/00.pcd"))):
        synset = os.path.basename(os.path.dirname(path))
        pts = load_pcd(path)
        xyz = np.zeros((2048, 3), np.float32)
        xyz[: len(pts)] = pts[:2048]
        t = torch.from_numpy(xyz)[None]
        ref_idx = ref_tu.knn_point(16, t, t)[0]                     # reference, order unspecified
        D = ref_tu.square_distance(t.double(), t.double())[0]       # float64 for tie classification
        srt = torch.sort(D, dim=-1).values
        gap = (srt[:, 16] - srt[:, 15])
        # rows whose neighbour SET is unambiguous even under the reference's FP32 expansion formula, whose absolute
        # error is ~4 * 2^-24 * (|a|^2 + |b|^2) <= 1.2e-7 for coordinates in [-0.5, 0.5]
        untied = gap > 2e-7
        oknn, _ = oracle.knn(16, xyz[None], xyz[None])
        ofps = oracle.fps(xyz[None], 512, True)[0]
        np.savez_compressed(
            os.path.join(HERE, f"sample_{synset}.npz"), xyz=xyz, n_real=np.int32(len(pts)),
            fps512=ofps.astype(np.int32), knn16=oknn[0].astype(np.int16),
            ref_knn16_sorted=torch.sort(ref_idx, dim=-1).values.numpy().astype(np.int16),
            ref_knn16_untied=untied.numpy())
        print(synset, len(pts), "untied rows", int(untied.sum()))
    # ---- loss modules: the reference's reductions + autograd over the oracle's search ----------------------------
    pred, gt = synthetic.pcn_pair(2, 512, seed=11)
    gt = gt[:, :384].contiguous()
    out = {"pred": pred.numpy(), "gt": gt.numpy()}
    for name in ("ChamferDistanceL1", "ChamferDistanceL2", "ChamferDistanceL2_split", "ChamferDistanceL1_PM"):
        p = pred.clone().requires_grad_(True)
        g = gt.clone().requires_grad_(True)
        val = getattr(ref_cd, name)()(p, g)
        val = torch.stack(list(val)) if isinstance(val, tuple) else val
        val.sum().backward()
        out[name + "_value"] = val.detach().numpy()
        out[name + "_grad_pred"] = p.grad.numpy()
        out[name + "_grad_gt"] = g.grad.numpy() if g.grad is not None else np.zeros_like(gt.numpy())
    # ignore_zeros (batch 1 only): a zero-padded partial against its gt
    gt1 = synthetic.pcn_gt(1, 1024, 12)
    part = synthetic.pcn_partial(gt1, 512, 13)
    part[0, 3] = torch.tensor([0.25, -0.5, 0.25])
    out["iz_partial"], out["iz_gt"] = part.numpy(), gt1.numpy()
    out["iz_L1"] = ref_cd.ChamferDistanceL1(ignore_zeros=True)(part, gt1).numpy()
    out["iz_L2"] = ref_cd.ChamferDistanceL2(ignore_zeros=True)(part, gt1).numpy()
    out["iz_L1_batch2"] = ref_cd.ChamferDistanceL1(ignore_zeros=True)(torch.cat([part, part]), torch.cat([gt1, gt1])).numpy()
    # index_points on a kNN result
    pts = torch.randn(2, 384, 5)
    idx = ref_tu.knn_point(4, gt, gt[:, :50].contiguous())
    out["ip_points"], out["ip_idx"], out["ip_out"] = pts.numpy(), idx.numpy(), ref_tu.index_points(pts, idx).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_losses.npz"), **out)
    # ---- DGCNN_Grouper geometry (base_blocks.py:547-588): the reference's own methods on CPU --------------------------
    g = torch.Generator().manual_seed(21)
    B, C, Nk, Nq, k = 2, 8, 300, 100, 16
    coor_k = torch.rand((B, 3, Nk), generator=g) - 0.5
    coor_q = coor_k[:, :, :Nq].contiguous()
    x_k = torch.randn((B, C, Nk), generator=g).requires_grad_(True)
    x_q = torch.randn((B, C, Nq), generator=g).requires_grad_(True)
    self_like = types.SimpleNamespace(k=k)
    feat = ref_bb.DGCNN_Grouper.get_graph_feature(self_like, coor_q, x_q, coor_k, x_k)
    go = torch.randn(feat.shape, generator=g)
    feat.backward(go)
    with torch.no_grad():
        idx = ref_tu.knn_point(k, coor_k.transpose(-1, -2).contiguous(), coor_q.transpose(-1, -2).contiguous())
    gg = {"coor_k": coor_k.numpy(), "coor_q": coor_q.numpy(), "x_k": x_k.detach().numpy(), "x_q": x_q.detach().numpy(),
          "ref_idx": idx.numpy(), "feature": feat.detach().numpy(), "grad_out": go.numpy(),
          "grad_x_k": x_k.grad.numpy(), "grad_x_q": x_q.grad.numpy()}
    coor = torch.rand((B, 3, 200), generator=g) - 0.5
    xf = torch.randn((B, 5, 200), generator=g)
    nc, nx = ref_bb.DGCNN_Grouper.fps_downsample(coor, xf, 40)
    gg.update({"fd_coor": coor.numpy(), "fd_x": xf.numpy(), "fd_new_coor": nc.numpy(), "fd_new_x": nx.numpy()})
    np.savez_compressed(os.path.join(HERE, "reference_grouper.npz"), **gg)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
