"""CPU oracle for the geometry hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker.  ``imgadapointr_b200`` never imports it (tests/test_no_oracle_in_product.py
enforces that) and raises if its CUDA library is missing.

Two layers:

* ``oracle.c`` (built to ``liboracle.so`` by ``oracle/Makefile``): exact-rounding C restatement of the reference
  algorithms (see the header of that file for the reference file:line each function follows and for the
  pinning status -- chamfer is pinned against the reference's own compiled kernel, FPS/gather/kNN/F-Score are
  "parity unpinned").
* ``torch_formulation``: the reference's *PyTorch CPU formulation* named by BASELINE.json's north_star
  (``torch.cdist``+``min`` Chamfer, ``square_distance``+``topk`` kNN verbatim from
  models/Transformer_utils.py:19-51, a Python-loop FPS).  This is what ``bench.py`` times as the CPU baseline.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc, a second or two)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def chamfer_forward(xyz1, xyz2):
    """chamfer.forward (extensions/chamfer_dist/chamfer.cu:147-171) -> dist1, dist2, idx1, idx2."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    d1 = np.zeros((B, N), np.float32)
    d2 = np.zeros((B, M), np.float32)
    i1 = np.zeros((B, N), np.int32)
    i2 = np.zeros((B, M), np.int32)
    lib().oracle_chamfer_forward(_p(xyz1), _p(xyz2), B, N, M, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    """chamfer.backward (chamfer.cu:203-229) -> grad_xyz1, grad_xyz2."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    idx1, idx2 = _i32(idx1), _i32(idx2)
    g1, g2 = _f32(grad_dist1), _f32(grad_dist2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    gx1 = np.zeros_like(xyz1)
    gx2 = np.zeros_like(xyz2)
    lib().oracle_chamfer_backward(_p(xyz1), _p(xyz2), _p(idx1), _p(idx2), _p(g1), _p(g2), B, N, M, _p(gx1), _p(gx2))
    return gx1, gx2


def fps(xyz, npoint: int, skip_near_origin: bool = True):
    """pointnet2_utils.furthest_point_sample (restated; parity unpinned) -> int32 [B, npoint]."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    lib().oracle_fps(_p(xyz), B, N, int(npoint), int(bool(skip_near_origin)), _p(idx))
    return idx


def fps_ex(xyz, npoint: int, skip_near_origin: bool = True, tie: str = "lowest"):
    """FPS with the tie order as a parameter and tie statistics (see oracle.c: oracle_fps_ex).  ``tie`` = "lowest" (ours)
    or "upstream" (smallest (k mod block, k), the published pointnet2_ops kernel's thread-strided scan + tree).
    Returns (idx int32 [B,npoint], stats int64 [B,2] = rounds with a tie between DIFFERENT points / between exact
    duplicates only)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, npoint), np.int32)
    stats = np.zeros((B, 2), np.int64)
    lib().oracle_fps_ex(_p(xyz), B, N, int(npoint), int(bool(skip_near_origin)), {"lowest": 0, "upstream": 1}[tie],
                        _p(idx), _p(stats))
    return idx, stats


def gather(features, idx):
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    m = idx.shape[1]
    out = np.zeros((B, C, m), np.float32)
    lib().oracle_gather(_p(features), _p(idx), B, C, N, m, _p(out))
    return out


def gather_grad(grad_out, idx, N: int):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = np.zeros((B, C, N), np.float32)
    lib().oracle_gather_grad(_p(grad_out), _p(idx), B, C, N, m, _p(out))
    return out


def knn(k: int, xyz, new_xyz, use_f64: bool = False):
    """knn_point(nsample, xyz, new_xyz) -> (int64 idx [B,S,k] ascending by (distance, index), float32 dist)."""
    ref, qry = _f32(xyz), _f32(new_xyz)
    B, R, _ = ref.shape
    Q = qry.shape[1]
    idx = np.zeros((B, Q, k), np.int64)
    dist = np.zeros((B, Q, k), np.float32)
    lib().oracle_knn(_p(ref), _p(qry), B, R, Q, int(k), int(bool(use_f64)), _p(idx), _p(dist))
    return idx, dist


def fscore_counts(pred, gt, th: float = 0.01):
    """Float64 NN counts behind utils/metrics.py:67-90 -> (counts [B,2] int64, boundary [B] int64)."""
    pred, gt = _f32(pred), _f32(gt)
    B, N, _ = pred.shape
    M = gt.shape[1]
    counts = np.zeros((B, 2), np.int64)
    boundary = np.zeros((B,), np.int64)
    lib().oracle_fscore_counts(_p(pred), _p(gt), B, N, M, ctypes.c_double(th), _p(counts), _p(boundary))
    return counts, boundary


def fscore(pred, gt, th: float = 0.01):
    """Metrics._get_f_score (utils/metrics.py:67-90): mean over the batch of per-cloud F."""
    counts, _ = fscore_counts(pred, gt, th)
    N, M = np.asarray(pred).shape[1], np.asarray(gt).shape[1]
    fs = []
    for c in counts:
        precision, recall = c[0] / float(N), c[1] / float(M)
        fs.append(2 * recall * precision / (recall + precision) if recall + precision else 0.0)
    return float(sum(fs) / len(fs))


from . import torch_formulation  # noqa: E402,F401


def emd_forward(xyz1, xyz2, eps: float, iters: int):
    """Auction EMD of extensions/emd (emd_cuda.cu:95-281) with explicit tie rules -> (dist [B,n] f32, assignment [B,n] i32)."""
    a, b = _f32(xyz1), _f32(xyz2)
    B, n, _ = a.shape
    assert b.shape == a.shape, "the reference requires clouds of the same size (emd_cuda.cu:238-241)"
    dist = np.empty((B, n), dtype=np.float32)
    asg = np.empty((B, n), dtype=np.int32)
    lib().oracle_emd_forward(_p(a), _p(b), B, n, ctypes.c_float(eps), int(iters), _p(dist), _p(asg))
    return dist, asg
