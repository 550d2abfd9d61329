"""CPU: the C-ABI library loads, exports every symbol include/iapt_b200.h declares, the ctypes table mirrors the
header, and argument validation / workspace arithmetic work without a GPU (no compute call is made)."""
import ctypes
import os
import re
import subprocess

import pytest

from imgadapointr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "iapt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"^\s*([A-Za-z_][\w \*]*?)\b(iapt_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(2)] = n
    return decls


def test_library_exports_every_declared_symbol():
    decls = _declared()
    assert len(decls) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in iapt_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(iapt_\w+)\b", out))
    assert exported == set(decls), exported ^ set(decls)


def test_ctypes_table_mirrors_header():
    decls = _declared()
    assert set(decls) == set(_lib.SIGNATURES)
    for name, n in decls.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_no_torch_types_cross_the_boundary():
    src = open(HEADER).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", src, flags=re.S).lower()
    assert 'extern "C"' in src
    # the library itself must not link against torch / python
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libc10" not in out and "libpython" not in out


def test_version_errors_and_workspace_arithmetic():
    L = _lib.lib()
    assert L.iapt_abi_version() == 2 == _lib.ABI_VERSION
    assert b"invalid" in L.iapt_error_string(-1)
    assert b"workspace" in L.iapt_error_string(-2)
    assert L.iapt_error_string(0) == b"success"
    assert L.iapt_chamfer_forward_workspace_bytes(0, 5, 5) == 0
    for (B, N, M) in [(1, 1, 1), (3, 700, 16384), (48, 16384, 16384), (256, 16384, 16384)]:
        fw = L.iapt_chamfer_forward_workspace_bytes(B, N, M)   # keys (8 B / point) + counters
        bw = L.iapt_chamfer_backward_workspace_bytes(B, N, M)  # counts + offsets + CSR (12 B / point) + counters
        sw = L.iapt_chamfer_step_workspace_bytes(B, N, M)      # everything + per-32-point partial sums
        pts = B * (N + M)
        assert 8 * pts <= fw <= 8 * pts + 16 * B + 4096
        assert 12 * pts <= bw <= 12 * pts + 24 * B + 4096
        assert fw + bw - 4096 <= sw <= 21 * pts + 64 * B + 8192
    # the one-launch step validates before any CUDA call
    assert L.iapt_chamfer_step(None, None, 1, 4, 4, 1, 0, 0.0, None, None, None, None, None, None, None, None, None, None,
                               None, None, 0, None) == -1
    assert L.iapt_chamfer_step(None, None, 0, 4, 4, 1, 0, 0.0, None, None, None, None, None, None, None, None, None, None,
                               None, None, 0, None) == 0  # empty batch
    assert L.iapt_chamfer_step(None, None, 1, 4, 4, 5, 0, 0.0, None, None, None, None, None, None, None, None, None, None,
                               None, None, 0, None) == -1  # unknown loss kind
    assert L.iapt_chamfer_backward(None, None, None, None, None, None, 1, 0, 4, None, None, None, 0, None) == 0
    assert L.iapt_scale(None, None, 0, None, None) == 0 and L.iapt_scale(None, None, 5, None, None) == -1
    assert L.iapt_fps_workspace_bytes(4, 2048, 512) == 0
    assert L.iapt_fps_workspace_bytes(2, 65536, 512) == 2 * 65536 * 4
    # argument validation happens before any CUDA call
    assert L.iapt_chamfer_forward(None, None, -1, 1, 1, None, None, None, None, None, 0, None) == -1
    assert L.iapt_chamfer_forward(None, None, 1, 4, 4, None, None, None, None, None, 0, None) == -1
    assert L.iapt_knn(None, None, 1, 10, 10, 65, None, None, None, 0, None) == -3
    assert L.iapt_knn(None, None, 1, 10, 10, 11, None, None, None, 0, None) == -1
    assert L.iapt_knn_workspace_bytes(4, 2048, 2048, 16) == 0 and L.iapt_knn_workspace_bytes(48, 512, 2048, 32) == 0
    # the culled path (dense queries, from 640 warps of them at 2048 references): curve-ordered copies of both clouds (16 B / point) + 32 B per tile of 32
    assert L.iapt_knn_workspace_bytes(48, 2048, 2048, 16) == 48 * (2048 * 16 + 64 * 32 + 2048 * 16) + 256
    assert L.iapt_knn_workspace_bytes(48, 16384, 64, 32) > 0  # few queries, many references: per-slice lists
    assert L.iapt_fps(None, 1, 0, 4, 0, None, None, 0, None) == -1
    assert L.iapt_chamfer_forward(None, None, 0, 4, 4, None, None, None, None, None, 0, None) == 0  # empty batch


def test_product_raises_without_gpu_or_library(monkeypatch, tmp_path):
    import torch

    from imgadapointr_b200 import chamfer
    from imgadapointr_b200.models.Transformer_utils import knn_point
    from imgadapointr_b200.pointnet2_ops import pointnet2_utils

    x = torch.zeros((1, 4, 3))
    for call in (lambda: chamfer.forward(x, x), lambda: knn_point(2, x, x),
                 lambda: pointnet2_utils.furthest_point_sample(x, 2)):
        with pytest.raises(_lib.IaptError):
            call()  # CPU tensors: there is no CPU path
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(_lib.IaptError):
        _lib.lib()
