"""ctypes binding of libiapt_b200.so (C ABI: include/iapt_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  PyTorch is used only
for device memory (outputs and workspaces come from its caching allocator) and for the current stream.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libiapt_b200.so")
ABI_VERSION = 2

_lib = None

_c_int = ctypes.c_int
_c_void_p = ctypes.c_void_p
_c_size_t = ctypes.c_size_t
_c_float = ctypes.c_float

# name -> (restype, argtypes); mirrors include/iapt_b200.h one to one (tests/test_abi.py checks the header against it)
SIGNATURES = {
    "iapt_abi_version": (_c_int, []),
    "iapt_error_string": (ctypes.c_char_p, [_c_int]),
    "iapt_launch_count": (ctypes.c_ulonglong, []),
    "iapt_profile_enable": (_c_int, [_c_int]),
    "iapt_profile_collect": (_c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_c_int)]),
    "iapt_chamfer_forward_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "iapt_chamfer_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                      _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_chamfer_backward_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "iapt_chamfer_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                       _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_chamfer_loss": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "iapt_chamfer_backward_fused": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                             _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                             _c_size_t, _c_void_p]),
    "iapt_chamfer_step_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "iapt_chamfer_step": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                                   _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                   _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_chamfer_step_stats": (_c_int, [_c_void_p, ctypes.POINTER(ctypes.c_ulonglong)]),
    "iapt_scale": (_c_int, [_c_void_p, _c_void_p, ctypes.c_longlong, _c_void_p, _c_void_p]),
    "iapt_chamfer_reduce": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_fscore_counts": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p]),
    "iapt_fps_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "iapt_fps": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_gather": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_gather_grad": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_gather_xyz": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_graph_feature": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                    _c_void_p]),
    "iapt_graph_feature_grad_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "iapt_graph_feature_grad": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                         _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_index_points": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, ctypes.c_longlong, _c_void_p, _c_void_p]),
    "iapt_index_points_grad": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, ctypes.c_longlong, _c_void_p,
                                        _c_void_p, _c_size_t, _c_void_p]),
    "iapt_graph_feature_cl": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                       _c_void_p]),
    "iapt_graph_feature_cl_grad": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                            _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_group_gather": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                   _c_void_p]),
    "iapt_group_gather_grad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int,
                                        _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "iapt_ball_query": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_int, _c_void_p, _c_void_p]),
    "iapt_emd_workspace_bytes": (_c_size_t, [_c_int, _c_int]),
    "iapt_emd_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_float, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                  _c_size_t, _c_void_p]),
    "iapt_emd_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_knn_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "iapt_square_distance": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "iapt_knn": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t,
                          _c_void_p]),
}


class IaptError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built: there is no CPU / eager fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IaptError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C imgadapointr_b200/csrc`). imgadapointr_b200 has no fallback path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.iapt_abi_version() != ABI_VERSION:
            raise IaptError(f"ABI mismatch: library {handle.iapt_abi_version()} != binding {ABI_VERSION}")
        _lib = handle
    return _lib


_ext = False  # not looked up yet


def torch_ext():
    """The compiled PyTorch binding over the same C ABI (csrc/torch_binding.cpp -> iapt_torch_ext*.so), or None when it
    is not built or ``IAPT_BINDING=ctypes`` / ``set_binding("ctypes")`` selects the torch-free ctypes table below.
    Both bindings call the same entry points of libiapt_b200.so; the compiled one validates, allocates and launches
    from C++ (TORCH_CHECK, at::empty, CUDAGuard, getCurrentCUDAStream) and costs a fraction of the host time."""
    global _ext
    if _ext is False:
        _ext = None
        if os.environ.get("IAPT_BINDING", "torch") != "ctypes":
            lib()  # the CUDA library must be present either way: no fallback
            try:
                from . import iapt_torch_ext as mod
            except ImportError:
                mod = None
            if mod is not None:
                if mod.abi_version() != ABI_VERSION:
                    raise IaptError(f"ABI mismatch: iapt_torch_ext {mod.abi_version()} != binding {ABI_VERSION}")
                _ext = mod
    return _ext


def set_binding(kind: str) -> None:
    """``"torch"`` (compiled extension; raises if it is not built) or ``"ctypes"`` - for tests and timing."""
    global _ext
    if kind == "ctypes":
        _ext = None
    elif kind == "torch":
        _ext = False
        os.environ.pop("IAPT_BINDING", None)
        if torch_ext() is None:
            raise IaptError("iapt_torch_ext is not built (python imgadapointr_b200/csrc/build_torch_ext.py)")
    else:
        raise ValueError(kind)


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().iapt_error_string(code)
        raise IaptError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(lib().iapt_launch_count())


def profile_enable(on: bool) -> None:
    check(lib().iapt_profile_enable(1 if on else 0), "iapt_profile_enable")


def profile_collect():
    """(total device ms of the dominant Chamfer kernel, number of launches) since profile_enable(True)."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int(0)
    check(lib().iapt_profile_collect(ctypes.byref(ms), ctypes.byref(n)), "iapt_profile_collect")
    return float(ms.value), int(n.value)


def require_cuda_f32(t: torch.Tensor, name: str, ndim: int) -> None:
    """The checks the reference leaves to data_ptr<float>() / CHECK_* macros, made explicit."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise IaptError(f"{name} must be a CUDA tensor (imgadapointr_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise IaptError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != ndim:
        raise IaptError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")


def ptr(t):
    return None if t is None else t.data_ptr()  # ctypes converts ints for c_void_p argtypes


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """`with on_device(dev):` - like torch.cuda.device(dev) but free when dev is already current (the common case;
    the torch context manager costs ~10 us per entry, comparable to a whole kernel launch)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            return self.ctx.__exit__(*a)
        return False
