"""Shim for the third-party ``pointnet2_ops`` package the reference imports (utils/misc.py:10,
models/Transformer_utils.py:14) but does not vendor.  Only the hot-path ops are implemented."""
from . import pointnet2_utils  # noqa: F401
