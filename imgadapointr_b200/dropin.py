"""Register the B200 implementations under the module names the reference imports.

After ``imgadapointr_b200.dropin.install()``:
    import chamfer                                  -> imgadapointr_b200.chamfer
    from pointnet2_ops import pointnet2_utils       -> imgadapointr_b200.pointnet2_ops.pointnet2_utils
    from extensions.chamfer_dist import ...         -> imgadapointr_b200.extensions.chamfer_dist   (optional)
    from extensions.emd import emd_module           -> imgadapointr_b200.extensions.emd.emd_module (optional)
and ``patch_reference_modules()`` swaps ``knn_point`` / ``fps`` / ``Metrics`` inside already-imported reference modules
(models.Transformer_utils, models.base_blocks, utils.misc, utils.metrics, ...), which is needed because the reference
star-imports ``knn_point`` into every model module (models/base_blocks.py:9, models/external/AdaPoinTr.py:14).
"""
from __future__ import annotations

import sys


def install(override_extensions: bool = True) -> None:
    from . import chamfer, pointnet2_ops
    from .extensions import chamfer_dist, emd
    from .pointnet2_ops import pointnet2_utils

    sys.modules["chamfer"] = chamfer
    sys.modules["pointnet2_ops"] = pointnet2_ops
    sys.modules["pointnet2_ops.pointnet2_utils"] = pointnet2_utils
    if override_extensions:
        import importlib.util

        from . import extensions as our_extensions

        sys.modules["extensions.chamfer_dist"] = chamfer_dist
        sys.modules["extensions.emd"] = emd  # utils/metrics.py:13 imports it at module load; install.sh never builds it
        sys.modules["extensions.emd.emd_module"] = emd.emd_module
        ext = sys.modules.get("extensions")
        if ext is None:
            try:
                found = importlib.util.find_spec("extensions") is not None
            except (ImportError, ValueError):
                found = False
            if not found:  # reference tree not on sys.path: our package stands in for the parent too
                sys.modules["extensions"] = ext = our_extensions
        if ext is not None:
            setattr(ext, "chamfer_dist", chamfer_dist)
            setattr(ext, "emd", emd)


def patch_reference_modules() -> list:
    """Replace knn_point / fps / Metrics in every loaded reference module that holds them. Returns what was patched."""
    from .models import Transformer_utils as tu
    from .utils import metrics as m
    from .utils import misc

    from .models import base_blocks as bb

    patched = []
    ref_bb = sys.modules.get("models.base_blocks")
    grouper = getattr(ref_bb, "DGCNN_Grouper", None) if ref_bb is not None else None
    if grouper is not None:  # geometry methods of the reference's grouper (base_blocks.py:547-588)
        grouper.fps_downsample = staticmethod(bb.fps_downsample)
        grouper.get_graph_feature = lambda self, coor_q, x_q, coor_k, x_k: bb.get_graph_feature(self.k, coor_q, x_q,
                                                                                                coor_k, x_k)
        patched += ["models.base_blocks.DGCNN_Grouper.fps_downsample", "models.base_blocks.DGCNN_Grouper.get_graph_feature"]
    for name, mod in list(sys.modules.items()):
        if mod is None or name.startswith("imgadapointr_b200"):
            continue
        root = name.split(".")[0]
        if root not in ("models", "utils", "tools", "extensions", "datasets"):
            continue
        from .models import SnowFlakeNet_utils as sf

        for attr, new in (("knn_point", tu.knn_point), ("index_points", tu.index_points), ("fps", misc.fps),
                          ("Metrics", m.Metrics), ("query_knn", sf.query_knn)):
            old = getattr(mod, attr, None)
            if old is not None and old is not new and callable(old):
                if attr == "fps" and not name.endswith("misc") and getattr(old, "__module__", "") != "utils.misc":
                    continue
                setattr(mod, attr, new)
                patched.append(f"{name}.{attr}")
    ref_tu = sys.modules.get("models.Transformer_utils")
    dga = getattr(ref_tu, "DynamicGraphAttention", None) if ref_tu is not None else None
    if dga is not None:  # gather / sub / cat of the geometry-aware attention (Transformer_utils.py:796-860)
        dga.forward = tu.dynamic_graph_attention_forward
        patched.append("models.Transformer_utils.DynamicGraphAttention.forward")
    ref_p3d = sys.modules.get("utils.from_pytorch3d")
    if ref_p3d is not None:  # the vendored copy calls knn_points without defining it (utils/from_pytorch3d.py:76)
        from .utils import from_pytorch3d as p3d

        ref_p3d.knn_points = p3d.knn_points
        patched.append("utils.from_pytorch3d.knn_points")
    return patched
