"""imgadapointr_b200 - B200-native (sm_100a) point-cloud geometry hot path for ImgAdaPoinTr.

Host-side mirror of the reference's interfaces for this path (same names, arguments and error behaviour), each a
thin layer over the C ABI in include/iapt_b200.h (libiapt_b200.so):

    imgadapointr_b200.chamfer                      <- pybind module `chamfer`
    imgadapointr_b200.extensions.chamfer_dist      <- extensions/chamfer_dist
    imgadapointr_b200.pointnet2_ops.pointnet2_utils<- pointnet2_ops.pointnet2_utils (FPS, gather)
    imgadapointr_b200.utils.misc.fps               <- utils/misc.py:fps
    imgadapointr_b200.models.Transformer_utils     <- knn_point / square_distance / index_points
    imgadapointr_b200.utils.metrics.Metrics        <- utils/metrics.py
    imgadapointr_b200.utils.dist_utils             <- utils/dist_utils.py (+ batch sharding helpers)

``imgadapointr_b200.dropin.install()`` registers these under the reference's own module names so the reference's
model / loss / metric code imports them unchanged.  There is no CPU or eager fallback anywhere:
without the built CUDA library every op raises.
"""
from . import _lib  # noqa: F401
from ._lib import IaptError, launch_count  # noqa: F401

__version__ = "0.2.0"
