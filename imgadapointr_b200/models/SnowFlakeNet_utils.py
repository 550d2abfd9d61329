"""Drop-in for the kNN helper of the reference's ``models/external/SnowFlakeNet_utils.py``: ``query_knn`` (:284-289),
which argsorts a materialised [B,S,N] distance matrix.  Same result from the fused kNN kernel (already ascending)."""
from .Transformer_utils import knn_point


def query_knn(nsample, xyz, new_xyz, include_self=True):
    """Find the k nearest neighbours of new_xyz [B,S,3] in xyz [B,N,3] -> int32 [B,S,nsample], ascending by distance;
    ``include_self=False`` drops the nearest one (the query itself when new_xyz is a subset of xyz)."""
    pad = 0 if include_self else 1
    idx = knn_point(nsample + pad, xyz, new_xyz)
    return idx[:, :, pad:nsample + pad].int()
