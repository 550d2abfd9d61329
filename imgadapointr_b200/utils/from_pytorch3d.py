"""`knn_points` with pytorch3d's signature on the B200 kNN kernel.

The reference estimates normals from K = 30 neighbourhoods through pytorch3d
(models/internal/pctransformer.py:176,239,294 -> pytorch3d.ops.points_normals -> knn_points) and carries a vendored
copy of that module whose `knn_points` name is unresolved (utils/from_pytorch3d.py:76).  This is that one function:
squared distances ascending, int64 indices, optional gathered neighbours.  Only what the call sites use is supported -
full-length clouds (lengths == P), 3-D points, norm 2.
"""
from collections import namedtuple

import torch

from ..models.Transformer_utils import index_points, knn_point

_KNN = namedtuple("KNN", "dists idx knn")


def knn_points(p1, p2, lengths1=None, lengths2=None, norm: int = 2, K: int = 1, version: int = -1,
               return_nn: bool = False, return_sorted: bool = True):
    """p1 [N,P1,3] queries, p2 [N,P2,3] references -> KNN(dists [N,P1,K] squared, idx [N,P1,K] int64, knn [N,P1,K,3] or
    None).  Neighbours are ascending by (distance, index), i.e. always 'sorted'; lowest index on exact ties."""
    if norm != 2:
        raise NotImplementedError("knn_points: only the squared L2 distance of the reference's call sites")
    for lengths, p in ((lengths1, p1), (lengths2, p2)):
        if lengths is not None and not bool((torch.as_tensor(lengths) == p.shape[1]).all()):
            raise NotImplementedError("knn_points: ragged clouds (lengths < P) are not used by the reference")
    idx, dists = knn_point(int(K), p2, p1, return_dist=True)
    nn = index_points(p2, idx) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=nn)
