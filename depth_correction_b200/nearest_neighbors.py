"""Drop-in for depth_correction.nearest_neighbors (nearest_neighbors.py:13-80), on the GPU.

Same signature and return convention: `(dist, ind)`, `ind` int64 [N,K] with -1 for missing
neighbours; kNN mode returns float64 distances (inf for missing), radius mode returns `dist=None`
and index-sorted rows padded to the longest row.  The search itself is kernel 1 (libdcb200.so);
the returned index tensor carries the internal sorted-space graph (`ind._dc_graph`) so that
DepthCloud / the fused step can reuse it without re-importing the padded matrix.
"""
import torch

from .graph import search

__all__ = ['ball_angle_to_distance', 'nearest_neighbors']


def ball_angle_to_distance(angle, radius=1.0):
    """Chord length of a ball of given angular radius (nearest_neighbors.py:13-19)."""
    assert isinstance(angle, torch.Tensor)
    angle = torch.clamp(angle, 0., torch.pi)
    dist = torch.sqrt(2. * (1. - torch.cos(angle)))
    if isinstance(radius, float) or radius != 1.0:
        dist = radius * dist
    return dist


def nearest_neighbors(points, query, k=None, r=None, n_jobs=-1):
    """Find nearest neighbors of query in points.

    :param points: Reference points in rows (CUDA tensor).
    :param query: Query points in rows.
    :param k: Number of neighbors.
    :param r: Radius in which to find neighbors (with k: strict upper bound on the distance).
    :return: Tuple with distances and indices. Distances may be None. Missing neighbors are indicated by -1.
    """
    assert isinstance(points, torch.Tensor)
    assert isinstance(query, torch.Tensor)
    assert k or r
    if not points.is_cuda:
        raise RuntimeError('depth_correction_b200.nearest_neighbors needs CUDA tensors; there is no CPU fallback')
    points = points.reshape([-1, points.shape[-1]])
    query = query.reshape([-1, points.shape[-1]])
    same = (query.data_ptr() == points.data_ptr() and query.shape == points.shape
            and query.stride() == points.stride())
    graph = search(points, None if same else query, k=k, r=r)
    ind = graph.neighbors()
    dist = graph.distances()
    if same:
        ind._dc_graph = graph
    return dist, ind
