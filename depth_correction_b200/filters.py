"""Masks over neighbourhood features (filters.py:85-113, 116-141, 184-254 of the reference)."""
import torch

from .depth_cloud import DepthCloud

from . import _lib as L

__all__ = ['filter_depth', 'filter_eigenvalue', 'filter_eigenvalue_ratio', 'filter_eigenvalue_ratios',
           'filter_eigenvalues', 'filter_shadow_points', 'filter_valid_neighbors', 'within_bounds']


def within_bounds(x, min=None, max=None, bounds=None, log_variable=None):
    """Mask of x being within bounds  min <= x <= max (inclusive; None / +-inf disable a side)."""
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    keep = torch.ones((x.numel(),), dtype=torch.bool, device=x.device)
    if bounds:
        assert min is None and max is None
        min, max = bounds
    if min is not None and min > -float('inf'):
        keep = keep & (x.flatten() >= min)
    if max is not None and max < float('inf'):
        keep = keep & (x.flatten() <= max)
    if log_variable is not None:
        print('%.3f = %i / %i points kept (%.3g <= %s <= %.3g).'
              % (keep.double().mean(), keep.sum(), keep.numel(),
                 min if min is not None else float('nan'), log_variable, max if max is not None else float('nan')))
    return keep


def filter_depth(cloud, min=None, max=None, only_mask=False, log=False):
    """Keep points with depth in bounds (filters.py:116-141)."""
    assert isinstance(cloud, DepthCloud)
    keep = within_bounds(cloud.depth, min=min, max=max, log_variable='depth' if log else None)
    return keep if only_mask else cloud[keep]


def filter_valid_neighbors(cloud, min=None, only_mask=False, log=False):
    """Keep points with enough valid neighbors."""
    assert isinstance(cloud, DepthCloud)
    keep = within_bounds(cloud.num_valid_neighbors(), min=min, log_variable='valid neighbors' if log else None)
    return keep if only_mask else cloud[keep]


def filter_eigenvalue(cloud, eigenvalue=0, min=None, max=None, only_mask=False, log=False):
    with torch.no_grad():
        keep = within_bounds(cloud.eigvals[:, eigenvalue], min=min, max=max,
                             log_variable='eigenvalue %i' % eigenvalue if log else None)
    return keep if only_mask else cloud[keep]


def filter_eigenvalues(cloud, bounds, only_mask=False, log=False):
    mask = None
    if bounds:
        for eig, min, max in bounds:
            eig_mask = filter_eigenvalue(cloud, eig, min=min, max=max, only_mask=True, log=log)
            mask = eig_mask if mask is None else mask & eig_mask
    else:
        mask = torch.ones((cloud.size(),), dtype=torch.bool, device=cloud.device())
    return mask if only_mask else cloud[mask]


def filter_eigenvalue_ratio(cloud, eigenvalues=(0, 1), min=None, max=None, only_mask=False, log=False):
    assert cloud.eigvals is not None
    assert len(eigenvalues) == 2
    assert all(0 <= i <= 2 for i in eigenvalues)
    i, j = eigenvalues
    with torch.no_grad():
        ratio = cloud.eigvals[:, i] / cloud.eigvals[:, j]
        keep = within_bounds(ratio, min=min, max=max,
                             log_variable='eigenvalue %i / eigenvalue %i' % tuple(eigenvalues) if log else None)
    return keep if only_mask else cloud[keep]


def filter_eigenvalue_ratios(cloud, bounds, only_mask=False, log=False):
    mask = None
    if bounds:
        for i, j, min, max in bounds:
            eig_mask = filter_eigenvalue_ratio(cloud, (i, j), min=min, max=max, only_mask=True, log=log)
            mask = eig_mask if mask is None else mask & eig_mask
    else:
        mask = torch.ones((cloud.size(),), dtype=torch.bool, device=cloud.device())
    return mask if only_mask else cloud[mask]


def filter_shadow_points(cloud, angle_bounds, only_mask=False, log=False):
    """Remove shadow points (filters.py:257-309): bound the minimum and maximum angle, at the point, between the
    ray back to the viewpoint and the directions to its neighbours among neighbouring BEAMS (dir_neighbors).

    One kernel (dc_shadow_mask) instead of the reference's [N,K,3] temporaries.  `only_mask=True` returns the mask
    (the reference returns the flag itself there, filters.py:303-304 -- a bug no caller relies on)."""
    import math
    assert cloud.vps is not None
    assert cloud.dir_neighbors is not None
    lo, hi = angle_bounds[0], angle_bounds[1]
    if lo is None or not (lo >= 0.0):
        lo = 0.0
    if hi is None or not (hi <= math.pi):
        hi = math.pi
    x = cloud.get_points().detach().contiguous()
    if not x.is_cuda:
        raise RuntimeError('filter_shadow_points needs a CUDA cloud; there is no CPU fallback')
    n = x.shape[0]
    vps = cloud.vps.detach().to(x.dtype).expand(n, 3).contiguous()
    nb = cloud.dir_neighbors.contiguous()
    w = None if cloud.dir_neighbor_weights is None else cloud.dir_neighbor_weights.to(torch.float32).contiguous()
    keep = torch.empty(n, dtype=torch.uint8, device=x.device)
    L.call('dc_shadow_mask', L.ptr(x), L.ptr(vps), L.dtype_code(x.dtype), L.ptr(nb), L.ptr(w), n, nb.shape[1], float(lo), float(hi),
           L.ptr(keep), None, None, L.stream())
    mask = keep.bool()
    if log:
        print('%.3f = %i / %i points kept (shadow points removed).' % (mask.double().mean(), mask.sum(), mask.numel()))
    if only_mask:
        return mask
    return cloud[mask]
