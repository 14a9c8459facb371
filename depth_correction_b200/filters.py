"""Masks over neighbourhood features (filters.py:85-113, 116-141, 184-254 of the reference).

On CUDA clouds every bound of a call -- and, through `feature_mask`, every bound of a whole configuration (eigenvalue
bounds, eigenvalue-ratio bounds, minimum number of valid neighbours, a starting mask) -- is evaluated by ONE launch of
dc_feature_mask instead of one comparison and one AND kernel per bound.  Host tensors (configuration glue, CPU tests of
the host logic) take the plain torch comparisons.
"""
import ctypes

import torch

from .depth_cloud import DepthCloud

from . import _lib as L

__all__ = ['feature_mask', 'filter_depth', 'filter_eigenvalue', 'filter_eigenvalue_ratio', 'filter_eigenvalue_ratios',
           'filter_eigenvalues', 'filter_shadow_points', 'filter_valid_neighbors', 'within_bounds']

_INF = float('inf')


def _side(v, default):
    return default if v is None else float(v)


def _mask_launch(vals, records, valid_counts=None, min_valid=0, mask=None):
    """records: [(kind, a, b, lo, hi)] -> bool [n]; `mask` (bool [n]) is ANDed in.  One kernel launch."""
    ref = vals if vals is not None else valid_counts
    n = ref.shape[0]
    dev = ref.device
    if mask is None:
        out = torch.empty(n, dtype=torch.bool, device=dev)
        init = 1
    else:
        out = mask.detach().to(device=dev, dtype=torch.bool).clone().contiguous()
        init = 0
    stride = 1
    code = L.DC_F64
    if vals is not None:
        vals = vals.detach()
        if vals.dim() == 1:
            vals = vals[:, None]
        vals = vals.contiguous()
        stride = vals.shape[1]
        code = L.dtype_code(vals.dtype)
    for first in range(0, max(len(records), 1), 16):
        chunk = records[first:first + 16]
        flat = [float(x) for rec in chunk for x in rec]
        table = (ctypes.c_double * max(len(flat), 1))(*flat)
        L.call('dc_feature_mask', L.ptr(vals) if chunk else None, code, n, stride, ctypes.cast(table, ctypes.c_void_p),
               len(chunk), L.ptr(valid_counts) if first == 0 else None, int(min_valid or 0), init,
               ctypes.c_void_p(out.data_ptr()), L.stream())
        init = 0
    return out


def _log_kept(keep, what, lo, hi):
    print('%.3f = %i / %i points kept (%.3g <= %s <= %.3g).'
          % (keep.double().mean(), keep.sum(), keep.numel(), lo if lo is not None else float('nan'), what,
             hi if hi is not None else float('nan')))


def within_bounds(x, min=None, max=None, bounds=None, log_variable=None):
    """Mask of x being within bounds  min <= x <= max (inclusive; None / +-inf disable a side; filters.py:85-113)."""
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    if bounds:
        assert min is None and max is None
        min, max = bounds
    if x.is_cuda and x.dtype in (torch.float32, torch.float64) and x.numel() > 0:
        keep = _mask_launch(x.reshape(-1), [(0, 0, 0, _side(min, -_INF), _side(max, _INF))])
    else:
        keep = torch.ones((x.numel(),), dtype=torch.bool, device=x.device)
        if min is not None and min > -_INF:
            keep = keep & (x.flatten() >= min)
        if max is not None and max < _INF:
            keep = keep & (x.flatten() <= max)
    if log_variable is not None:
        _log_kept(keep, log_variable, min, max)
    return keep


def feature_mask(cloud, eigenvalue_bounds=None, eigenvalue_ratio_bounds=None, min_valid_neighbors=None, mask=None):
    """AND of every feature bound of a configuration, one kernel launch:
    eigenvalue_bounds [(eig, min, max)] (filters.py:196-221), eigenvalue_ratio_bounds [(i, j, min, max)]
    (filters.py:224-254), min_valid_neighbors (filters.py:184-193), starting `mask` (bool [N] or None)."""
    records = []
    for eig, lo, hi in (eigenvalue_bounds or []):
        assert 0 <= eig <= 2
        records.append((0, eig, 0, _side(lo, -_INF), _side(hi, _INF)))
    for i, j, lo, hi in (eigenvalue_ratio_bounds or []):
        assert 0 <= i <= 2 and 0 <= j <= 2
        records.append((1, i, j, _side(lo, -_INF), _side(hi, _INF)))
    counts = None
    if min_valid_neighbors:
        counts = cloud.num_valid_neighbors().to(torch.int64).contiguous()
    vals = None
    if records:
        assert cloud.eigvals is not None
        vals = cloud.eigvals
    if vals is None and counts is None:
        n = cloud.size()
        return torch.ones((n,), dtype=torch.bool, device=cloud.device()) if mask is None else mask.clone()
    ref = vals if vals is not None else counts
    if not ref.is_cuda:
        # host tensors: the torch comparisons of the reference
        keep = torch.ones((ref.shape[0],), dtype=torch.bool) if mask is None else mask.clone()
        if counts is not None:
            keep = keep & (counts >= min_valid_neighbors)
        for kind, a, b, lo, hi in records:
            x = vals[:, a] if kind == 0 else vals[:, a] / vals[:, b]
            if lo > -_INF:
                keep = keep & (x >= lo)
            if hi < _INF:
                keep = keep & (x <= hi)
        return keep
    with torch.no_grad():
        return _mask_launch(vals, records, counts, min_valid_neighbors, mask)


def filter_depth(cloud, min=None, max=None, only_mask=False, log=False):
    """Keep points with depth in bounds (filters.py:116-141)."""
    assert isinstance(cloud, DepthCloud)
    keep = within_bounds(cloud.depth, min=min, max=max, log_variable='depth' if log else None)
    return keep if only_mask else cloud[keep]


def filter_valid_neighbors(cloud, min=None, only_mask=False, log=False):
    """Keep points with enough valid neighbors (filters.py:184-193)."""
    assert isinstance(cloud, DepthCloud)
    keep = feature_mask(cloud, min_valid_neighbors=min)
    if log:
        _log_kept(keep, 'valid neighbors', min, None)
    return keep if only_mask else cloud[keep]


def filter_eigenvalue(cloud, eigenvalue=0, min=None, max=None, only_mask=False, log=False):
    keep = feature_mask(cloud, eigenvalue_bounds=[(eigenvalue, min, max)])
    if log:
        _log_kept(keep, 'eigenvalue %i' % eigenvalue, min, max)
    return keep if only_mask else cloud[keep]


def filter_eigenvalues(cloud, bounds, only_mask=False, log=False):
    mask = feature_mask(cloud, eigenvalue_bounds=bounds)
    if log:
        for eig, lo, hi in (bounds or []):
            filter_eigenvalue(cloud, eig, min=lo, max=hi, only_mask=True, log=True)
    return mask if only_mask else cloud[mask]


def filter_eigenvalue_ratio(cloud, eigenvalues=(0, 1), min=None, max=None, only_mask=False, log=False):
    assert cloud.eigvals is not None
    assert len(eigenvalues) == 2
    i, j = eigenvalues
    keep = feature_mask(cloud, eigenvalue_ratio_bounds=[(i, j, min, max)])
    if log:
        _log_kept(keep, 'eigenvalue %i / eigenvalue %i' % tuple(eigenvalues), min, max)
    return keep if only_mask else cloud[keep]


def filter_eigenvalue_ratios(cloud, bounds, only_mask=False, log=False):
    mask = feature_mask(cloud, eigenvalue_ratio_bounds=bounds)
    if log:
        for i, j, lo, hi in (bounds or []):
            filter_eigenvalue_ratio(cloud, (i, j), min=lo, max=hi, only_mask=True, log=True)
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
