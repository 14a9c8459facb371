"""Voxel-grid filter (filter_grid, filters.py:24-82 of the reference) -- SURVEY.md section 8(f) row 1.

The reference walks a Python dict over all points (tuple keys, last value wins, dict order out).  Here the voxel keys
are computed, grouped (radix sort) and the survivors picked and ordered on the device (dc_voxel_keys / dc_sort_pairs
/ dc_voxel_pick); the only host work is the seeded permutation of keep='random', drawn from the same numpy
Generator the reference shuffles with, so that the same points survive in the same order.
"""
import numpy as np
import torch

from . import _lib as L

__all__ = ['default_rng', 'filter_grid']

default_rng = np.random.default_rng(135)      # filters.py:20


def filter_grid(cloud, grid_res, only_mask=False, keep='random', preserve_order=False, log=False, rng=default_rng):
    """Keep a single point within each cell (filters.py:24-82).  Returns the filtered cloud, or (only_mask) the
    indices of the kept points (int64 tensor; the reference returns them as a Python list)."""
    from .depth_cloud import DepthCloud
    assert isinstance(cloud, (DepthCloud, np.ndarray, torch.Tensor))
    assert isinstance(grid_res, float) and grid_res > 0.0
    assert keep in ('first', 'random', 'last')
    if isinstance(cloud, DepthCloud):
        x = cloud.get_points().detach()
    elif isinstance(cloud, np.ndarray):
        if cloud.dtype.names:
            x = torch.as_tensor(np.stack([cloud[f] for f in ('x', 'y', 'z')], axis=-1))
        else:
            x = torch.as_tensor(cloud)
    else:
        x = cloud.detach()
    if not x.is_cuda:
        raise RuntimeError('filter_grid needs a CUDA cloud / tensor; there is no CPU fallback')
    x = x.reshape(-1, 3).contiguous()
    n = x.shape[0]
    dev = x.device
    st = L.stream()
    seq = None
    if keep == 'random':
        # the reference shuffles the index list with this generator (and advances its state by one shuffle)
        perm = np.arange(n)
        rng.shuffle(perm)
        seq = L.upload(perm.astype(np.int32), torch.int32, dev)
    reversed_ = 1 if keep == 'first' else 0       # "make the first item last" (filters.py:49-52)
    if n == 0:
        ind = torch.zeros(0, dtype=torch.int64, device=dev)
    else:
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        ids = torch.empty(n, dtype=torch.int32, device=dev)
        skeys = torch.empty(n, dtype=torch.int64, device=dev)
        sids = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)     # [bad points, voxels]
        L.call('dc_voxel_keys', L.ptr(x), L.dtype_code(x.dtype), n, float(grid_res), L.ptr(seq), reversed_, L.ptr(keys),
               L.ptr(ids), L.ptr(cnt[0:1]), st)
        L.call_with_temp('dc_sort_pairs', dev, L.ptr(keys), L.ptr(skeys), L.ptr(ids), L.ptr(sids), n, 63, after=(st,))
        L.call('dc_voxel_pick', L.ptr(skeys), L.ptr(sids), n, L.ptr(seq), reversed_, 1 if preserve_order else 0, L.ptr(keys),
               L.ptr(ids), L.ptr(cnt[1:2]), st)
        L.call_with_temp('dc_sort_pairs', dev, L.ptr(keys), L.ptr(skeys), L.ptr(ids), L.ptr(sids), n, 33, after=(st,))
        bad, n_vox = cnt.tolist()                                # the one read-back: the output size is data dependent
        if bad:
            raise ValueError('filter_grid: %i points are not finite or farther than 2^20 cells from the origin' % bad)
        ind = sids[:n_vox].long()
    if log:
        print('%.3f = %i / %i points kept (grid res. %.3f m).' % (len(ind) / max(n, 1), len(ind), n, grid_res))
    if only_mask:
        return ind
    if isinstance(cloud, np.ndarray):
        return cloud[ind.cpu().numpy()]
    return cloud[ind]
