"""Voxel-grid filter (filter_grid, filters.py:24-82 of the reference) -- SURVEY.md section 8(f) row 1.

The reference walks a Python dict over all points; here the voxel keys are sorted on the device and
the first point of every voxel is kept (keep='first'; 'random' with a seeded permutation)."""
import torch

__all__ = ['filter_grid']


def filter_grid(cloud, grid_res, only_mask=False, keep='first', rng=None):
    assert grid_res > 0.0
    pts = cloud.get_points().detach() if hasattr(cloud, 'get_points') else cloud
    n = pts.shape[0]
    keys = torch.floor(pts.double() / grid_res).long()
    keys = keys - keys.min(dim=0).values
    dims = keys.max(dim=0).values + 1
    lin = (keys[:, 0] * dims[1] + keys[:, 1]) * dims[2] + keys[:, 2]
    if keep == 'random':
        g = torch.Generator(device='cpu')
        g.manual_seed(135 if rng is None else int(rng.integers(1 << 31)))
        perm = torch.randperm(n, generator=g).to(pts.device)
    else:
        perm = torch.arange(n, device=pts.device)
    order = torch.argsort(lin[perm], stable=True)
    sorted_lin = lin[perm][order]
    first = torch.ones(n, dtype=torch.bool, device=pts.device)
    first[1:] = sorted_lin[1:] != sorted_lin[:-1]
    kept = perm[order[first]]
    mask = torch.zeros(n, dtype=torch.bool, device=pts.device)
    mask[kept] = True
    if only_mask:
        return mask
    return cloud[mask]
