"""ICP-style losses between consecutive scans: icp_loss / point_to_plane_dist / point_to_point_dist
(loss.py:373-565 of the reference), same signatures.

Per pair of scans: nearest neighbour of every point of scan 1 in scan 2 (dc_knn, k = 1, cross query), inlier
threshold = nanquantile of the distances (radix sort of order-preserving keys), then ONE kernel for the residuals
of the kept correspondences and ONE for their gradients to both point sets and both normal sets (dc_icp_forward /
dc_icp_backward).  The reference's `differentiable` flag selects pytorch3d's fp32 brute-force search or scipy's
exact one; both feed the same loss (gradients only flow through the gathered points and normals), and the exact
search is what runs here.
"""
import warnings

import torch

from . import _lib as L
from .depth_cloud import DepthCloud
from .graph import search

__all__ = ['icp_loss', 'point_to_plane_dist', 'point_to_point_dist', 'nanquantile']


def nanquantile(x, q):
    """torch.nanquantile(x, q) (linear interpolation) of a 1-D fp64 CUDA tensor through the library's radix sort."""
    x = x.detach().reshape(-1).to(torch.float64).contiguous()
    n = x.numel()
    dev = x.device
    st = L.stream()
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    skeys = torch.empty(n, dtype=torch.int64, device=dev)
    n_nan = torch.zeros(1, dtype=torch.int32, device=dev)
    L.call('dc_f64_sort_keys', L.ptr(x), n, L.ptr(keys), L.ptr(n_nan), st)
    L.call_with_temp('dc_sort_keys', dev, L.ptr(keys), L.ptr(skeys), n, 0, 64, after=(st,))
    srt = torch.empty(n, dtype=torch.float64, device=dev)
    L.call('dc_f64_from_sort_keys', L.ptr(skeys), n, L.ptr(srt), st)
    m = n - int(n_nan.item())
    if m <= 0:
        return torch.full((), float('nan'), dtype=torch.float64, device=dev)
    rank = q * (m - 1)
    lo = int(rank)                       # floor (rank >= 0)
    hi = min(lo + 1, m - 1) if rank > lo else lo
    return torch.lerp(srt[lo], srt[hi], rank - lo)


class _IcpPair(torch.autograd.Function):
    """loss of one pair of scans; inputs: points1, points2, normals1, normals2 (normals may be None)."""

    @staticmethod
    def forward(ctx, p1, p2, n1, n2, nn, dist, th, sel1, sel2, point_to_plane):
        dev = p1.device
        st = L.stream()
        a = p1.detach().contiguous()
        b = p2.detach().contiguous()
        if a.dtype not in (torch.float32, torch.float64):
            a = a.float()
        b = b.to(a.dtype)
        na = nb = None
        ncode = L.DC_F64
        if point_to_plane:
            na = n1.detach().contiguous()
            nb = n2.detach().to(na.dtype).contiguous()
            ncode = L.dtype_code(na.dtype)
        m = a.shape[0] if sel1 is None else sel1.numel()
        blocks = (m + 255) // 256
        out = torch.empty(4, dtype=torch.float64, device=dev)
        partials = torch.zeros(4 * blocks + 2, dtype=torch.float64, device=dev)
        L.call('dc_icp_forward', L.ptr(a), L.ptr(b), L.dtype_code(a.dtype), L.ptr(na), L.ptr(nb), ncode, L.ptr(nn), L.ptr(dist),
               float(th), L.ptr(sel1), L.ptr(sel2), m, 1 if point_to_plane else 0, L.ptr(out), L.ptr(partials),
               partials.numel() * 8, st)
        ctx.saved = (a, b, na, nb, nn, dist, float(th), sel1, sel2, m, ncode, bool(point_to_plane), out)
        ctx.meta = (p1.dtype, p2.dtype, None if n1 is None else n1.dtype, None if n2 is None else n2.dtype)
        if point_to_plane:
            loss = 0.5 * (out[0] + out[1]) / out[2]          # 0.5 * (dist12 + dist21), loss.py:465
        else:
            loss = out[0] / out[2]
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        a, b, na, nb, nn, dist, th, sel1, sel2, m, ncode, p2pl, out = ctx.saved
        dev = a.device
        need = ctx.needs_input_grad
        c = (0.5 if p2pl else 1.0) * g_loss.to(torch.float64) / out[2]
        coef = torch.stack([c, c]).contiguous()
        g1 = torch.zeros((a.shape[0], 3), dtype=torch.float64, device=dev) if need[0] else None
        g2 = torch.zeros((b.shape[0], 3), dtype=torch.float64, device=dev) if need[1] else None
        gn1 = torch.zeros((a.shape[0], 3), dtype=torch.float64, device=dev) if (p2pl and need[2]) else None
        gn2 = torch.zeros((b.shape[0], 3), dtype=torch.float64, device=dev) if (p2pl and need[3]) else None
        L.call('dc_icp_backward', L.ptr(a), L.ptr(b), L.dtype_code(a.dtype), L.ptr(na), L.ptr(nb), ncode, L.ptr(nn), L.ptr(dist), th,
               L.ptr(sel1), L.ptr(sel2), m, 1 if p2pl else 0, L.ptr(coef), L.ptr(g1), L.ptr(g2), L.ptr(gn1), L.ptr(gn2), L.stream())
        dt1, dt2, dn1, dn2 = ctx.meta
        cast = lambda g, dt: None if g is None else g.to(dt)
        return cast(g1, dt1), cast(g2, dt2), cast(gn1, dn1), cast(gn2, dn2), None, None, None, None, None, None


def _points_of(cloud):
    if isinstance(cloud, DepthCloud):
        return cloud.to_points() if cloud.points is None else cloud.points
    return cloud


def _pair_loss(cloud1, cloud2, icp_inlier_ratio, mask, point_to_plane, verbose, i):
    points1, points2 = _points_of(cloud1), _points_of(cloud2)
    if not points1.is_cuda:
        raise RuntimeError('ICP losses need CUDA clouds; there is no CPU fallback')
    n1 = n2 = None
    if point_to_plane:
        assert cloud1.normals is not None, 'Cloud must have normals computed to estimate point to plane distance'
        n1, n2 = cloud1.normals, cloud2.normals
    nn = dist = sel1 = sel2 = None
    th = 0.0
    if mask is None:
        # nearest neighbour of every point of cloud 1 in cloud 2, on the float32 values the reference searches
        p1f = points1.detach().float()
        p2f = points2.detach().float()
        g = search(p2f, p1f, k=1)
        nn = g.neighbors().reshape(-1).contiguous()
        dist = g.distances().reshape(-1).contiguous()
        th = nanquantile(dist, icp_inlier_ratio).item()
    else:
        mask1, mask2 = mask
        mask1 = torch.as_tensor(mask1, device=points1.device)
        sel1 = (torch.nonzero(mask1)[:, 0] if mask1.dtype == torch.bool else mask1.long()).contiguous()
        sel2 = torch.as_tensor(mask2, device=points1.device).long().contiguous()
        assert sel1.numel() == sel2.numel()
    loss, out = _IcpPair.apply(points1, points2, n1, n2, nn, dist, th, sel1, sel2, point_to_plane)
    stats = out.tolist() if (verbose or mask is None) else None
    if stats is not None:
        assert stats[2] > 0, 'Point clouds do not intersect. Try to sample lidar scans more frequently'
        inl_err = stats[3] / stats[2] if mask is None else -1.0
        if inl_err > 0.3:
            warnings.warn('ICP inliers error is too big: %.3f (> 0.3) [m] for pairs (%i, %i)' % (inl_err, i, i + 1))
        if verbose:
            print('Mean point to %s distance: %.3f [m] for scans: (%i, %i), inliers error: %.6f'
                  % ('plane' if point_to_plane else 'point', loss.item(), i, i + 1, inl_err))
    return loss


def _consecutive_pairs(clouds, icp_inlier_ratio, masks, point_to_plane, verbose):
    assert 0.0 <= icp_inlier_ratio <= 1.0
    if masks is not None:
        assert len(clouds) == len(masks) + 1
    n_pairs = len(clouds) - 1
    total = 0.0
    for i in range(n_pairs):
        total = total + _pair_loss(clouds[i], clouds[i + 1], icp_inlier_ratio, None if masks is None else masks[i],
                                   point_to_plane, verbose, i)
    return torch.as_tensor(total / n_pairs)


def point_to_plane_dist(clouds, icp_inlier_ratio=0.5, masks=None, differentiable=True, verbose=False, **kwargs):
    """ICP-like point to plane distance over consecutive pairs of scans (loss.py:407-479)."""
    return _consecutive_pairs(clouds, icp_inlier_ratio, masks, True, verbose)


def point_to_point_dist(clouds, icp_inlier_ratio=0.5, masks=None, differentiable=True, verbose=False, **kwargs):
    """ICP-like point to point distance over consecutive pairs of scans (loss.py:482-559)."""
    return _consecutive_pairs(clouds, icp_inlier_ratio, masks, False, verbose)


def icp_loss(clouds, poses=None, model=None, masks=None, **kwargs):
    """ICP-like loss over lists of sequences of scans (loss.py:373-404): returns (loss, [concatenated cloud per sequence])."""
    transformed = clouds
    if model is not None:
        transformed = [[model(c) for c in seq] for seq in transformed]
    if poses is not None:
        transformed = [[c.transform(p) for c, p in zip(seq, seq_poses)] for seq, seq_poses in zip(transformed, poses)]
    loss = 0.
    loss_cloud = []
    loss_fun = point_to_plane_dist if kwargs['icp_point_to_plane'] else point_to_point_dist
    for i, seq in enumerate(transformed):
        seq_masks = None if masks is None else masks[i]
        loss = loss + loss_fun(seq, masks=seq_masks, **kwargs)
        cloud = DepthCloud.concatenate(seq)
        cloud.loss = loss
        loss_cloud.append(cloud)
    loss = loss / len(transformed)
    return loss, loss_cloud
