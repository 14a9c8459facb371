"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp64 + scipy) of the reference's
map-consistency hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product package
(depth_correction_b200) never does.

Parity status: PINNED.  Every function below is checked in tests/test_oracle.py against
golden vectors produced by running the unmodified reference (/root/reference, loaded by
oracle/ref_shim.py) in the build container (generator: oracle/make_golden.py, fixtures:
tests/golden/*.npz), and -- when /root/reference is present -- live against the reference.
One exception: `axis_angle_to_matrix` restates pytorch3d (not vendored, not installed,
unpinned `@stable` in the reference's docs/install.md:17); it is pinned to scipy's
Rotation.from_rotvec instead ("parity unpinned" against pytorch3d itself).

All `file:line` citations are relative to /root/reference/src/depth_correction/.
The arithmetic is fp64 throughout (reference default, config.py:179).
"""
import numpy as np
import torch
from scipy.spatial import cKDTree

F64 = torch.float64


# ----------------------------------------------------------------------------------------
# pytorch3d.transforms restatement (transform.py:4-9,73 call site)
# ----------------------------------------------------------------------------------------
def axis_angle_to_quaternion(axis_angle):
    """pytorch3d `axis_angle_to_quaternion`: q = (cos(t/2), w * sin(t/2)/t), with the
    Taylor branch sin(t/2)/t ~= 1/2 - t^2/48 for |t| < 1e-6."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    eps = 1e-6
    small = angles.abs() < eps
    safe = torch.where(small, torch.ones_like(angles), angles)
    s = torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / safe)
    return torch.cat([torch.cos(half), axis_angle * s], dim=-1)


def quaternion_to_matrix(q):
    """pytorch3d `quaternion_to_matrix` (real part first, two_s = 2 / |q|^2)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_matrix(axis_angle):
    return quaternion_to_matrix(axis_angle_to_quaternion(axis_angle))


def xyz_axis_angle_to_matrix(xyz_axis_angle):
    """transform.py:68-78."""
    lead = xyz_axis_angle.shape[:-1]
    bottom = torch.zeros(lead + (1, 4), dtype=xyz_axis_angle.dtype)
    bottom[..., 0, 3] = 1.0
    top = torch.cat([axis_angle_to_matrix(xyz_axis_angle[..., 3:]), xyz_axis_angle[..., :3, None]], dim=-1)
    return torch.cat([top, bottom], dim=-2)


def create_corrected_poses(poses, pose_deltas):
    """eval.py:68-82 for one sequence: poses[S,4,4] @ delta matrices ([S,6] or broadcast [1,6])."""
    return torch.matmul(poses, xyz_axis_angle_to_matrix(pose_deltas))


# ----------------------------------------------------------------------------------------
# Neighbour search (nearest_neighbors.py:22-80; scipy cKDTree is the un-vendored native
# dependency, pinned scipy==1.7.2 in python_requirements.txt:12, here scipy 1.18.1)
# ----------------------------------------------------------------------------------------
def nearest_neighbors(points, query=None, k=None, r=None, workers=-1):
    """Returns (dist or None, idx int64 [N,K]); missing neighbours are -1 / inf.

    kNN mode: distance-sorted, `distance_upper_bound=r` is strict.  Radius mode: `<= r`,
    rows ascending by point index, padded to the longest row with -1, dist is None.
    """
    assert k or r
    pts = np.ascontiguousarray(points.detach().cpu().numpy().reshape(-1, points.shape[-1]), dtype=np.float64)
    qry = pts if query is None else np.ascontiguousarray(
        query.detach().cpu().numpy().reshape(-1, points.shape[-1]), dtype=np.float64)
    tree = cKDTree(pts)
    if k:
        kw = {'distance_upper_bound': r} if r else {}
        dist, idx = tree.query(qry, k, workers=workers, **kw)
        if k == 1:
            dist, idx = dist.reshape(-1, 1), idx.reshape(-1, 1)
        idx = idx.astype(np.int64)
        idx[idx == tree.n] = -1
        return torch.from_numpy(dist), torch.from_numpy(idx)
    rows = tree.query_ball_point(qry, r, workers=workers, return_sorted=True)
    lens = np.fromiter((len(x) for x in rows), dtype=np.int64, count=len(rows))
    width = int(lens.max()) if len(rows) else 0
    idx = np.full((len(rows), width), -1, dtype=np.int64)
    # vectorised padding (the reference pads with a Python list comprehension, :69-73)
    col = np.arange(width)[None, :]
    keep = col < lens[:, None]
    idx[keep] = np.concatenate([np.asarray(x, dtype=np.int64) for x in rows]) if len(rows) else 0
    return None, torch.from_numpy(idx)


# ----------------------------------------------------------------------------------------
# covs / trace (utils.py:109-154)
# ----------------------------------------------------------------------------------------
def covs(x, weights):
    """x [N,K,3], weights [N,K,1] -> [N,3,3]; weighted, centred, Bessel W-1 clamped to 1e-6."""
    w = weights.sum(dim=-2, keepdim=True)
    xm = (weights * x).sum(dim=-2, keepdim=True) / w
    xc = x - xm
    xx = weights.unsqueeze(-1) * (xc.unsqueeze(-1) * xc.unsqueeze(-2))
    xx = xx.sum(dim=-3)
    return xx / (w - 1).clamp(1e-6, None)


def trace(x):
    return x.diagonal(dim1=-2, dim2=-1).sum(dim=-1)


# ----------------------------------------------------------------------------------------
# Model (model.py:149-286)
# ----------------------------------------------------------------------------------------
def model_bias(inc_angles, w, exponent):
    """model.py:243-248: pow(gamma, e) @ w^T, [n,1]."""
    return torch.matmul(torch.pow(inc_angles, exponent.reshape(1, -1)), w.reshape(1, -1).t()).reshape(-1, 1)


def correct_depth(depth, inc_angles, mask, w, exponent, scaled=True):
    """model.py:194-205 (Polynomial: d - bias) / :250-261 (ScaledPolynomial: d (1 - bias)).
    Only rows with mask are corrected when a mask is given."""
    if w is None:
        return depth
    if mask is None:
        bias = model_bias(inc_angles, w, exponent)
        return depth * (1.0 - bias) if scaled else depth - bias
    bias = model_bias(inc_angles[mask], w, exponent)
    out = depth.clone()
    out[mask] = out[mask] * (1.0 - bias) if scaled else out[mask] - bias
    return out


def invert_depth(depth, inc_angles, w, exponent):
    """ScaledPolynomial.inverse without mask (model.py:263-268): d / (1 - bias)."""
    return depth / (1.0 - model_bias(inc_angles, w, exponent))


# ----------------------------------------------------------------------------------------
# Cloud pipeline (depth_cloud.py, preproc.py)
# ----------------------------------------------------------------------------------------
def from_points(pts, vps=None):
    """depth_cloud.py:592-638 -> (vps, dirs, depth[N,1])."""
    vps = torch.zeros_like(pts) if vps is None else vps
    dirs = pts - vps
    depth = dirs.norm(dim=-1, keepdim=True)
    valid = depth[:, 0] > 0.0
    dirs = dirs.clone()
    dirs[valid] = dirs[valid] / depth[valid]
    return vps, dirs, depth


def global_points(scans, poses, w=None, exponent=None, scaled=True):
    """preproc.py:80-119 + depth_cloud.py:122-152, 536-575.

    scans: list of dicts with fp64 tensors 'vps','dirs' [n,3], 'depth','inc_angles' [n,1], 'mask' bool[n] or None.
    poses: [S,4,4] (already corrected).  Returns world points [N,3] and world dirs [N,3].
    """
    pts, dirs_w = [], []
    for s, sc in enumerate(scans):
        d = correct_depth(sc['depth'], sc.get('inc_angles'), sc.get('mask'), w, exponent, scaled)
        R, t = poses[s, :3, :3], poses[s, :3, 3:]
        vps = torch.matmul(sc['vps'], R.t()) + t.t()
        dirs = torch.matmul(sc['dirs'], R.t())
        pts.append(vps + d * dirs)
        dirs_w.append(dirs)
    return torch.cat(pts), torch.cat(dirs_w)


def neighborhood_features(points, neighbors, weights=None, dirs=None, eigvecs=True):
    """depth_cloud.py:291-295 (mean), :356-369 (weights, cov), :376-399 (eigh),
    :401-424 (normals, incidence angles).  neighbors int64 [N,K] with -1 padding, which
    gathers the LAST point with weight 0 exactly like the reference (depth_cloud.py:304)."""
    if weights is None:
        weights = (neighbors >= 0).float()[..., None]       # float32 [N,K,1] (depth_cloud.py:213)
    nb = points[neighbors]                                    # [N,K,3]
    wsum = weights.sum(dim=(-2, -1))[..., None]
    out = {'mean': (weights * nb).sum(dim=-2) / wsum}
    out['cov'] = covs(nb, weights)
    if eigvecs:
        out['eigvals'], out['eigvecs'] = torch.linalg.eigh(out['cov'])
    else:
        out['eigvals'] = torch.linalg.eigvalsh(out['cov'])
    if dirs is not None and eigvecs:
        n = out['eigvecs'][..., 0]
        n = -torch.sign((dirs * n).sum(dim=-1))[..., None] * n
        out['normals'] = n
        out['inc_angles'] = torch.arccos((dirs * n).sum(dim=-1).abs()).unsqueeze(-1)
    return out


def within_bounds(x, lo=None, hi=None):
    """filters.py:85-113 (bounds inclusive; None / +-inf disable a side)."""
    keep = torch.ones((x.numel(),), dtype=torch.bool)
    if lo is not None and lo > -float('inf'):
        keep = keep & (x.flatten() >= lo)
    if hi is not None and hi < float('inf'):
        keep = keep & (x.flatten() <= hi)
    return keep


def eigenvalue_masks(eigvals, eigenvalue_bounds=(), eigenvalue_ratio_bounds=()):
    """filters.py:196-254."""
    mask = torch.ones((eigvals.shape[0],), dtype=torch.bool)
    for i, lo, hi in eigenvalue_bounds or ():
        mask = mask & within_bounds(eigvals[:, i], lo, hi)
    for i, j, lo, hi in eigenvalue_ratio_bounds or ():
        mask = mask & within_bounds(eigvals[:, i] / eigvals[:, j], lo, hi)
    return mask


def valid_neighbor_mask(neighbors, min_valid):
    """filters.py:184-193."""
    return within_bounds((neighbors >= 0).sum(dim=-1), lo=min_valid)


# ----------------------------------------------------------------------------------------
# Losses (loss.py:125-150, 216-370)
# ----------------------------------------------------------------------------------------
# ----------------------------------------------------------------------------------------
# per-scan preprocessing filters
# ----------------------------------------------------------------------------------------
def filter_grid(x, grid_res, keep='random', preserve_order=False, rng=None):
    """filters.py:24-82: one point per occupied voxel -> kept indices (int64 array).

    The reference zips voxel keys (tuples of floor(x / grid_res), computed in x's own dtype) with the
    point indices of a sequence (reversed for 'first', shuffled by `rng` for 'random') into a dict: the
    LAST index of every key survives, and the dict yields them in the order the keys were FIRST seen
    (or sorted, with preserve_order)."""
    x = np.asarray(x)
    n = len(x)
    vox = np.floor(x / grid_res).astype(np.int64)
    seq = np.arange(n)
    if keep == 'first':
        seq = seq[::-1]
    elif keep == 'random':
        seq = np.arange(n)
        rng.shuffle(seq)
    vox = vox[seq]
    _, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    n_vox = int(inv.max()) + 1 if n else 0
    pos = np.arange(n)
    first = np.full(n_vox, n, dtype=np.int64)
    last = np.full(n_vox, -1, dtype=np.int64)
    np.minimum.at(first, inv, pos)
    np.maximum.at(last, inv, pos)
    kept = seq[last]
    return np.sort(kept) if preserve_order else kept[np.argsort(first, kind='stable')]


def shadow_mask(points, vps, dir_neighbors, dir_neighbor_weights, angle_bounds):
    """filters.py:257-309 -> bool mask of the points that are kept."""
    lo, hi = angle_bounds
    if lo is None or not (lo >= 0.0):
        lo = 0.0
    if hi is None or not (hi <= np.pi):
        hi = np.pi
    x = torch.as_tensor(points, dtype=F64)
    o = torch.as_tensor(vps, dtype=F64).expand_as(x)
    nb = torch.as_tensor(dir_neighbors)
    to_vp = (o - x)[:, None, :]
    to_nb = x[nb] - x[:, None, :]
    cos = torch.nn.functional.cosine_similarity(to_vp, to_nb, dim=-1)
    ang = torch.acos(cos)
    ang[torch.as_tensor(dir_neighbor_weights) != 1.0] = 0.5 * (lo + hi)
    return (ang.amin(dim=-1) >= lo) & (ang.amax(dim=-1) <= hi)


# ----------------------------------------------------------------------------------------
# ICP-style losses
# ----------------------------------------------------------------------------------------
def icp_pairs_loss(points, normals=None, inlier_ratio=0.5, point_to_plane=True, masks=None):
    """loss.py:407-559 for one sequence: mean over consecutive pairs of scans of the point-to-plane
    (0.5 * (1->2 + 2->1)) or point-to-point distance over the correspondences whose nearest-neighbour
    distance is within the `inlier_ratio` quantile.  points / normals: lists of [n,3] tensors (world frame);
    masks: optional list of (bool mask or indices into scan i, indices into scan i+1)."""
    total = 0.0
    n_pairs = len(points) - 1
    for i in range(n_pairs):
        a, b = points[i].float(), points[i + 1].float()           # loss.py:424-425, 509-510
        if masks is None:
            _, idx = cKDTree(b.detach().double().numpy()).query(a.detach().double().numpy(), k=1)
            idx = torch.as_tensor(idx, dtype=torch.int64)
            d = ((a - b[idx]) ** 2).sum(dim=-1).sqrt().detach()
            keep = d <= torch.nanquantile(d, inlier_ratio)
            sel1, sel2 = torch.nonzero(keep)[:, 0], idx[keep]
        else:
            m1, sel2 = torch.as_tensor(masks[i][0]), torch.as_tensor(masks[i][1]).long()
            sel1 = torch.nonzero(m1)[:, 0] if m1.dtype == torch.bool else m1.long()
        diff = b[sel2] - a[sel1]
        if point_to_plane:
            n1, n2 = normals[i][sel1], normals[i + 1][sel2]
            d12 = ((n1 * diff).sum(dim=-1, keepdim=True) * n1).norm(dim=-1).mean()
            d21 = ((n2 * -diff).sum(dim=-1, keepdim=True) * n2).norm(dim=-1).mean()
            total = total + 0.5 * (d12 + d21)
        else:
            total = total + diff.norm(dim=-1).mean()
    return total / n_pairs


def reduce(x, reduction='mean', only_finite=False, skip_nans=False):
    if only_finite:
        x = x[x.isfinite()]
    elif skip_nans:
        x = x[~x.isnan()]
    if reduction == 'mean':
        return x.mean()
    if reduction == 'sum':
        return x.sum()
    return x


def _finish_loss(loss, sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult, only_finite, skip_nans):
    if inlier_ratio < 1.0:
        q = torch.quantile(loss, inlier_ratio, dim=0)
        if inlier_loss_mult != 1.0:
            q = inlier_loss_mult * q
        inlier_max_loss = q if inlier_max_loss is None else torch.min(torch.as_tensor(inlier_max_loss, dtype=q.dtype), q)
    if inlier_max_loss is not None:
        loss = loss[loss <= inlier_max_loss]
    loss = torch.relu(loss)
    if sqrt:
        loss = torch.sqrt(loss)
    return reduce(loss, reduction, only_finite, skip_nans), loss


def min_eigval_loss(eigvals, mask=None, sqrt=False, normalization=False, reduction='mean',
                    inlier_max_loss=None, inlier_ratio=1.0, inlier_loss_mult=1.0,
                    only_finite=False, skip_nans=False):
    """loss.py:216-294 -> (reduced loss, per-point loss after mask)."""
    if mask is not None:
        eigvals = eigvals[mask]
    loss = eigvals[:, 0]
    if normalization:
        loss = loss / eigvals.sum(dim=-1).clamp(min=1e-6)
    return _finish_loss(loss, sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult, only_finite, skip_nans)


def trace_loss(cov, mask=None, sqrt=None, reduction='mean',
               inlier_max_loss=None, inlier_ratio=1.0, inlier_loss_mult=1.0,
               only_finite=False, skip_nans=False):
    """loss.py:297-370."""
    if mask is not None:
        cov = cov[mask]
    return _finish_loss(trace(cov), sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult,
                        only_finite, skip_nans)


# ----------------------------------------------------------------------------------------
# One training iteration, as scripts/model_poses_learning:119-135 runs it
# ----------------------------------------------------------------------------------------
def map_consistency_step(scans, poses, neighbors, w, exponent, pose_deltas=None, loss_mask=None,
                         loss='min_eigval_loss', scaled=True, normalization=True, sqrt=False,
                         reduction='mean', backward=True):
    """Fixed-graph step: corrected poses -> global cloud -> features -> loss [-> backward].

    Returns dict(loss, per_point, eigvals, cov, points, w_grad, exponent_grad, pose_deltas_grad, poses_grad).
    """
    w = w.detach().clone().requires_grad_(backward)
    exponent = exponent.detach().clone().requires_grad_(backward)
    poses = poses.detach().clone()
    if pose_deltas is not None:
        pose_deltas = pose_deltas.detach().clone().requires_grad_(backward)
        poses_c = create_corrected_poses(poses, pose_deltas)
        if backward:
            poses_c.retain_grad()
    else:
        poses_c = poses.requires_grad_(backward)
    points, _ = global_points(scans, poses_c, w, exponent, scaled)
    feats = neighborhood_features(points, neighbors, eigvecs=False)
    if loss == 'min_eigval_loss':
        val, per_point = min_eigval_loss(feats['eigvals'], loss_mask, sqrt=sqrt, normalization=normalization,
                                         reduction=reduction)
    else:
        val, per_point = trace_loss(feats['cov'], loss_mask, sqrt=sqrt, reduction=reduction)
    out = {'loss': val.detach(), 'per_point': per_point.detach(), 'eigvals': feats['eigvals'].detach(),
           'cov': feats['cov'].detach(), 'points': points.detach()}
    if backward:
        val.backward()
        out['w_grad'] = w.grad
        out['exponent_grad'] = exponent.grad
        out['pose_deltas_grad'] = None if pose_deltas is None else pose_deltas.grad
        out['poses_grad'] = poses_c.grad    # d loss / d corrected poses [S,4,4]
    return out
