"""Shared helpers for the parity tests (test infrastructure)."""
import ast

import numpy as np
import torch

STEP_TAGS = ['scaled_mineig_norm_r', 'scaled_mineig_sqrt_kr', 'poly_trace_r', 'scaled_trace_k_common',
             'scaled_mineig_norm_sum']

# north_star tolerance: loss / eigenvalues / gradients within 1e-5 relative of the fp64 reference.
RTOL = 1e-5


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) elementwise."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if b.size else 0.0


def rel_err_norm(a, b):
    """||a-b||_inf / ||b||_inf : relative to the largest entry (for gradient vectors)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def step_inputs(g):
    """Unpack a tests/golden/step_*.npz into oracle-style inputs (fp32 values up-cast to fp64)."""
    from oracle import oracle
    S = int(g['n_scans'])
    scans = []
    for i in range(S):
        pts = torch.as_tensor(g['scan%d_points' % i].astype(np.float64))
        vps, dirs, depth = oracle.from_points(pts)
        scans.append({'vps': vps, 'dirs': dirs, 'depth': depth, 'points32': g['scan%d_points' % i],
                      'inc_angles': torch.as_tensor(g['scan%d_inc_angles' % i]),
                      'mask': torch.as_tensor(g['scan%d_mask' % i])})
    kw = ast.literal_eval(str(g['loss_kwargs']))
    return dict(
        scans=scans, poses=torch.as_tensor(g['poses']),
        neighbors=torch.as_tensor(g['neighbors'].astype(np.int64)),
        w=torch.as_tensor(g['w']).reshape(1, -1), exponent=torch.as_tensor(g['exponent']).reshape(1, -1),
        pose_deltas=torch.as_tensor(g['pose_deltas']) if 'pose_deltas' in g.files else None,
        loss_mask=torch.as_tensor(g['loss_mask']) if 'loss_mask' in g.files else None,
        loss=str(g['loss_name']), scaled=str(g['model']) == 'ScaledPolynomial',
        normalization=kw.get('normalization', False), sqrt=bool(kw.get('sqrt', False)),
        reduction=kw.get('reduction', 'mean'))


def eig_close(lam, ref, rtol):
    """Eigenvalue parity: |lam - ref| <= rtol * |ref| + 1e-13 * lambda_max(row).  The absolute term is the
    rounding floor of any symmetric eigen-solver (LAPACK included): eigenvalues of rank-deficient
    neighbourhoods (<= 3 points) are +-1e-17-ish noise in the reference as well."""
    lam = np.asarray(lam, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    tol = rtol * np.abs(ref) + 1e-13 * np.abs(ref).max(axis=1, keepdims=True)
    return bool(np.all(np.abs(lam - ref) <= tol))


def same_neighbor_sets(a, b):
    """Row-wise set equality of two padded index matrices (-1 = missing)."""
    a = np.sort(np.asarray(a), axis=1)
    b = np.sort(np.asarray(b), axis=1)
    if a.shape[1] != b.shape[1]:
        w = max(a.shape[1], b.shape[1])
        a = np.pad(a, ((0, 0), (w - a.shape[1], 0)), constant_values=-1)
        b = np.pad(b, ((0, 0), (w - b.shape[1], 0)), constant_values=-1)
        a, b = np.sort(a, axis=1), np.sort(b, axis=1)
    return np.array_equal(a, b)


def well_separated(eigvals, gap=1e-3):
    """Rows whose three eigenvalues are separated by more than `gap` relative to the largest
    (eigenvectors of nearly repeated eigenvalues are not comparable between solvers)."""
    ev = np.asarray(eigvals)
    scale = np.maximum(np.abs(ev[:, 2]), 1e-300)
    return ((ev[:, 1] - ev[:, 0]) > gap * scale) & ((ev[:, 2] - ev[:, 1]) > gap * scale) & np.isfinite(ev).all(axis=1)


def icp_inputs(g):
    """tests/golden/icp.npz -> per-scan fp64 tensors (float32 values) for the oracle / the CUDA path."""
    S = int(g['n_scans'])
    scans = []
    for i in range(S):
        scans.append({'points': torch.as_tensor(g['scan%d_points' % i].astype(np.float64)),
                      'inc_angles': torch.as_tensor(g['scan%d_inc_angles' % i]), 'mask': torch.as_tensor(g['scan%d_mask' % i]),
                      'normals': torch.as_tensor(g['scan%d_normals' % i])})
    masks = [(torch.as_tensor(g['mask%d_1' % i]), torch.as_tensor(g['mask%d_2' % i])) for i in range(S - 1)]
    return scans, torch.as_tensor(g['poses']), torch.as_tensor(g['w']).reshape(1, -1), torch.as_tensor(g['exponent']).reshape(1, -1), masks
