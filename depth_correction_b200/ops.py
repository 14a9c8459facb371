"""Differentiable per-op wrappers of the unfused feature kernels (dc_features.cu).

These back the staged DepthCloud API (update_mean / update_cov / update_eig / update_normals /
update_incidence_angles) so that code written against the reference keeps working, including
autograd through the stages.  The training loop uses the fused kernels (fused.py) instead.
"""
import torch

from . import _lib as L

__all__ = ['neighborhood_mean_cov', 'eigh3', 'normals_and_angles', 'pose_compose']


def _check(points, neighbors):
    assert points.is_cuda, 'depth_correction_b200 kernels need CUDA tensors (no CPU fallback)'
    assert neighbors.dtype == torch.int64 and neighbors.dim() == 2 and neighbors.shape[0] == points.shape[0]


class _MeanCov(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, neighbors, weights, want_mean, want_cov):
        ctx.set_materialize_grads(False)
        _check(points, neighbors)
        pts = points.detach().contiguous()
        nb = neighbors.contiguous()
        wt = None if weights is None else weights.detach().reshape(nb.shape).to(torch.float32).contiguous()
        n, K = nb.shape
        mean = torch.empty((n, 3), dtype=pts.dtype, device=pts.device) if want_mean else None
        cov = torch.empty((n, 3, 3), dtype=pts.dtype, device=pts.device) if want_cov else None
        L.call('dc_features', L.ptr(pts), L.dtype_code(pts.dtype), n, L.ptr(nb), L.ptr(wt), K, L.ptr(mean), L.ptr(cov),
               L.stream())
        ctx.save_for_backward(pts, nb, wt)
        ctx.want = (want_mean, want_cov)
        outs = tuple(x for x in (mean, cov) if x is not None)
        return outs if len(outs) > 1 else outs[0]

    @staticmethod
    def backward(ctx, *grads):
        pts, nb, wt = ctx.saved_tensors
        want_mean, want_cov = ctx.want
        grads = list(grads)
        gmean = grads.pop(0) if want_mean else None
        gcov = grads.pop(0) if want_cov else None
        gmean = None if gmean is None else gmean.to(pts.dtype).contiguous()
        gcov = None if gcov is None else gcov.to(pts.dtype).contiguous()
        gp = torch.zeros_like(pts)
        n, K = nb.shape
        L.call('dc_features_backward', L.ptr(pts), L.dtype_code(pts.dtype), n, L.ptr(nb), L.ptr(wt), K, L.ptr(gmean),
               L.ptr(gcov), L.ptr(gp), L.stream())
        return gp, None, None, None, None


def neighborhood_mean_cov(points, neighbors, weights=None, mean=True, cov=True):
    """Weighted neighbourhood mean [N,3] and covariance [N,3,3] (depth_cloud.py:291-295, utils.py:109-149).
    weights: float [N,K] / [N,K,1] or None (= neighbors >= 0); treated as constants by autograd."""
    return _MeanCov.apply(points, neighbors, weights, mean, cov)


class _Eigh3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cov):
        ctx.set_materialize_grads(False)
        assert cov.is_cuda and cov.shape[-2:] == (3, 3)
        c = cov.detach().contiguous()
        n = c.shape[0]
        eigvals = torch.empty((n, 3), dtype=c.dtype, device=c.device)
        eigvecs = torch.empty((n, 3, 3), dtype=c.dtype, device=c.device)
        L.call('dc_eigh3', L.ptr(c), L.dtype_code(c.dtype), n, L.ptr(eigvals), L.ptr(eigvecs), L.stream())
        ctx.save_for_backward(eigvals, eigvecs)
        return eigvals, eigvecs

    @staticmethod
    def backward(ctx, gl, gv):
        eigvals, eigvecs = ctx.saved_tensors
        n = eigvals.shape[0]
        gl = None if gl is None else gl.to(eigvals.dtype).contiguous()
        gv = None if gv is None else gv.to(eigvals.dtype).contiguous()
        gcov = torch.empty_like(eigvecs)
        L.call('dc_eigh3_backward', L.ptr(eigvals), L.ptr(eigvecs), L.dtype_code(eigvals.dtype), n, L.ptr(gl), L.ptr(gv),
               L.ptr(gcov), L.stream())
        return gcov


def eigh3(cov):
    """Batched symmetric 3x3 eigen-decomposition, ascending eigenvalues, eigenvectors in columns
    (replaces torch.linalg.eigh on the host, depth_cloud.py:376-399)."""
    return _Eigh3.apply(cov)


def normals_and_angles(dirs, eigvecs, use_normal_sign=False, want_normals=True, want_angles=True):
    """normals = -sign(dirs . v0) v0; inc_angles = arccos(|dirs . n|) (depth_cloud.py:401-424).  No autograd."""
    d = dirs.detach().contiguous()
    v = eigvecs.detach().to(d.dtype).contiguous()
    n = d.shape[0]
    normals = torch.empty((n, 3), dtype=d.dtype, device=d.device) if want_normals else None
    inc = torch.empty((n, 1), dtype=d.dtype, device=d.device) if want_angles else None
    L.call('dc_normals_angles', L.ptr(d), L.ptr(v), L.dtype_code(d.dtype), n, int(bool(use_normal_sign)), L.ptr(normals),
           L.ptr(inc), L.stream())
    return normals, inc


class _PoseCompose(torch.autograd.Function):
    @staticmethod
    def forward(ctx, poses, deltas):
        S = poses.shape[0]
        p = poses.detach().to(torch.float64).reshape(S, 16).contiguous()
        d = deltas.detach().to(torch.float64).reshape(-1, 6).contiguous()
        out12 = torch.empty((S, 12), dtype=torch.float64, device=p.device)
        L.call('dc_pose_compose', L.ptr(p), L.ptr(d), S, d.shape[0], L.ptr(out12), L.stream())
        ctx.save_for_backward(p, d)
        ctx.meta = (deltas.shape, deltas.dtype, poses.dtype)
        out = torch.zeros((S, 4, 4), dtype=torch.float64, device=p.device)
        out[:, :3, :] = out12.reshape(S, 3, 4)
        out[:, 3, 3] = 1.0
        return out.to(poses.dtype)

    @staticmethod
    def backward(ctx, gout):
        p, d = ctx.saved_tensors
        S = p.shape[0]
        g12 = gout.to(torch.float64)[:, :3, :].reshape(S, 12).contiguous()
        gd = torch.empty_like(d)
        L.call('dc_pose_compose_backward', L.ptr(p), L.ptr(d), S, d.shape[0], L.ptr(g12), L.ptr(gd), L.stream())
        shape, dtype, _ = ctx.meta
        return None, gd.reshape(shape).to(dtype)


def pose_compose(poses, deltas):
    """poses [S,4,4] @ xyz_axis_angle_to_matrix(deltas [S,6] or [1,6]) (eval.py:68-82), differentiable
    w.r.t. deltas; the poses themselves are constants of the optimisation."""
    return _PoseCompose.apply(poses, deltas)
