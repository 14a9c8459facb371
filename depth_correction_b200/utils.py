"""Math helpers of the reference's utils.py that live on the hot path (utils.py:54-64, 109-154)."""
from timeit import default_timer as timer

import torch

__all__ = ['covs', 'timing', 'trace']


def timing(f):
    def timing_wrapper(*args, **kwargs):
        t0 = timer()
        try:
            return f(*args, **kwargs)
        finally:
            print('%s %.6f s' % (f.__name__, timer() - t0))
    return timing_wrapper


def covs(x, obs_axis=-2, var_axis=-1, center=True, correction=True, weights=None):
    """Covariance matrices of samples laid out along `obs_axis` with variables along `var_axis` (generic utility with
    the semantics of utils.py:109-149: optional per-sample weights, Bessel correction on the weight sum, a floor of
    1e-6 on a floating-point normaliser).  Neighbourhood covariances of a DepthCloud do not go through this dense
    [.., K, 3, 3] form; they use the gather kernel (ops.neighborhood_mean_cov)."""
    assert isinstance(x, torch.Tensor)
    assert obs_axis != var_axis
    assert weights is None or isinstance(weights, torch.Tensor)
    # bring the data to [..., observations, variables]
    xs = x.movedim((obs_axis, var_axis), (-2, -1))
    if weights is None:
        norm = xs.shape[-2]
        mean = xs.mean(dim=-2, keepdim=True)
    else:
        ws = weights.movedim((obs_axis, var_axis), (-2, -1)) if weights.dim() == x.dim() else weights
        norm = ws.sum(dim=-2, keepdim=True)
        mean = (ws * xs).sum(dim=-2, keepdim=True) / norm
        norm = norm.squeeze(-2).unsqueeze(-1) if norm.dim() >= 2 else norm
    d = xs - mean if center else xs
    if weights is None:
        scatter = torch.einsum('...ki,...kj->...ij', d, d)
    else:
        scatter = torch.einsum('...ki,...kj->...ij', ws * d, d)
    if correction:
        norm = norm - 1
    if isinstance(norm, torch.Tensor) and norm.dtype.is_floating_point:
        norm = norm.clamp(1e-6, None)
    return scatter / norm


def trace(x, dim1=-2, dim2=-1):
    return x.diagonal(dim1=dim1, dim2=dim2).sum(dim=-1)
