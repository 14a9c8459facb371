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
    """Covariance matrices from samples along `obs_axis` (generic utility, utils.py:109-149).

    Neighbourhood covariances of a DepthCloud do not go through this dense [.., K, 3, 3] form;
    they use the gather kernel (ops.neighborhood_mean_cov)."""
    assert isinstance(x, torch.Tensor)
    assert obs_axis != var_axis
    assert weights is None or isinstance(weights, torch.Tensor)
    w = weights.sum(dim=obs_axis, keepdim=True) if weights is not None else x.shape[obs_axis]
    if center:
        xm = (weights * x).sum(dim=obs_axis, keepdim=True) / w if weights is not None else x.mean(dim=obs_axis, keepdim=True)
        xc = x - xm
    else:
        xc = x
    var_axis_2 = var_axis + 1 if var_axis >= 0 else var_axis - 1
    xx = xc.unsqueeze(var_axis) * xc.unsqueeze(var_axis_2)
    if weights is not None:
        xx = weights.unsqueeze(var_axis) * xx
    if obs_axis < var_axis and obs_axis < 0:
        obs_axis -= 1
    elif obs_axis > var_axis and obs_axis > 0:
        obs_axis += 1
    xx = xx.sum(dim=obs_axis)
    if correction:
        w = w - 1
    if isinstance(w, torch.Tensor) and w.dtype.is_floating_point:
        w = w.clamp(1e-6, None)
    return xx / w


def trace(x, dim1=-2, dim2=-1):
    return x.diagonal(dim1=dim1, dim2=dim2).sum(dim=-1)
