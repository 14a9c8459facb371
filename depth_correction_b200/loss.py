"""Map-consistency losses: drop-in for min_eigval_loss / trace_loss / batch_loss / reduce
(loss.py:125-150, 181-370 of the reference), same signatures and return value `(loss, cloud)`.

Dispatch
  * lazy GlobalCloud with a fixed graph (the training loop)  -> fused kernels 2/3 (fused.py):
      - mean / sum reduction without inlier selection -> loss reduced inside the kernel (fast path)
      - anything else (inlier_ratio, inlier_max_loss, offset, only_finite, skip_nans, reduction none)
        -> per-point raw loss from the kernel, remaining element-wise steps in torch, backward through
        the same kernel with a per-point upstream gradient
  * any other cloud with `eigvals` / `cov` already present -> the reference's element-wise arithmetic.
"""
from enum import Enum

import torch

from . import _lib as L
from .depth_cloud import DepthCloud
from .fused import fused_loss
from .utils import trace

__all__ = ['batch_loss', 'create_loss', 'loss_by_name', 'min_eigval_loss', 'reduce', 'Reduction', 'trace_loss']


class Reduction(Enum):
    NONE = 'none'
    MEAN = 'mean'
    SUM = 'sum'


def reduce(x, reduction=Reduction.MEAN, weights=None, only_finite=False, skip_nans=False):
    assert reduction in Reduction
    keep = None
    if only_finite:
        keep = x.isfinite()
    elif skip_nans:
        keep = ~x.isnan()
    if keep is not None:
        if weights is not None:
            weights = weights[keep]
        x = x[keep]
    if reduction == Reduction.MEAN:
        x = x.mean() if weights is None else (weights * x).sum() / weights.sum()
    elif reduction == Reduction.SUM:
        x = x.sum() if weights is None else (weights * x).sum()
    return x


def batch_loss(loss_fun, clouds, masks=None, offsets=None, reduction=Reduction.MEAN,
               only_finite=False, skip_nans=False, **kwargs):
    """General batch loss of a sequence of clouds (loss.py:181-213)."""
    assert callable(loss_fun)
    assert isinstance(clouds, (list, tuple))
    if masks is None:
        masks = len(clouds) * [None]
    if offsets is None:
        offsets = len(clouds) * [None]
    assert isinstance(masks, (list, tuple)) and len(masks) == len(clouds)
    assert isinstance(offsets, (list, tuple)) and len(offsets) == len(clouds)
    simple = (reduction in (Reduction.MEAN, Reduction.SUM) and not only_finite and not skip_nans
              and all(o is None for o in offsets) and _simple_kwargs(kwargs)
              and all(_is_fusable(c) for c in clouds))
    if simple:
        # sum of in-kernel partial sums over the clouds, one division at the end
        parts, loss_clouds = [], []
        for cloud, mask in zip(clouds, masks):
            out, loss_cloud = _fused_reduced(loss_fun, cloud, mask, kwargs)
            parts.append(out)
            loss_clouds.append(loss_cloud)
        tot = torch.stack(parts).sum(dim=0)
        loss = tot[0] / tot[1] if reduction == Reduction.MEAN else tot[0]
        return loss, loss_clouds
    losses, loss_clouds = [], []
    for cloud, mask, offset in zip(clouds, masks, offsets):
        loss, loss_cloud = loss_fun(cloud, mask=mask, offset=offset, reduction=Reduction.NONE, **kwargs)
        losses.append(loss)
        loss_clouds.append(loss_cloud)
    loss = reduce(torch.cat(losses), reduction=reduction, only_finite=only_finite, skip_nans=skip_nans)
    return loss, loss_clouds


def _simple_kwargs(kw):
    return (kw.get('inlier_max_loss') is None and kw.get('inlier_ratio', 1.0) >= 1.0)


def _is_fusable(cloud):
    return hasattr(cloud, 'fusable') and cloud.fusable()


def _kind_flags(loss_fun, kw):
    if loss_fun is trace_loss:
        return L.LOSS_TRACE, (L.FLAG_SQRT if kw.get('sqrt') else 0)
    flags = (L.FLAG_SQRT if kw.get('sqrt') else 0) | (L.FLAG_NORMALIZATION if kw.get('normalization') else 0)
    return L.LOSS_MIN_EIGVAL, flags


class _LazyLossCloud(object):
    """The `cloud` half of the `(loss, cloud)` return value: a shallow copy whose per-point `.loss`
    (after the mask, like cloud[mask].loss in the reference) is gathered from the kernel output on demand."""

    def __init__(self, cloud, state, mask):
        self._cloud, self._state, self._mask = cloud, state, mask
        self._gen = state.generation
        self._loss = None

    @property
    def loss(self):
        if self._loss is None:
            if self._state.generation != self._gen:
                raise RuntimeError('per-point loss was overwritten by a newer forward pass')
            pp = self._state.loss_pp[self._state.graph.map.inv_order.long()]
            self._loss = pp if self._mask is None else pp[self._mask]
        return self._loss

    def __getattr__(self, name):
        c = self.__dict__['_cloud']
        m = self.__dict__['_mask']
        if m is not None and name in DepthCloud.sliced_fields:
            x = getattr(c, name)
            return None if x is None else x[m]
        return getattr(c, name)


def _fused_reduced(loss_fun, cloud, mask, kw):
    state = cloud.step_state()
    kind, flags = _kind_flags(loss_fun, kw)
    out = fused_loss(state, cloud._model, cloud.poses_tensor(), kind, flags, mask=mask)
    return out, _LazyLossCloud(cloud, state, mask)


def fused_sum_count(cloud, mask=None, loss='min_eigval_loss', sqrt=False, normalization=False):
    """Tensor [2] = (sum of per-point losses over `mask`, number of masked points) from the fused kernels, with
    autograd history.  Building block of multi-GPU reductions (parallel.reduce_step): the mean over all ranks
    is sum-of-sums / sum-of-counts."""
    assert _is_fusable(cloud), 'fused_sum_count needs a lazy global cloud with a fixed graph on the GPU'
    loss_fun = trace_loss if loss in ('trace_loss', trace_loss) else min_eigval_loss
    out, _ = _fused_reduced(loss_fun, cloud, mask, dict(sqrt=sqrt, normalization=normalization))
    return out


def _finish(cloud, loss, mask, offset, sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult,
            only_finite, skip_nans):
    """Element-wise tail shared by both losses (loss.py:256-293 / 332-369)."""
    if inlier_ratio < 1.0:
        assert offset is None
        loss_quantile = torch.quantile(loss, inlier_ratio, dim=0)
        if inlier_loss_mult != 1.0:
            loss_quantile = inlier_loss_mult * loss_quantile
        if inlier_max_loss is None:
            inlier_max_loss = loss_quantile
        else:
            inlier_max_loss = torch.min(torch.as_tensor(inlier_max_loss, dtype=loss.dtype, device=loss.device), loss_quantile)
    if inlier_max_loss is not None:
        assert offset is None
        mask = (loss <= inlier_max_loss)
    if mask is not None:
        cloud = cloud[mask]
        loss = loss[mask]
    if offset is not None:
        loss = loss - offset
    loss = torch.relu(loss)
    if sqrt:
        loss = torch.sqrt(loss)
    cloud = cloud.copy()
    cloud.loss = loss
    loss = reduce(loss, reduction=reduction, only_finite=only_finite, skip_nans=skip_nans)
    return loss, cloud


def _generic_loss(loss_fun, cloud, mask, offset, sqrt, normalization, reduction, inlier_max_loss, inlier_ratio,
                  inlier_loss_mult, only_finite, skip_nans):
    assert isinstance(cloud, DepthCloud) or hasattr(cloud, 'eigvals')
    assert offset is None or isinstance(offset, (DepthCloud, torch.Tensor))
    kw = dict(sqrt=sqrt, normalization=normalization)
    simple = (reduction in (Reduction.MEAN, Reduction.SUM) and offset is None and not only_finite and not skip_nans
              and inlier_max_loss is None and inlier_ratio >= 1.0)
    if _is_fusable(cloud):
        if simple:
            out, loss_cloud = _fused_reduced(loss_fun, cloud, mask, kw)
            return (out[0] / out[1] if reduction == Reduction.MEAN else out[0]), loss_cloud
        # general path: raw per-point values from the kernel, tail in torch
        kind, flags = _kind_flags(loss_fun, dict(kw, sqrt=False))
        raw = fused_loss(cloud.step_state(), cloud._model, cloud.poses_tensor(), kind, flags | L.FLAG_RAW, mask=None)
        view = _RawView(cloud)
        if mask is not None:
            raw = raw[mask]
            view = view[mask]
        return _finish(view, raw, None, offset, sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult,
                       only_finite, skip_nans)
    # staged clouds: the reference's element-wise arithmetic on precomputed features
    if loss_fun is trace_loss:
        assert cloud.cov is not None
    else:
        assert cloud.eigvals is not None
    if mask is not None:
        cloud = cloud[mask]
    if loss_fun is trace_loss:
        loss = trace(cloud.cov)
    else:
        eigvals = cloud.eigvals
        loss = eigvals[:, 0]
        if normalization:
            loss = loss / eigvals.sum(dim=-1).clamp(min=1e-6)
    return _finish(cloud, loss, None, offset, sqrt, reduction, inlier_max_loss, inlier_ratio, inlier_loss_mult,
                   only_finite, skip_nans)


class _RawView(object):
    """Minimal stand-in for `cloud[mask]` on the general fused path: slicing composes index masks and
    `.copy()` yields an object that accepts `.loss`; feature fields are fetched from the lazy cloud on demand."""

    def __init__(self, cloud, index=None):
        self._cloud, self._index, self.loss = cloud, index, None

    def __getitem__(self, item):
        if self._index is None:
            idx = torch.arange(len(self._cloud), device=item.device)[item]
        else:
            idx = self._index[item]
        return _RawView(self._cloud, idx)

    def copy(self):
        return _RawView(self._cloud, self._index)

    def __getattr__(self, name):
        c = self.__dict__['_cloud']
        idx = self.__dict__['_index']
        x = getattr(c, name)
        if idx is not None and name in DepthCloud.sliced_fields and x is not None:
            return x[idx]
        return x


def min_eigval_loss(cloud, mask=None, offset=None, sqrt=False, normalization=False, reduction=Reduction.MEAN,
                    inlier_max_loss=None, inlier_ratio=1.0, inlier_loss_mult=1.0,
                    only_finite=False, skip_nans=False, **kwargs):
    """Map consistency loss based on the smallest eigenvalue (loss.py:216-294).

    :param cloud: DepthCloud (or list of clouds -> batch_loss).
    :param mask: Points used in the loss reduction.
    :param offset: Offset point-wise loss values, optional.
    :param sqrt: Whether to use square root of eigenvalue.
    :param normalization: Whether to normalize minimum eigenvalue by total variance.
    :return: (reduced loss, cloud with per-point `.loss`)
    """
    if isinstance(cloud, (list, tuple)):
        return batch_loss(min_eigval_loss, cloud, masks=mask, offsets=offset, sqrt=sqrt, normalization=normalization,
                          reduction=reduction, inlier_max_loss=inlier_max_loss, inlier_ratio=inlier_ratio,
                          inlier_loss_mult=inlier_loss_mult, only_finite=only_finite, skip_nans=skip_nans)
    return _generic_loss(min_eigval_loss, cloud, mask, offset, sqrt, normalization, reduction, inlier_max_loss,
                         inlier_ratio, inlier_loss_mult, only_finite, skip_nans)


def trace_loss(cloud, mask=None, offset=None, sqrt=None, reduction=Reduction.MEAN,
               inlier_max_loss=None, inlier_ratio=1.0, inlier_loss_mult=1.0,
               only_finite=False, skip_nans=False, **kwargs):
    """Map consistency loss based on the trace of covariance matrix (loss.py:297-370)."""
    if isinstance(cloud, (list, tuple)):
        return batch_loss(trace_loss, cloud, masks=mask, offsets=offset, sqrt=sqrt, reduction=reduction,
                          inlier_max_loss=inlier_max_loss, inlier_ratio=inlier_ratio,
                          inlier_loss_mult=inlier_loss_mult, only_finite=only_finite, skip_nans=skip_nans)
    return _generic_loss(trace_loss, cloud, mask, offset, sqrt, False, reduction, inlier_max_loss,
                         inlier_ratio, inlier_loss_mult, only_finite, skip_nans)


def loss_by_name(name):
    assert name in ('min_eigval_loss', 'trace_loss', 'icp_loss'), name
    if name == 'icp_loss':
        from .icp import icp_loss
        return icp_loss
    return globals()[name]


def create_loss(cfg):
    loss = loss_by_name(cfg.loss)

    def loss_fun(*args, **kwargs):
        return loss(*args, **kwargs, **cfg.loss_kwargs)

    return loss_fun
