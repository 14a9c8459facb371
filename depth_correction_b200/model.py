"""Depth-correction models named by the north star: Polynomial and ScaledPolynomial
(model.py:149-286 of the reference), same constructor, parameters and state dict
({'w': float64 [1,n]} + 'exponent' if learnable), so reference checkpoints load unchanged.

`model(cloud)` / `correct_depth` keep the reference's staged semantics (elementwise, returns a
shallow copy with a new `depth`); inside the training loop the same arithmetic runs fused in
kernel 2 (pass A) and its gradient in kernel 3 -- see fused.py.
"""
import numpy as np
import torch

from .depth_cloud import DepthCloud

__all__ = ['BaseModel', 'load_model', 'model_by_name', 'Polynomial', 'ScaledPolynomial']


class BaseModel(torch.nn.Module):

    def __init__(self, device=torch.device('cpu')):
        super(BaseModel, self).__init__()
        self.device = device

    def forward(self, dc: DepthCloud) -> DepthCloud:
        return self.correct_depth(dc, dc.mask)

    def correct_depth(self, dc: DepthCloud, mask=None) -> DepthCloud:
        return dc

    def inverse(self, dc: DepthCloud, mask=None) -> DepthCloud:
        return dc

    def __str__(self):
        return 'BaseModel()'

    def construct(self, *args, **kwargs):
        return type(self)(*args, **kwargs)

    def detach(self):
        return self.construct(**{k: v.detach() for k, v in self.named_parameters()})

    def clone(self):
        return self.construct(**{k: v.clone() for k, v in self.named_parameters()})


class _PolynomialBase(BaseModel):
    scaled = False

    def __init__(self, p0=None, p1=None, w=None, exponent=None, learnable_exponents=False,
                 device=torch.device('cpu')):
        super().__init__(device=device)
        if exponent is None:
            assert w is None, w
            self.legacy = True
            exponent = [2.0, 4.0]
            w = [p0 or 0.0, p1 or 0.0]
        else:
            self.legacy = False
        if w is None:
            w = [0.0] * len(exponent)
        elif isinstance(w, float):
            w = [w]
        w = torch.as_tensor(w, dtype=torch.float64, device=device).view((1, -1))
        exponent = torch.as_tensor(exponent, dtype=torch.float64, device=device).view((1, -1))
        assert w.numel() == exponent.numel(), (w, exponent)
        self.w = torch.nn.Parameter(w)
        self.exponent = torch.nn.Parameter(exponent) if learnable_exponents else exponent

    def bias(self, inc_angles):
        assert inc_angles.dim() == 2
        assert inc_angles.shape[1] == 1
        x = torch.pow(inc_angles, self.exponent.to(inc_angles.device))
        return torch.matmul(x, self.w.t().to(x.dtype)).view((-1, 1))

    def _apply_bias(self, depth, bias, inverse):
        if self.scaled:
            return depth / (1. - bias) if inverse else depth * (1. - bias)
        return depth + bias if inverse else depth - bias

    def _correct(self, dc, mask, inverse):
        assert dc.inc_angles is not None
        dc_corr = dc.copy()
        if mask is None:
            bias = self.bias(dc.inc_angles).to(dc.depth.dtype)
            if inverse and not self.scaled:
                # the reference's unmasked Polynomial.inverse divides (model.py:210), kept for parity
                dc_corr.depth = dc_corr.depth / (1. - bias)
            else:
                dc_corr.depth = self._apply_bias(dc_corr.depth, bias, inverse)
        else:
            bias = self.bias(dc.inc_angles[mask]).to(dc.depth.dtype)
            depth = dc_corr.depth.clone()       # avoid modifying depth in-place
            depth[mask] = self._apply_bias(depth[mask], bias, inverse)
            dc_corr.depth = depth
        return dc_corr

    def correct_depth(self, dc: DepthCloud, mask=None) -> DepthCloud:
        return self._correct(dc, mask, inverse=False)

    def inverse(self, dc: DepthCloud, mask=None) -> DepthCloud:
        return self._correct(dc, mask, inverse=True)

    def to(self, *args, **kwargs):
        ret = super().to(*args, **kwargs)
        if not isinstance(ret.exponent, torch.nn.Parameter):
            ret.exponent = ret.exponent.to(*args, **kwargs)
        return ret

    def __str__(self):
        return '%s(%s)' % (type(self).__name__, ', '.join(
            '%.6gx^%.6g' % (w, e) for w, e in zip(self.w.detach().flatten().tolist(), self.exponent.detach().flatten().tolist())))


class Polynomial(_PolynomialBase):
    """d' = d - sum_k w_k gamma^e_k (model.py:149-215)."""
    scaled = False


class ScaledPolynomial(_PolynomialBase):
    """d' = d (1 - sum_k w_k gamma^e_k) (model.py:218-286)."""
    scaled = True


def model_by_name(name):
    assert name in ('BaseModel', 'Polynomial', 'ScaledPolynomial'), name
    return globals()[name]


def load_model(class_name=None, model_args=None, model_kwargs=None, state_dict=None, device=None, cfg=None,
               eval_mode=True):
    """model.py:19-67."""
    if cfg is not None:
        class_name = class_name if class_name is not None else cfg.model_class
        model_args = model_args if model_args is not None else (cfg.model_args[:] if cfg.model_args else [])
        model_kwargs = model_kwargs if model_kwargs is not None else (cfg.model_kwargs.copy() if cfg.model_kwargs else {})
        state_dict = state_dict if state_dict is not None else cfg.model_state_dict
        device = device if device is not None else cfg.device
    model_args = model_args or []
    model_kwargs = model_kwargs or {}
    if isinstance(state_dict, str) and state_dict:
        state_dict = torch.load(state_dict)
    if isinstance(device, str):
        device = torch.device(device)
    if 'device' not in model_kwargs:
        model_kwargs['device'] = device
    model = model_by_name(class_name)(*model_args, **model_kwargs)
    assert isinstance(model, BaseModel)
    if state_dict:
        model.load_state_dict(state_dict)
    if eval_mode:
        model.eval()
    model.to(device)
    return model
