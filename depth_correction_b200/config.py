"""The handful of configuration fields the hot path reads (config.py:143-292 of the reference).

The reference's Config is a ~70-field YAML/argparse/rosparam object; re-implementing that system is
out of scope.  This Config carries the same attribute names and
defaults for the fields the map-consistency path uses, and any reference Config (or a
SimpleNamespace) with these attributes works in its place.
"""
import numpy as np
import torch

__all__ = ['Config', 'Loss', 'Model', 'NeighborhoodType', 'PoseCorrection']


class _ValueEnum(type):
    def __iter__(cls):
        return iter(v for k, v in vars(cls).items() if not k.startswith('_'))

    def __contains__(cls, item):
        return item in list(iter(cls))


class NeighborhoodType(metaclass=_ValueEnum):
    ball = 'ball'
    plane = 'plane'


class Loss(metaclass=_ValueEnum):
    min_eigval_loss = 'min_eigval_loss'
    trace_loss = 'trace_loss'
    icp_loss = 'icp_loss'


class Model(metaclass=_ValueEnum):
    Polynomial = 'Polynomial'
    ScaledPolynomial = 'ScaledPolynomial'


class PoseCorrection(metaclass=_ValueEnum):
    none = 'none'
    common = 'common'
    sequence = 'sequence'
    pose = 'pose'


class Config(object):
    def __init__(self, **kwargs):
        self.random_seed = 135
        self.float_type = 'float32'          # the B200 path stores float32 (reference default: float64)
        self.device = 'cuda'
        self.min_depth = 5.0
        self.max_depth = 25.0
        self.grid_res = 0.2
        self.nn_type = NeighborhoodType.ball
        self.nn_k = 0
        self.nn_r = 0.25
        self.nn_scale = None
        self.min_valid_neighbors = 5
        self.shadow_neighborhood_angle = 0.017453
        self.shadow_angle_bounds = []
        self.dir_dispersion_bounds = []
        self.vp_dispersion_bounds = [0.36, float('inf')]
        self.vp_dispersion_to_depth2_bounds = []
        self.eigenvalue_bounds = []
        self.eigenvalue_ratio_bounds = [[0, 1, 0, 0.25], [1, 2, 0.25, 1.]]
        self.log_filters = False
        self.model_class = Model.ScaledPolynomial
        self.model_args = []
        self.model_kwargs = {'w': [0.0], 'exponent': [4.0]}
        self.model_state_dict = ''
        self.loss = Loss.min_eigval_loss
        self.loss_offset = False
        self.loss_kwargs = {'sqrt': False, 'normalization': True, 'inlier_max_loss': None,
                            'inlier_loss_mult': 1.0, 'inlier_ratio': 1.0}
        self.pose_correction = PoseCorrection.none
        self.optimizer = 'Adam'
        self.optimizer_args = []
        self.optimizer_kwargs = {}
        self.lr = 1e-4
        self.n_opt_iters = 100
        self.optimize_model = True
        self.log_dir = None                  # train() creates a temporary directory when unset
        # how dL/dp is accumulated by the fused step: 'auto' (float32 L2 reductions on maps of >= 2^19 points, fast, not
        # bitwise reproducible), 'gather' (fp64, atomic-free: deterministic like the reference's autograd), 'scatter'
        self.backward_form = 'auto'
        self.train_names = []
        self.val_names = []
        for k, v in kwargs.items():
            setattr(self, k, v)

    def numpy_float_type(self):
        return getattr(np, self.float_type)

    def torch_float_type(self):
        return getattr(torch, self.float_type)

    def to_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, (bool, int, float, str, list, dict)) or v is None:
                out[k] = v
        return out

    def to_yaml(self, path=None):
        """config.py:270-283: plain-value fields as YAML (to a file when a path is given)."""
        import yaml
        text = yaml.safe_dump(self.to_dict())
        if path is None:
            return text
        with open(path, 'w') as f:
            f.write(text)

    def from_yaml(self, path):
        import yaml
        with open(path) as f:
            for k, v in (yaml.safe_load(f) or {}).items():
                setattr(self, k, v)
        return self

    def copy(self):
        c = Config()
        c.__dict__.update({k: (v.copy() if isinstance(v, (dict, list)) else v) for k, v in self.__dict__.items()})
        return c
