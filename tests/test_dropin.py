"""Drop-in acceptance test of the boundary: the optimisation loop
of the reference's own script, VERBATIM (tests/golden/dropin_loop.txt = scripts/model_poses_learning:121-135, pinned to
the reference text by tests/test_oracle.py), runs unchanged with `depth_correction` aliased to `depth_correction_b200`,
and its loss trajectory is the CPU oracle's under the same optimiser."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _alias_reference_package():
    """The recipe of INTEGRATION.md section A."""
    import depth_correction_b200
    sys.modules['depth_correction'] = depth_correction_b200
    for m in ('depth_cloud', 'model', 'loss', 'preproc', 'eval', 'filters', 'nearest_neighbors', 'transform', 'utils', 'config'):
        sys.modules['depth_correction.' + m] = getattr(__import__('depth_correction_b200.' + m), m)


def test_reference_loop_runs_unchanged_and_follows_the_oracle():
    from oracle import oracle
    from depth_correction_b200.synthetic import make_sequence
    _alias_reference_package()
    try:
        # the script's own imports (scripts/model_poses_learning:8-16), resolved through the alias
        from depth_correction.depth_cloud import DepthCloud
        from depth_correction.model import ScaledPolynomial
        from depth_correction.preproc import local_feature_cloud
        from depth_correction.config import Config, PoseCorrection, Loss, NeighborhoodType
        from depth_correction.loss import create_loss
        from depth_correction.eval import create_corrected_poses, global_cloud
        from depth_correction.preproc import establish_neighborhoods, compute_neighborhood_features

        # its configuration (:58-74), on the GPU and with a smaller synthetic sequence
        cfg = Config()
        cfg.grid_res = 0.0
        cfg.min_depth = 0.0
        cfg.nn_r = 0.4
        cfg.lr = 0.001
        cfg.n_opt_iters = 5
        cfg.device = 'cuda'
        cfg.float_type = 'float64'
        cfg.pose_correction = PoseCorrection.pose
        cfg.loss = Loss.trace_loss
        cfg.nn_type = NeighborhoodType.ball
        scans_np, _, poses = make_sequence('corridor', n_scans=4, pattern='os0-32', seed=17, grid_res=0.15, step=0.9,
                                           pose_noise=(0.01, 0.005), bias_w=[-0.01], bias_exponent=[4.0], depth_clip=(1.0, 8.0))
        train_clouds = [local_feature_cloud(cloud=DepthCloud.from_points(torch.as_tensor(s['points'].astype(np.float64), device=cfg.device)), cfg=cfg)
                        for s in scans_np]
        train_poses = torch.as_tensor(poses, device=cfg.device, dtype=cfg.torch_float_type())
        train_pose_deltas = torch.zeros((len(train_poses), 6), dtype=cfg.torch_float_type(), requires_grad=True, device=cfg.device)
        train_ns = establish_neighborhoods(clouds=train_clouds, poses=train_poses, cfg=cfg)          # :93
        model = ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=cfg.device)                    # :95
        loss_fn = create_loss(cfg)                                                                     # :98
        optimizer = torch.optim.Adam([{'params': train_pose_deltas, 'lr': cfg.lr},                     # :101-102
                                      {'params': model.parameters(), 'lr': cfg.lr}], **cfg.optimizer_kwargs)
        losses_train = []
        loop = compile(open(os.path.join(HERE, 'golden', 'dropin_loop.txt')).read(), 'scripts/model_poses_learning:121-135', 'exec')
        scope = dict(globals(), **locals())
        for it in range(cfg.n_opt_iters):
            exec(loop, scope)
        losses_train = scope['losses_train']
    finally:
        for name in [m for m in sys.modules if m == 'depth_correction' or m.startswith('depth_correction.')]:
            del sys.modules[name]
    assert len(losses_train) == 5 and losses_train[-1] < losses_train[0]

    # oracle: same per-scan records and graph, same Adam (parameter groups in the script's order)
    scans = [{'vps': c.vps.double().cpu().expand(len(c), 3), 'dirs': c.dirs.double().cpu(), 'depth': c.depth.double().cpu(),
              'inc_angles': c.inc_angles.double().cpu(), 'mask': c.mask.cpu()} for c in train_clouds]
    poses_t = torch.as_tensor(poses)
    pts0, _ = oracle.global_points(scans, poses_t)
    _, nb = oracle.nearest_neighbors(pts0, r=0.4)
    assert torch.equal(train_ns[0].cpu(), nb)
    w = torch.zeros((1, 2), dtype=torch.float64, requires_grad=True)
    deltas = torch.zeros((len(scans), 6), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([{'params': [deltas], 'lr': 1e-3}, {'params': [w], 'lr': 1e-3}])
    ref = []
    for it in range(5):
        out = oracle.map_consistency_step(scans, poses_t, nb, w.detach(), torch.tensor([[2.0, 4.0]], dtype=torch.float64),
                                          pose_deltas=deltas.detach(), loss='trace_loss', sqrt=bool(cfg.loss_kwargs.get('sqrt')))
        ref.append(float(out['loss']))
        opt.zero_grad()
        w.grad = out['w_grad'].reshape(1, 2).clone()
        deltas.grad = out['pose_deltas_grad'].clone()
        opt.step()
    assert np.allclose(losses_train, ref, rtol=1e-6), (losses_train, ref)
