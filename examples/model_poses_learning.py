#!/usr/bin/env python
"""Joint learning of the depth-correction model and per-scan SE(3) pose corrections with trace_loss
(BASELINE.json configs[3]; the loop of the reference's scripts/model_poses_learning:93-135, unchanged apart from the
import and the synthetic FEE-corridor-shaped data in place of the dataset reader).

    python examples/model_poses_learning.py [--scans 12] [--iters 200] [--loss trace_loss]

Prints the loss, the learned weights and the mean position error of the corrected poses against the ground truth
(the reference plots the same three curves).  Needs a B200 (there is no CPU fallback).
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from depth_correction_b200.config import Config, PoseCorrection                       # noqa: E402
from depth_correction_b200.depth_cloud import DepthCloud                              # noqa: E402
from depth_correction_b200.eval import create_corrected_poses                         # noqa: E402
from depth_correction_b200.loss import create_loss                                    # noqa: E402
from depth_correction_b200.model import ScaledPolynomial                              # noqa: E402
from depth_correction_b200.preproc import (compute_neighborhood_features, establish_neighborhoods, filtered_cloud,   # noqa: E402
                                           global_cloud, local_feature_cloud)
from depth_correction_b200.synthetic import make_sequence                             # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scans', type=int, default=12)
    ap.add_argument('--iters', type=int, default=200)
    ap.add_argument('--loss', default='trace_loss', choices=['trace_loss', 'min_eigval_loss'])
    ap.add_argument('--lr', type=float, default=5e-4)
    args = ap.parse_args()

    cfg = Config(min_depth=1.0, max_depth=25.0, grid_res=0.1, nn_k=0, nn_r=0.25, loss=args.loss, lr=args.lr,
                 loss_kwargs={'sqrt': False} if args.loss == 'trace_loss' else {'sqrt': False, 'normalization': True},
                 pose_correction=PoseCorrection.pose, n_opt_iters=args.iters, device='cuda')
    # FEE-corridor-shaped data: noisy initial poses (first pose exact), depth bias injected with the model's inverse
    scans, poses_gt, poses_init = make_sequence('fee', n_scans=args.scans, pattern='os0-128', seed=5, rings=64, azimuths=512,
                                                pose_noise=(0.01, 0.005), bias_w=[-0.01], bias_exponent=[4.0])
    train_clouds = []
    for s in scans:
        cloud = DepthCloud.from_points(torch.as_tensor(s['points'], device=cfg.device))
        cloud = filtered_cloud(cloud, cfg)                    # depth + voxel-grid filters (preproc.py:25-32)
        train_clouds.append(local_feature_cloud(cloud, cfg))  # normals, incidence angles, planarity mask
    train_poses = torch.as_tensor(poses_init, device=cfg.device)
    gt_xyz = torch.as_tensor(poses_gt[:, :3, 3], device=cfg.device)
    train_pose_deltas = torch.zeros((len(train_poses), 6), dtype=torch.float64, requires_grad=True, device=cfg.device)

    train_ns = establish_neighborhoods(clouds=train_clouds, poses=train_poses, cfg=cfg)
    model = ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=cfg.device)
    loss_fn = create_loss(cfg)
    optimizer = torch.optim.Adam([{'params': train_pose_deltas, 'lr': cfg.lr}, {'params': model.parameters(), 'lr': cfg.lr}])

    n = sum(len(c) for c in train_clouds)
    print('%d scans, %d points after filtering, %s' % (len(train_clouds), n, args.loss))
    for it in range(cfg.n_opt_iters):
        train_poses_corr = torch.stack(create_corrected_poses(train_poses, train_pose_deltas, cfg))
        cloud = global_cloud(clouds=train_clouds, model=model, poses=train_poses_corr)
        feats = compute_neighborhood_features(cloud=cloud, model=None, neighborhoods=train_ns, cfg=cfg)
        loss_train, _ = loss_fn(feats)
        optimizer.zero_grad()
        loss_train.backward()
        train_pose_deltas.grad[0].zero_()                     # keep the first pose fixed (train.py:306-309)
        optimizer.step()
        if it % max(cfg.n_opt_iters // 10, 1) == 0 or it == cfg.n_opt_iters - 1:
            with torch.no_grad():
                pose_err = torch.linalg.norm(gt_xyz - train_poses_corr[:, :3, 3], dim=1).mean().item()
            print('it %4d  loss %.9f  model %s  mean position error %.4f m' % (it, loss_train.item(), model, pose_err))


if __name__ == '__main__':
    main()
