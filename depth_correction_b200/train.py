"""Training driver: train() of the reference (train.py:46-327) without its ROS / tensorboard coupling
(SURVEY.md section 8(f) row 4).

Same loop: per-scan features once, one neighbourhood graph and one loss mask per sequence, then per iteration
train loss -> validation loss -> best-checkpoint rule -> optimiser step (first pose of every sequence frozen for
per-pose corrections) -> validation-pose optimiser step.  Checkpoints keep the reference's file names
(`%03i_%.6g_state_dict.pth`, `..._pose_deltas.pth`, `..._poses_upd.pth`, `best.yaml`).  Datasets must be passed in
(iterables of (cloud, pose)); the reference's dataset readers (dataset.py: ASL laser, KITTI-360, ...) are out of
scope.  Every loss evaluation runs the fused kernels; the optimiser is torch's (a handful of scalars).
"""
import os
import tempfile

import numpy as np
import torch
from torch.optim import Adam, SGD, LBFGS  # noqa: F401  (cfg.optimizer is evaluated by name, train.py:148)

from .config import Config, NeighborhoodType, PoseCorrection
from .depth_cloud import DepthCloud
from .eval import eval_loss_clouds, initialize_pose_corrections
from .icp import nanquantile
from .graph import search
from .loss import create_loss
from .model import load_model
from .preproc import (compute_neighborhood_features, establish_neighborhoods, global_cloud, global_cloud_mask, local_feature_clouds,
                      local_feature_cloud)

__all__ = ['TrainCallbacks', 'train']


class TrainCallbacks(object):
    """train.py:24-43."""

    def __init__(self, cfg=None):
        self.cfg = cfg

    def iteration_started(self, iter):
        pass

    def train_inputs(self, iter, clouds, poses):
        pass

    def val_inputs(self, iter, clouds, poses):
        pass

    def train_loss(self, iter, model, clouds, pose_deltas, poses, masks, loss):
        pass

    def val_loss(self, iter, model, clouds, pose_deltas, poses, masks, loss):
        pass


def _prepare(datasets, cfg):
    """Per-scan feature clouds and poses of every sequence (train.py:92-110)."""
    all_clouds, all_poses = [], []
    for ds in datasets:
        if cfg.nn_type != NeighborhoodType.ball:
            raise NotImplementedError('plane neighbourhoods are out of scope of the B200 hot path')
        clouds, poses = [], []
        for cloud, pose in ds:
            clouds.append(cloud)
            poses.append(np.asarray(pose.detach().cpu() if isinstance(pose, torch.Tensor) else pose))
        # the reference calls local_feature_cloud scan by scan (train.py:97-104); here all scans of a sequence go through
        # one stacked search and one neighbourhood pass
        clouds = local_feature_clouds(clouds, cfg)
        all_clouds.append(clouds)
        all_poses.append(torch.as_tensor(np.stack(poses).astype(np.float64), device=cfg.device))
    return all_clouds, all_poses


def _icp_masks(clouds, poses, ratio):
    """Correspondences between consecutive scans at the initial poses (train.py:181-210), searched on the device."""
    out = []
    for seq_clouds, seq_poses in zip(clouds, poses):
        seq = []
        for j in range(len(seq_clouds) - 1):
            p1 = seq_clouds[j].transform(seq_poses[j]).to_points().detach()
            p2 = seq_clouds[j + 1].transform(seq_poses[j + 1]).to_points().detach()
            g = search(p2, p1, k=1)
            dist = g.distances().reshape(-1)
            mask1 = dist <= nanquantile(dist, ratio)
            seq.append((mask1, g.neighbors().reshape(-1)[mask1]))
        out.append(seq)
    return out


def train(cfg, callbacks=None, train_datasets=None, val_datasets=None):
    """Train the depth correction model (and pose corrections), validate it, and return the best config."""
    if not callbacks:
        callbacks = TrainCallbacks(cfg)
    if not train_datasets or val_datasets is None:
        raise NotImplementedError('pass train_datasets / val_datasets (iterables of (cloud, pose)); the dataset readers of '
                                  'the reference are outside the hot path')
    from .fused import set_backward_form
    set_backward_form(getattr(cfg, 'backward_form', 'auto'))
    if not cfg.log_dir:
        cfg.log_dir = tempfile.mkdtemp(prefix='depth_correction_b200_')
    os.makedirs(cfg.log_dir, exist_ok=True)
    cfg_path = os.path.join(cfg.log_dir, 'train.yaml')
    if not os.path.exists(cfg_path):
        cfg.to_yaml(cfg_path)

    loss_fun = create_loss(cfg)
    train_clouds, train_poses = _prepare(train_datasets, cfg)
    train_pose_deltas = initialize_pose_corrections(train_datasets, cfg)
    val_clouds, val_poses = _prepare(val_datasets, cfg)
    if cfg.pose_correction == PoseCorrection.common:
        val_pose_deltas = len(val_datasets) * [train_pose_deltas[0]]      # reuse the correction from training
    else:
        val_pose_deltas = initialize_pose_corrections(val_datasets, cfg)

    model = load_model(cfg=cfg, eval_mode=False)
    params = []
    if cfg.optimize_model and len(list(model.parameters())) > 0:
        params.append({'params': model.parameters(), 'lr': cfg.lr})
    if cfg.pose_correction != PoseCorrection.none:
        params.append({'params': train_pose_deltas, 'lr': cfg.lr})
    args = cfg.optimizer_args[:] if cfg.optimizer_args else []
    kwargs = cfg.optimizer_kwargs.copy() if cfg.optimizer_kwargs else {}
    optimizer = eval(cfg.optimizer)(params, *args, **kwargs)
    val_optimizer = None
    if cfg.pose_correction in (PoseCorrection.sequence, PoseCorrection.pose) and val_datasets:
        val_optimizer = eval(cfg.optimizer)([{'params': val_pose_deltas, 'lr': cfg.lr}], *args, **kwargs)

    # neighbourhoods and masks of the initial global clouds: fixed over the optimisation (train.py:164-215)
    train_global = [global_cloud(clouds=c, poses=p) for c, p in zip(train_clouds, train_poses)]
    val_global = [global_cloud(clouds=c, poses=p) for c, p in zip(val_clouds, val_poses)]
    train_ns = [establish_neighborhoods(cloud=c, cfg=cfg) for c in train_global]
    val_ns = [establish_neighborhoods(cloud=c, cfg=cfg) for c in val_global]
    if cfg.loss == 'icp_loss':
        ratio = cfg.loss_kwargs['icp_inlier_ratio']
        train_masks, val_masks = _icp_masks(train_clouds, train_poses, ratio), _icp_masks(val_clouds, val_poses, ratio)
    else:
        def feature_mask(cloud, ns):
            # the reference's establish_neighborhoods leaves the features on the initial global cloud
            # (preproc.py:180-185); here the fixed graph is attached and the mask statistics read them lazily
            cloud = compute_neighborhood_features(cloud=cloud, neighborhoods=ns, cfg=cfg)
            return global_cloud_mask(cloud, cloud.mask if hasattr(cloud, 'mask') else None, cfg)
        train_masks = [feature_mask(c, ns) for c, ns in zip(train_global, train_ns)]
        val_masks = [feature_mask(c, ns) for c, ns in zip(val_global, val_ns)]

    min_train_loss = np.inf
    min_val_loss = np.inf
    best_cfg = None
    history = []
    for it in range(cfg.n_opt_iters):
        callbacks.iteration_started(it)
        train_loss, _, train_poses_upd, train_feat = eval_loss_clouds(train_clouds, train_poses, train_pose_deltas, train_masks,
                                                                      train_ns, model, loss_fun, cfg)
        callbacks.train_loss(it, model, train_feat, train_pose_deltas, train_poses_upd, train_masks, train_loss)
        if val_datasets:
            val_loss, _, val_poses_upd, val_feat = eval_loss_clouds(val_clouds, val_poses, val_pose_deltas, val_masks, val_ns,
                                                                    model, loss_fun, cfg)
            callbacks.val_loss(it, model, val_feat, val_pose_deltas, val_poses_upd, val_masks, val_loss)
        else:
            val_loss = train_loss.detach()
        tl, vl = train_loss.item(), val_loss.item()
        history.append((tl, vl))
        if tl < min_train_loss and vl < min_val_loss:
            # (the reference never updates min_train_loss, train.py:242-244: kept)
            saved = True
            min_val_loss = vl
            state_dict_path = '%s/%03i_%.6g_state_dict.pth' % (cfg.log_dir, it, min_val_loss)
            torch.save(model.state_dict(), state_dict_path)
            pose_deltas_path = '%s/%03i_%.6g_pose_deltas.pth' % (cfg.log_dir, it, min_val_loss)
            torch.save([p.detach().clone() for p in train_pose_deltas if p is not None], pose_deltas_path)
            poses_upd_path = '%s/%03i_%.6g_poses_upd.pth' % (cfg.log_dir, it, min_val_loss)
            torch.save([torch.stack(list(p)).detach().clone() if isinstance(p, (list, tuple)) else p.detach().clone()
                        for p in train_poses_upd if p is not None], poses_upd_path)
            best_cfg = cfg.copy()
            best_cfg.model_state_dict = state_dict_path
            best_cfg.train_pose_deltas = pose_deltas_path
            best_cfg.to_yaml(os.path.join(cfg.log_dir, 'best.yaml'))
        else:
            saved = False
        print('It. %03i: train loss: %.9f, val.: %.9f. Model %s %s.' % (it, tl, vl, model, 'saved' if saved else 'not saved'))

        if cfg.optimizer == 'LBFGS':
            def closure():
                optimizer.zero_grad()
                train_loss.backward(retain_graph=True)
                return train_loss
        else:
            optimizer.zero_grad()
            train_loss.backward()
        if cfg.pose_correction == PoseCorrection.pose:                      # keep the first pose fixed
            for d in train_pose_deltas:
                if d.grad is not None:
                    d.grad[0].zero_()
        optimizer.step(closure) if cfg.optimizer == 'LBFGS' else optimizer.step()
        if val_optimizer is not None:
            val_optimizer.zero_grad()
            val_loss.backward()
            if cfg.pose_correction == PoseCorrection.pose:
                for d in val_pose_deltas:
                    if d.grad is not None:
                        d.grad[0].zero_()
            val_optimizer.step()
    if best_cfg is not None:
        best_cfg.loss_history = [list(h) for h in history]
    return best_cfg
