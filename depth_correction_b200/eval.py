"""Pose-correction helpers and one evaluation of the hot path (eval.py:31-112 of the reference)."""
import torch

from .config import NeighborhoodType, PoseCorrection
from .depth_cloud import DepthCloud
from . import ops
from .preproc import compute_neighborhood_features, global_cloud, global_cloud_mask, local_feature_cloud, offset_cloud
from .transform import xyz_axis_angle_to_matrix

__all__ = ['create_corrected_poses', 'eval_loss_clouds', 'initialize_pose_corrections']


def initialize_pose_corrections(datasets, cfg):
    """Zero-initialised pose corrections per sequence (eval.py:31-65)."""
    pose_deltas = []
    kwargs = {'dtype': torch.float64, 'device': cfg.device, 'requires_grad': True}
    for ds in datasets:
        if cfg.pose_correction == PoseCorrection.common:
            pose_delta = pose_deltas[0] if pose_deltas else torch.zeros((1, 6), **kwargs)
        elif cfg.pose_correction == PoseCorrection.sequence:
            pose_delta = torch.zeros((1, 6), **kwargs)
        elif cfg.pose_correction == PoseCorrection.pose:
            pose_delta = torch.zeros((len(ds), 6), **kwargs)
        else:
            pose_delta = None
        pose_deltas.append(pose_delta)
    return pose_deltas


def _compose(poses, deltas):
    if isinstance(poses, (list, tuple)):
        poses = torch.stack(list(poses))
    if poses.is_cuda and poses.dim() == 3 and deltas.dim() == 2 and deltas.shape[0] in (1, poses.shape[0]):
        return ops.pose_compose(poses, deltas.to(poses.device))       # dc_pose_compose kernel (+ backward)
    return torch.matmul(poses, xyz_axis_angle_to_matrix(deltas).to(poses.dtype))


def create_corrected_poses(poses, pose_deltas, cfg):
    """poses[i] @ xyz_axis_angle_to_matrix(pose_deltas[i]) (eval.py:68-82).

    Accepts what the reference's callers pass: a tensor [S,4,4] with deltas [S,6] (scripts/model_poses_learning:121)
    or lists over sequences of [S_i,4,4] poses and [S_i,6] / [1,6] deltas (train.py:225)."""
    if cfg.pose_correction == PoseCorrection.none:
        return poses
    assert len(poses) == len(pose_deltas)
    if isinstance(poses, torch.Tensor) and isinstance(pose_deltas, torch.Tensor) and poses.dim() == 3:
        return list(_compose(poses, pose_deltas).unbind(0))
    if cfg.pose_correction == PoseCorrection.common:
        assert all(d is pose_deltas[0] for d in pose_deltas[1:])
    poses_upd = []
    for i in range(len(poses)):
        d = pose_deltas[i]
        if d.dim() == 1:
            poses_upd.append(torch.matmul(poses[i], xyz_axis_angle_to_matrix(d).to(poses[i].dtype)))
        else:
            poses_upd.append(_compose(poses[i], d))
    return poses_upd


def eval_loss_clouds(clouds, poses, pose_deltas, masks, ns, model, loss_fun, cfg):
    """Evaluate loss on given clouds, poses, deltas, etc. (eval.py:85-112, ball neighbourhoods)."""
    if cfg.nn_type != NeighborhoodType.ball:
        raise NotImplementedError('plane neighbourhoods are out of scope of the B200 hot path')
    offsets = [offset_cloud(c, model) for c in clouds] if cfg.loss_offset else None
    poses_upd = create_corrected_poses(poses, pose_deltas, cfg)
    global_clouds = [global_cloud(clouds=c, model=model, poses=p) for c, p in zip(clouds, poses_upd)]
    feat_clouds = [compute_neighborhood_features(cloud=cloud, model=None, neighborhoods=nn, cfg=cfg)
                   for cloud, nn in zip(global_clouds, ns)]
    if cfg.loss == 'icp_loss':
        if clouds[0][0].normals is None:
            clouds = [[local_feature_cloud(cloud, cfg) for cloud in seq_clouds] for seq_clouds in clouds]
        loss, loss_cloud = loss_fun(clouds, poses_upd, model, masks=masks)
    else:
        if (not masks or masks[0] is None) and isinstance(feat_clouds[0], DepthCloud):
            masks = [global_cloud_mask(cloud, cloud.mask if hasattr(cloud, 'mask') else None, cfg) for cloud in feat_clouds]
        loss, loss_cloud = loss_fun(feat_clouds, mask=masks, offset=offsets)
    return loss, loss_cloud, poses_upd, feat_clouds
