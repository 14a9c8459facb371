"""SE(3) pose parametrisation (transform.py:68-91 of the reference; pytorch3d is not required)."""
import torch

__all__ = ['axis_angle_to_matrix', 'matrix_to_xyz_axis_angle', 'xyz_axis_angle_to_matrix']


def axis_angle_to_matrix(axis_angle):
    """Axis-angle -> quaternion -> rotation matrix, the map pytorch3d implements (transform.py:73).
    Small differentiable torch version for arbitrary leading dims; the per-scan training path uses
    the dc_pose_compose kernel (ops.pose_compose) instead."""
    angles = torch.norm(axis_angle, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angles), angles)
    s = torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / safe)
    q = torch.cat([torch.cos(half), axis_angle * s], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def xyz_axis_angle_to_matrix(xyz_axis_angle):
    assert isinstance(xyz_axis_angle, torch.Tensor)
    assert xyz_axis_angle.shape[-1] == 6
    mat = torch.zeros(xyz_axis_angle.shape[:-1] + (4, 4), dtype=xyz_axis_angle.dtype, device=xyz_axis_angle.device)
    mat[..., :3, :3] = axis_angle_to_matrix(xyz_axis_angle[..., 3:])
    mat[..., :3, 3] = xyz_axis_angle[..., :3]
    mat[..., 3, 3] = 1.
    return mat


def matrix_to_xyz_axis_angle(T):
    """Inverse of xyz_axis_angle_to_matrix for [n,4,4] poses (transform.py:81-91)."""
    assert isinstance(T, torch.Tensor)
    assert T.dim() == 3 and T.shape[1:] == (4, 4)
    R = T[:, :3, :3]
    cos = ((R.diagonal(dim1=-2, dim2=-1).sum(-1) - 1.0) / 2.0).clamp(-1.0, 1.0)
    angle = torch.arccos(cos)
    axis = torch.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], dim=-1)
    sin = torch.sin(angle)
    scale = torch.where(sin.abs() < 1e-9, torch.full_like(sin, 0.5), angle / (2.0 * sin.clamp(min=1e-300)))
    return torch.cat([T[:, :3, 3], axis * scale[:, None]], dim=1)
