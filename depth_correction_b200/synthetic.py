"""Seeded synthetic lidar sequences shaped like the reference's datasets.

No data ships with the reference (data/ only holds a download script), and its synthetic
generators (dataset.py:39-414) need pytorch3d/open3d; these are our own seeded generators with
the same role: analytically planar scenes scanned by a spinning lidar moving along a path.
Everything is produced as float32 (the storage type of the B200 path); the oracle consumes the
same float32 values up-cast to float64.

scenes
  corridor : 3 m x 3 m box cross-section, unbounded along x, sensor 1 m above the floor,
             poses every `step` metres along x with a small yaw (BASELINE.json configs 0/1)
  street   : ground plane + two facades 12 m apart (KITTI-360-shaped, config 2)
  fee      : the corridor with a side room behind a doorway and a flight of stairs up to a landing (FEE-corridor
             shaped, config 3); used with noisy initial poses (first pose exact) for pose learning
patterns
  os0-128  : 128 rings x 1024 azimuths, elevation +-45 deg
  hdl-64   : 64 rings (+2 ... -24.8 deg) x 1900 azimuths
"""
import numpy as np

__all__ = ['beam_pattern', 'make_sequence', 'scene_planes', 'voxel_keep_first']

PATTERNS = {
    'os0-128': dict(rings=128, azimuths=1024, elev=(-45.0, 45.0)),
    'os0-32': dict(rings=32, azimuths=256, elev=(-45.0, 45.0)),     # small test pattern
    'hdl-64': dict(rings=64, azimuths=1900, elev=(-24.8, 2.0)),
}


def beam_pattern(name, rings=None, azimuths=None):
    p = dict(PATTERNS[name])
    if rings:
        p['rings'] = rings
    if azimuths:
        p['azimuths'] = azimuths
    el = np.deg2rad(np.linspace(p['elev'][0], p['elev'][1], p['rings']))
    az = np.linspace(-np.pi, np.pi, p['azimuths'], endpoint=False)
    el, az = np.meshgrid(el, az, indexing='ij')
    return el.ravel(), az.ravel()


INF = float('inf')


def scene_planes(scene):
    """Surfaces as (normal[3], offset, lo[3], hi[3]): the part of the plane n.x = offset inside the axis-aligned box
    [lo, hi] (infinite bounds = the whole plane); rays hit the nearest surface in front."""
    whole = ((-INF, -INF, -INF), (INF, INF, INF))
    if scene == 'corridor':
        return [((0, 0, 1), 0.0) + whole, ((0, 0, 1), 3.0) + whole, ((0, 1, 0), -1.5) + whole, ((0, 1, 0), 1.5) + whole]
    if scene == 'street':
        return [((0, 0, 1), 0.0) + whole, ((0, 1, 0), -6.0) + whole, ((0, 1, 0), 6.0) + whole]
    if scene == 'fee':
        # corridor 3 m x 3 m along x; side room x in [6, 12], y in [1.5, 5.5] behind a doorway x in [8, 10];
        # stairs from x = 20: eight steps of 0.5 m x 0.15 m up to a landing at z = 1.2 m (x >= 24)
        srf = [((0, 0, 1), 3.0) + whole,                                              # ceiling
               ((0, 1, 0), -1.5) + whole,                                             # right wall
               ((0, 0, 1), 0.0, (-INF, -INF, -INF), (20.0, INF, INF)),                # floor up to the stairs
               ((0, 1, 0), 1.5, (-INF, -INF, -INF), (8.0, INF, INF)),                 # left wall before the door
               ((0, 1, 0), 1.5, (10.0, -INF, -INF), (INF, INF, INF)),                 # left wall after the door
               ((0, 1, 0), 5.5, (6.0, -INF, -INF), (12.0, INF, INF)),                 # side room: back wall
               ((1, 0, 0), 6.0, (-INF, 1.5, -INF), (INF, 5.5, INF)),                  # side room: side walls
               ((1, 0, 0), 12.0, (-INF, 1.5, -INF), (INF, 5.5, INF)),
               ((0, 0, 1), 1.2, (24.0, -INF, -INF), (INF, INF, INF))]                 # landing
        for i in range(8):
            x0 = 20.0 + 0.5 * i
            srf.append(((1, 0, 0), x0, (-INF, -INF, 0.15 * i), (INF, INF, 0.15 * (i + 1))))          # riser
            if i < 7:
                srf.append(((0, 0, 1), 0.15 * (i + 1), (x0, -INF, -INF), (x0 + 0.5, INF, INF)))      # tread
        return srf
    raise ValueError(scene)


def _yaw_pose(x, y, z, yaw):
    T = np.eye(4)
    c, s = np.cos(yaw), np.sin(yaw)
    T[:3, :3] = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    T[:3, 3] = [x, y, z]
    return T


def scan_pose(scene, sid, step=1.0):
    """Ground-truth sensor-to-world pose of scan `sid` (a function of the scan id only, so that every rank of
    a multi-GPU run can rebuild the poses of all scans without generating their points)."""
    sensor_z = 1.0 if scene != 'street' else 1.73
    yaw = 0.05 * np.sin(0.7 * sid)
    y0 = 0.2 * np.sin(0.3 * sid) if scene != 'street' else 1.5 * np.sin(0.02 * sid)
    return _yaw_pose(step * sid, y0, sensor_z, yaw)


def make_poses(scene, n_scans, step=1.0, first_scan=0):
    return np.stack([scan_pose(scene, first_scan + k, step) for k in range(n_scans)])


def voxel_keep_first(points, grid_res):
    """Keep the first point of every occupied voxel (generator-side density control; plays the
    role of the reference's filter_grid, filters.py:24-82, but deterministic)."""
    keys = np.floor(points / grid_res).astype(np.int64)
    keys -= keys.min(axis=0)
    dims = keys.max(axis=0) + 1
    lin = (keys[:, 0] * dims[1] + keys[:, 1]) * dims[2] + keys[:, 2]
    _, first = np.unique(lin, return_index=True)
    return np.sort(first)


def make_sequence(scene='corridor', n_scans=10, pattern='os0-128', seed=0, step=1.0,
                  range_noise=0.005, angle_jitter=1e-4, depth_clip=(1.0, 25.0), grid_res=0.0,
                  rings=None, azimuths=None, pose_noise=(0.0, 0.0), bias_w=None, bias_exponent=None,
                  first_scan=0):
    """Returns (scans, poses_gt, poses_init).

    scans: list of dict(points float32 [n,3] in the sensor frame, vps float32 [n,3] (zeros));
    poses_*: float64 [S,4,4] sensor-to-world.  `pose_noise=(sigma_xyz, sigma_rot)` perturbs
    poses_init (first pose exact).  `bias_w/bias_exponent` inject a ScaledPolynomial bias
    through the model's inverse formula d / (1 - sum w g^e) (model.py:263-274) using the
    analytic incidence angle.  `first_scan` offsets scan ids (per-rank generation of a shard).
    """
    el0, az0 = beam_pattern(pattern, rings, azimuths)
    planes = scene_planes(scene)
    scans, poses_gt, poses_init = [], [], []
    for k in range(n_scans):
        sid = first_scan + k
        rng = np.random.default_rng([seed, sid])
        T = scan_pose(scene, sid, step)
        el = el0 + angle_jitter * rng.standard_normal(el0.shape)
        az = az0 + angle_jitter * rng.standard_normal(az0.shape)
        d_local = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        d_world = d_local @ T[:3, :3].T
        o = T[:3, 3]
        depth = np.full(len(d_world), np.inf)
        cosi = np.zeros(len(d_world))
        for n, off, lo, hi in planes:
            n = np.asarray(n, dtype=np.float64)
            denom = d_world @ n
            with np.errstate(divide='ignore', invalid='ignore'):
                t = (off - o @ n) / denom
            hit = (t > 0) & (t < depth)
            if np.isfinite(lo).any() or np.isfinite(hi).any():
                with np.errstate(invalid='ignore'):
                    pt = o + t[:, None] * d_world
                    hit &= np.all((pt >= np.asarray(lo) - 1e-9) & (pt <= np.asarray(hi) + 1e-9), axis=1)
            depth = np.where(hit, t, depth)
            cosi = np.where(hit, np.abs(denom), cosi)
        keep = np.isfinite(depth) & (depth >= depth_clip[0]) & (depth <= depth_clip[1])
        depth, d_local, cosi = depth[keep], d_local[keep], cosi[keep]
        if bias_w is not None:
            g = np.arccos(np.clip(cosi, 0.0, 1.0))
            bias = sum(w * g ** e for w, e in zip(bias_w, bias_exponent))
            depth = depth / (1.0 - bias)
        depth = depth + range_noise * rng.standard_normal(depth.shape)
        pts = (depth[:, None] * d_local).astype(np.float32)
        if grid_res > 0.0:
            pts = pts[voxel_keep_first(pts.astype(np.float64), grid_res)]
        scans.append({'points': pts, 'vps': np.zeros_like(pts)})
        poses_gt.append(T)
        Tn = T.copy()
        if k + first_scan > 0 and (pose_noise[0] > 0 or pose_noise[1] > 0):
            Tn = T @ _yaw_pose(*(pose_noise[0] * rng.standard_normal(3)), pose_noise[1] * rng.standard_normal())
        poses_init.append(Tn)
    return scans, np.stack(poses_gt), np.stack(poses_init)
