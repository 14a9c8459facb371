"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, via oracle/ref_shim.py) on small seeded synthetic scenes.

Run in the build container (the reference tree does not exist on the GPU box):

    python -m oracle.make_golden

Inputs are float32 (stored) and handed to the reference up-cast to float64, which is what
the parity tests feed both the oracle restatement and the CUDA path.
"""
import contextlib
import io
import os
import types

import numpy as np
import torch

from depth_correction_b200.synthetic import make_sequence
from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ref_cfg(ref, **kw):
    cfg = types.SimpleNamespace(
        nn_type=ref.NeighborhoodType.ball, nn_k=0, nn_r=0.4, nn_scale=None,
        shadow_angle_bounds=[], shadow_neighborhood_angle=None,
        eigenvalue_bounds=[], eigenvalue_ratio_bounds=[[0, 1, 0, 0.25], [1, 2, 0.25, 1.0]],
        min_valid_neighbors=5, dir_dispersion_bounds=[], vp_dispersion_bounds=[],
        vp_dispersion_to_depth2_bounds=[], log_filters=False, device='cpu', float_type='float64')
    cfg.numpy_float_type = lambda: np.float64
    cfg.torch_float_type = lambda: torch.float64
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def gen_nn(ref):
    rng = np.random.default_rng(11)
    # noisy plane + a few isolated points (rows with only self) + one far outlier
    n = 1500
    pts = np.concatenate([
        np.stack([rng.uniform(0, 4, n), rng.uniform(0, 3, n), 0.01 * rng.standard_normal(n)], 1),
        rng.uniform(20, 40, (6, 3)),
    ]).astype(np.float32)
    p64 = torch.as_tensor(pts.astype(np.float64))
    out = {'points': pts}
    _, out['radius_r0.4'] = ref.nearest_neighbors(p64, p64, r=0.4)
    _, out['radius_r0.15'] = ref.nearest_neighbors(p64, p64, r=0.15)
    out['knn8_dist'], out['knn8'] = ref.nearest_neighbors(p64, p64, k=8)
    out['knn16_r0.3_dist'], out['knn16_r0.3'] = ref.nearest_neighbors(p64, p64, k=16, r=0.3)
    out['knn32_r0.1_dist'], out['knn32_r0.1'] = ref.nearest_neighbors(p64, p64, k=32, r=0.1)
    # cross query (query != points), used by the ICP-style callers
    q = (pts[::7] + 0.01).astype(np.float32)
    out['query'] = q
    out['cross_knn4_dist'], out['cross_knn4'] = ref.nearest_neighbors(p64, torch.as_tensor(q.astype(np.float64)), k=4)
    # boundary / tie semantics on an exact lattice: <= r for balls, < r for kNN upper bound
    lat = np.array([[0, 0, 0], [0.5, 0, 0], [0, 0.5, 0], [0, 0, 0.5], [-0.5, 0, 0], [1, 0, 0], [0.25, 0, 0]],
                   dtype=np.float32)
    l64 = torch.as_tensor(lat.astype(np.float64))
    out['lattice'] = lat
    _, out['lattice_radius_r0.5'] = ref.nearest_neighbors(l64, l64, r=0.5)
    out['lattice_knn3_r0.5_dist'], out['lattice_knn3_r0.5'] = ref.nearest_neighbors(l64, l64, k=3, r=0.5)
    np.savez_compressed(os.path.join(OUT, 'nn.npz'),
                        **{k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print('nn.npz', {k: tuple(np.shape(v)) for k, v in out.items()})


def build_scans(ref, cfg, scans_np):
    clouds = []
    for sc in scans_np:
        cloud = ref.DepthCloud.from_points(torch.as_tensor(sc['points'].astype(np.float64)),
                                           vps=torch.as_tensor(sc['vps'].astype(np.float64)))
        cloud = quiet(ref.local_feature_cloud, cloud, cfg)
        clouds.append(cloud)
    return clouds


def gen_features(ref):
    scans_np, poses, _ = make_sequence('corridor', n_scans=1, pattern='os0-32', seed=3, grid_res=0.15)
    pts = scans_np[0]['points']
    for tag, kw in (('r', dict(r=0.4)), ('kr', dict(k=12, r=0.5))):
        cloud = ref.DepthCloud.from_points(torch.as_tensor(pts.astype(np.float64)))
        quiet(cloud.update_all, **kw)
        out = {'points': pts, 'neighbors': cloud.neighbors.numpy().astype(np.int32),
               'weights_dtype': str(cloud.weights.dtype), 'weights_shape': np.array(cloud.weights.shape)}
        for f in ('mean', 'cov', 'eigvals', 'eigvecs', 'normals', 'inc_angles', 'dirs', 'depth'):
            out[f] = getattr(cloud, f).numpy()
        if cloud.distances is not None:
            out['distances'] = cloud.distances.numpy()
        cfg = ref_cfg(ref)
        out['mask_ratio'] = ref.filters.filter_eigenvalue_ratios(cloud, cfg.eigenvalue_ratio_bounds, only_mask=True).numpy()
        out['mask_eig'] = ref.filters.filter_eigenvalues(cloud, [[0, None, 0.01], [1, 0.0025, None]], only_mask=True).numpy()
        out['mask_valid'] = ref.filters.filter_valid_neighbors(cloud, min=5, only_mask=True).numpy()
        out['vp_dispersion'] = cloud.vp_dispersion().numpy()
        out['dir_dispersion'] = cloud.dir_dispersion().numpy()
        np.savez_compressed(os.path.join(OUT, 'features_%s.npz' % tag), **out)
        print('features_%s.npz' % tag, pts.shape, cloud.neighbors.shape)


def gen_steps(ref):
    """The north-star training iteration (scripts/model_poses_learning:119-135), variants."""
    variants = [
        # tag, model, w, exponent, loss, loss_kwargs, nn(k,r), use loss mask, pose_correction
        ('scaled_mineig_norm_r', 'ScaledPolynomial', [0.002, -0.004], [2.0, 4.0], 'min_eigval_loss',
         dict(normalization=True), (0, 0.4), True, 'pose'),
        ('scaled_mineig_sqrt_kr', 'ScaledPolynomial', [-0.01], [4.0], 'min_eigval_loss',
         dict(normalization=False, sqrt=True), (16, 0.5), False, 'pose'),
        ('poly_trace_r', 'Polynomial', [0.01, 0.003], [2.0, 4.0], 'trace_loss', dict(), (0, 0.4), True, 'pose'),
        ('scaled_trace_k_common', 'ScaledPolynomial', [0.004], [2.0], 'trace_loss', dict(sqrt=True), (10, 0), False,
         'common'),
        ('scaled_mineig_norm_sum', 'ScaledPolynomial', [0.0, 0.0], [2.0, 4.0], 'min_eigval_loss',
         dict(normalization=True, reduction='sum'), (24, 0.4), True, 'none'),
    ]
    for tag, model_name, w, e, loss_name, loss_kwargs, (k, r), use_mask, pose_corr in variants:
        cfg = ref_cfg(ref, nn_k=k, nn_r=r or None)
        scans_np, poses_gt, poses_init = make_sequence('fee', n_scans=4, pattern='os0-32', seed=5, grid_res=0.2,
                                                       pose_noise=(0.01, 0.005), bias_w=[-0.01], bias_exponent=[4.0])
        clouds = build_scans(ref, cfg, scans_np)
        poses = torch.as_tensor(poses_init)
        rng = np.random.default_rng(17)
        S = len(clouds)
        if pose_corr == 'pose':
            deltas = torch.as_tensor(rng.normal(0, [0.01] * 3 + [0.004] * 3, (S, 6)), dtype=torch.float64)
        elif pose_corr == 'common':
            deltas = torch.as_tensor(rng.normal(0, [0.01] * 3 + [0.004] * 3, (1, 6)), dtype=torch.float64)
        else:
            deltas = None
        neighbors, weights = quiet(ref.establish_neighborhoods, clouds=clouds, poses=poses, cfg=cfg)
        model = getattr(ref, model_name)(w=w, exponent=e)
        if deltas is not None:
            deltas.requires_grad_(True)
            # eval.py:68-82 (create_corrected_poses; eval.py itself needs ROS-only imports)
            poses_c = torch.stack([torch.matmul(poses[i], ref.xyz_axis_angle_to_matrix(deltas[i if len(deltas) > 1 else 0]))
                                   for i in range(S)])
            poses_c.retain_grad()
        else:
            poses_c = poses.clone().requires_grad_(True)
        cloud = ref.global_cloud(clouds=clouds, model=model, poses=poses_c)
        feats = quiet(ref.compute_neighborhood_features, cloud=cloud, model=None, neighborhoods=(neighbors, weights), cfg=cfg)
        mask = None
        if use_mask:
            mask = quiet(ref.global_cloud_mask, feats, feats.mask.clone() if feats.mask is not None else None, cfg)
        kw = dict(loss_kwargs)
        if 'reduction' in kw:
            kw['reduction'] = ref.Reduction(kw['reduction'])
        loss, loss_cloud = quiet(getattr(ref, loss_name), feats, mask=mask, **kw)
        loss.backward()
        out = {
            'n_scans': S, 'model': model_name, 'w': np.array(w), 'exponent': np.array(e), 'loss_name': loss_name,
            'loss_kwargs': repr(loss_kwargs), 'nn_k': k, 'nn_r': r, 'pose_correction': pose_corr,
            'poses': poses.numpy(), 'neighbors': neighbors.numpy().astype(np.int32),
            'loss': loss.detach().numpy(), 'per_point': loss_cloud.loss.detach().numpy(),
            'eigvals': feats.eigvals.detach().numpy(), 'cov': feats.cov.detach().numpy(),
            'points': feats.points.detach().numpy(), 'w_grad': model.w.grad.numpy(),
            'poses_grad': poses_c.grad.numpy(),
        }
        if deltas is not None:
            out['pose_deltas'] = deltas.detach().numpy()
            out['pose_deltas_grad'] = deltas.grad.numpy()
        if mask is not None:
            out['loss_mask'] = mask.numpy()
        for i, (sc, c) in enumerate(zip(scans_np, clouds)):
            out['scan%d_points' % i] = sc['points']
            out['scan%d_inc_angles' % i] = c.inc_angles.numpy()
            out['scan%d_mask' % i] = c.mask.numpy()
        np.savez_compressed(os.path.join(OUT, 'step_%s.npz' % tag), **out)
        print('step_%s.npz' % tag, 'N=%d K=%d loss=%.9g' % (neighbors.shape[0], neighbors.shape[1], loss.item()),
              'w_grad', model.w.grad.numpy().ravel())


def gen_misc(ref):
    # model.py:357-364 `test_model` values (the reference prints them without asserting)
    model = ref.ScaledPolynomial(exponent=[2, 4], w=[-0.06, -0.06])
    angles = np.linspace(10, 85, 10) / 180.0
    cloud = ref.DepthCloud.from_points(torch.tensor([[20.0, 0.0, 0.0]] * 10, dtype=torch.float64))
    cloud.inc_angles = torch.as_tensor(angles)[:, None]
    d_scaled = model(cloud).depth.detach().numpy()
    d_poly = ref.Polynomial(exponent=[2, 4], w=[-0.06, -0.06])(cloud).depth.detach().numpy()
    cloud.mask = torch.tensor([True, False] * 5)
    d_masked = model(cloud).depth.detach().numpy()
    rng = np.random.default_rng(23)
    xyzaa = np.concatenate([rng.normal(0, 1, (16, 6)), np.zeros((1, 6)), [[1, 2, 3, 1e-9, 0, 0]],
                            [[0, 0, 0, np.pi, 0, 0]]])
    mats = ref.xyz_axis_angle_to_matrix(torch.as_tensor(xyzaa)).numpy()
    np.savez_compressed(os.path.join(OUT, 'misc.npz'), angles=angles, d_scaled=d_scaled, d_poly=d_poly,
                        d_masked=d_masked, xyz_axis_angle=xyzaa, matrices=mats)
    print('misc.npz')


def gen_filters(ref):
    """filter_grid (filters.py:24-82) for every keep mode / order, and filter_shadow_points (filters.py:257-309)."""
    scans, _, _ = make_sequence('corridor', n_scans=1, pattern='os0-32', seed=21)
    pts = scans[0]['points']                                  # float32 [n,3], sensor frame
    out = {'points': pts}
    for dt, tag in ((np.float32, 'f32'), (np.float64, 'f64')):
        x = pts.astype(dt)
        for keep in ('first', 'last', 'random'):
            for po in (False, True):
                rng = np.random.default_rng(135)
                ind = quiet(ref.filters.filter_grid, x, 0.2, only_mask=True, keep=keep, preserve_order=po, rng=rng)
                out['grid_%s_%s_%d' % (tag, keep, int(po))] = np.asarray(ind, dtype=np.int64)
    # two consecutive draws from ONE generator (the reference's module-level default_rng is stateful)
    rng = np.random.default_rng(7)
    out['grid_two_draws_a'] = np.asarray(quiet(ref.filters.filter_grid, pts, 0.35, only_mask=True, keep='random', rng=rng), dtype=np.int64)
    out['grid_two_draws_b'] = np.asarray(quiet(ref.filters.filter_grid, pts, 0.35, only_mask=True, keep='random', rng=rng), dtype=np.int64)
    # shadow points: a wall behind a thin pole seen from the origin produces mixed-depth beams
    rng = np.random.default_rng(5)
    az = rng.uniform(-0.6, 0.6, 4000)
    el = rng.uniform(-0.3, 0.3, 4000)
    d = np.where(np.abs(az) < 0.05, 2.0, 6.0) / np.cos(az) + 0.01 * rng.standard_normal(4000)
    edge = (np.abs(np.abs(az) - 0.05) < 0.004)
    d[edge] = rng.uniform(2.0, 6.0, edge.sum())               # mixed pixels along the occlusion boundary
    sp = (d[:, None] * np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], 1)).astype(np.float32)
    out['shadow_points'] = sp
    for tag, bounds in (('a', [0.0873, None]), ('b', [0.2, 2.8])):
        cloud = ref.DepthCloud.from_points(torch.as_tensor(sp.astype(np.float64)))
        cloud.update_dir_neighbors(angle=0.02)
        cloud.loss = torch.arange(len(sp), dtype=torch.float64)[:, None]
        kept = quiet(ref.filters.filter_shadow_points, cloud, list(bounds))
        out['shadow_kept_' + tag] = kept.loss[:, 0].numpy().astype(np.int64)
        out['shadow_K_' + tag] = np.asarray(cloud.dir_neighbors.shape[1])
    np.savez_compressed(os.path.join(OUT, 'filters.npz'), **out)
    print('filters.npz', {k: v.shape for k, v in out.items() if k.startswith('grid_f32') or k.startswith('shadow_kept')})


def gen_stats(ref):
    """Neighbourhood statistics and the loss mask of the GLOBAL cloud (depth_cloud.py:314-354, preproc.py:122-164)."""
    scans_np, poses, _ = make_sequence('corridor', n_scans=3, pattern='os0-32', seed=17, grid_res=0.15, step=1.5)
    clouds = []
    cfg = ref_cfg(ref)
    for s in scans_np:
        c = ref.DepthCloud.from_points(torch.as_tensor(s['points'].astype(np.float64)))
        clouds.append(quiet(ref.local_feature_cloud, c, cfg))
    poses_t = torch.as_tensor(poses)
    cloud = quiet(ref.global_cloud, clouds=clouds, model=None, poses=poses_t)
    quiet(cloud.update_all, r=0.4)
    out = {'n_scans': np.asarray(len(scans_np)), 'poses': poses}
    for i, s in enumerate(scans_np):
        out['scan%d_points' % i] = s['points']
    for f in ('vp_dispersion', 'dir_dispersion', 'mean_depth', 'mean_vp_dist', 'vp_dispersion_to_depth2', 'vp_dist_to_depth'):
        out[f] = getattr(cloud, f)().numpy()
    mcfg = ref_cfg(ref, min_valid_neighbors=8, eigenvalue_bounds=[[0, None, 0.01]], dir_dispersion_bounds=[0.0, 0.02],
                   vp_dispersion_bounds=[0.05, float('inf')], vp_dispersion_to_depth2_bounds=[0.001, None])
    out['mask'] = quiet(ref.global_cloud_mask, cloud, cloud.mask.clone(), mcfg).numpy()
    # the individual criteria, for a readable failure
    out['mask_start'] = cloud.mask.numpy()
    out['mask_valid'] = ref.filters.filter_valid_neighbors(cloud, min=8, only_mask=True).numpy()
    out['mask_eig'] = ref.filters.filter_eigenvalues(cloud, mcfg.eigenvalue_bounds, only_mask=True).numpy()
    out['mask_ratio'] = ref.filters.filter_eigenvalue_ratios(cloud, mcfg.eigenvalue_ratio_bounds, only_mask=True).numpy()
    out['eigvals'] = cloud.eigvals.numpy()
    np.savez_compressed(os.path.join(OUT, 'stats.npz'), **out)
    print('stats.npz', cloud.neighbors.shape, 'mask keeps', int(out['mask'].sum()))


def gen_icp(ref):
    """icp_loss (loss.py:373-565): loss and gradients to the model weights and the poses, point-to-plane and
    point-to-point, searched and precomputed correspondences.  The reference's scipy branch (differentiable=False)
    cannot run with gradients (it hands tensors that require grad to cKDTree, loss.py:442), so the search goes
    through its pytorch3d branch with ref_shim.knn_points_exact standing in for the absent pytorch3d."""
    from depth_correction.loss import icp_loss
    scans_np, poses, _ = make_sequence('corridor', n_scans=3, pattern='os0-32', seed=29, grid_res=0.15, step=0.7,
                                       pose_noise=(0.02, 0.01))
    cfg = ref_cfg(ref)
    out = {'n_scans': np.asarray(len(scans_np)), 'poses': poses, 'w': np.array([0.002, -0.001]), 'exponent': np.array([2.0, 4.0])}
    for i, s in enumerate(scans_np):
        out['scan%d_points' % i] = s['points']

    def run(p2pl, masks=None):
        clouds = []
        for s in scans_np:
            c = ref.DepthCloud.from_points(torch.as_tensor(s['points'].astype(np.float64)))
            clouds.append(quiet(ref.local_feature_cloud, c, cfg))
        model = ref.ScaledPolynomial(w=[0.002, -0.001], exponent=[2.0, 4.0])
        poses_t = torch.as_tensor(poses).clone().requires_grad_(True)
        loss, _ = quiet(icp_loss, [clouds], poses=[list(poses_t)], model=model, masks=masks, icp_point_to_plane=p2pl,
                        icp_inlier_ratio=0.5, differentiable=True)
        loss.backward()
        return loss.item(), model.w.grad.numpy().copy(), poses_t.grad.numpy().copy(), clouds

    for tag, p2pl in (('plane', True), ('point', False)):
        loss, gw, gp, clouds = run(p2pl)
        out['%s_loss' % tag], out['%s_w_grad' % tag], out['%s_poses_grad' % tag] = np.asarray(loss), gw, gp
    for i, c in enumerate(clouds):
        out['scan%d_inc_angles' % i] = c.inc_angles.numpy()
        out['scan%d_mask' % i] = c.mask.numpy()
        out['scan%d_normals' % i] = c.normals.numpy()
    # precomputed correspondences: a boolean mask over scan i and an index list into scan i+1
    rng = np.random.default_rng(2)
    masks = []
    for i in range(len(scans_np) - 1):
        n1, n2 = len(scans_np[i]['points']), len(scans_np[i + 1]['points'])
        m1 = rng.random(n1) < 0.3
        m2 = rng.integers(0, n2, int(m1.sum()))
        out['mask%d_1' % i], out['mask%d_2' % i] = m1, m2
        masks.append((torch.as_tensor(m1), torch.as_tensor(m2)))
    loss, gw, gp, _ = run(True, masks=[masks])
    out['masked_plane_loss'], out['masked_plane_w_grad'], out['masked_plane_poses_grad'] = np.asarray(loss), gw, gp
    np.savez_compressed(os.path.join(OUT, 'icp.npz'), **out)
    print('icp.npz', {k: float(out[k]) for k in ('plane_loss', 'point_loss', 'masked_plane_loss')})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shim.load()
    torch.set_default_dtype(torch.float32)
    gen_nn(ref)
    gen_features(ref)
    gen_steps(ref)
    gen_misc(ref)
    gen_filters(ref)
    gen_stats(ref)
    gen_icp(ref)


if __name__ == '__main__':
    main()
