"""GPU parity tests: the sm_100a path (through the C-ABI library) against
  (a) golden vectors produced by the unmodified reference (tests/golden, float64 clouds), and
  (b) the CPU oracle on the same float32 values (the storage type of the bench path).
Tolerance (north_star): neighbour indices bit-exact; loss / eigenvalues / gradients <= 1e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from helpers import RTOL, STEP_TAGS, eig_close, icp_inputs, rel_err, rel_err_norm, same_neighbor_sets, step_inputs, well_separated

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dc():
    import depth_correction_b200 as dc
    return dc


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


# ------------------------------------------------------------------------------------------------
# kernel 1: neighbour search
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
@pytest.mark.parametrize('key,kw', [('radius_r0.4', dict(r=0.4)), ('radius_r0.15', dict(r=0.15)), ('knn8', dict(k=8)),
                                    ('knn16_r0.3', dict(k=16, r=0.3)), ('knn32_r0.1', dict(k=32, r=0.1))])
def test_nn_bit_exact_vs_reference(dc, dev, golden, key, kw, dtype):
    g = golden('nn')
    p = torch.as_tensor(g['points'], device=dev).to(dtype)      # float32 values, exact in both dtypes
    dist, idx = dc.nearest_neighbors(p, p, **kw)
    assert idx.dtype == torch.int64 and idx.is_cuda
    assert np.array_equal(idx.cpu().numpy(), g[key]), key
    if key + '_dist' in g.files:
        assert dist.dtype == torch.float64
        assert np.array_equal(dist.cpu().numpy(), g[key + '_dist'])     # sqrt of the same fp64 d2: bit-exact
    else:
        assert dist is None


def test_nn_cross_query(dc, dev, golden):
    g = golden('nn')
    p = torch.as_tensor(g['points'], device=dev)
    q = torch.as_tensor(g['query'], device=dev)
    dist, idx = dc.nearest_neighbors(p, q, k=4)
    assert np.array_equal(idx.cpu().numpy(), g['cross_knn4'])
    assert np.array_equal(dist.cpu().numpy(), g['cross_knn4_dist'])


def test_nn_boundary_and_ties(dc, dev, golden):
    """d == r is inside a ball and outside a kNN upper bound; exact ties are compared as sets
    (cKDTree returns them in tree-traversal order, we use (d2, index) order)."""
    g = golden('nn')
    lat = torch.as_tensor(g['lattice'], device=dev)
    _, idx = dc.nearest_neighbors(lat, lat, r=0.5)
    assert np.array_equal(idx.cpu().numpy(), g['lattice_radius_r0.5'])
    dist, idx = dc.nearest_neighbors(lat, lat, k=3, r=0.5)
    assert np.array_equal(dist.cpu().numpy(), g['lattice_knn3_r0.5_dist'])
    # row 6 has an exact tie (points 0 and 1 both at 0.25): either member of the tie group is correct,
    # so indices are validated through the distances they realise
    ours, ref = idx.cpu().numpy(), g['lattice_knn3_r0.5']
    assert np.array_equal(ours >= 0, ref >= 0)
    lat64 = g['lattice'].astype(np.float64)
    for row in range(len(ours)):
        for c in range(3):
            if ours[row, c] >= 0:
                assert np.linalg.norm(lat64[ours[row, c]] - lat64[row]) == g['lattice_knn3_r0.5_dist'][row, c]
    assert ours[6, 1] in (0, 1)  # our rule: (d2, position in the cell-sorted map) lexicographic
    assert idx[0].tolist() == [0, 6, -1]


def test_nn_edge_cases(dc, dev):
    one = torch.zeros((1, 3), device=dev)
    d, i = dc.nearest_neighbors(one, one, r=1.0)
    assert i.tolist() == [[0]] and d is None
    d, i = dc.nearest_neighbors(one, one, k=3)
    assert i.tolist() == [[0, -1, -1]] and d[0, 0].item() == 0.0 and torch.isinf(d[0, 1:]).all()
    # 33 points: a full slice plus a ragged tail; duplicates of one location
    p = torch.zeros((33, 3), device=dev)
    p[:, 0] = torch.arange(33, device=dev) // 3
    _, i = dc.nearest_neighbors(p, p, r=0.5)
    assert i.shape == (33, 3) and (i >= 0).all()
    assert i[4].tolist() == [3, 4, 5]
    with pytest.raises(AssertionError):
        dc.nearest_neighbors(p, p)


def test_nn_large_random_vs_ckdtree(dc, dev):
    """Full-size style check against the oracle (cKDTree) on 200k clustered points."""
    from oracle import oracle
    rng = np.random.default_rng(7)
    c = rng.uniform(-20, 20, (400, 3))
    pts = (c[rng.integers(0, 400, 200000)] + rng.normal(0, 0.3, (200000, 3))).astype(np.float32)
    p = torch.as_tensor(pts, device=dev)
    p64 = torch.as_tensor(pts.astype(np.float64))
    for kw in (dict(r=0.12), dict(k=16, r=0.25), dict(k=8)):
        dist, idx = dc.nearest_neighbors(p, p, **kw)
        d_ref, i_ref = oracle.nearest_neighbors(p64, k=kw.get('k'), r=kw.get('r'))
        assert torch.equal(idx.cpu(), i_ref), kw
        if d_ref is not None:
            assert torch.equal(dist.cpu(), d_ref)


def _corridor_world_points(n_scans, rings, azimuths):
    from depth_correction_b200.synthetic import make_sequence
    scans, poses, _ = make_sequence('corridor', n_scans=n_scans, pattern='os0-128', seed=3, rings=rings, azimuths=azimuths)
    out = []
    for s, T in zip(scans, poses):
        out.append(s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    return np.concatenate(out)


@pytest.mark.parametrize('cell', [None, 0.05, 0.23])
def test_knn_lidar_map_vs_ckdtree(dc, dev, cell):
    """Lidar-shaped map (density varies by orders of magnitude, points on planes, fp64 world coordinates) with
    the automatic cell size, much smaller and much larger cells: ring skipping / growth must reproduce cKDTree."""
    from oracle import oracle
    from depth_correction_b200.graph import search
    pts = _corridor_world_points(5, 64, 512)
    p = torch.as_tensor(pts, device=dev)
    p64 = torch.as_tensor(pts)
    for kw in (dict(k=32, r=0.4), dict(k=12), dict(k=48, r=0.25)):
        g = search(p, None, cell=cell, **kw)
        idx, dist = g.neighbors(), g.distances()
        d_ref, i_ref = oracle.nearest_neighbors(p64, **kw)
        assert torch.equal(idx.cpu(), i_ref), (cell, kw)
        assert torch.equal(dist.cpu(), d_ref), (cell, kw)


def test_knn_duplicates_and_sparse_tail(dc, dev):
    """Exact duplicates (more than 8 ties in the boundary bin -> second histogram level / repeated minimum) and a
    sparse tail (ring growth, ring skipping from the cell table): distances bit-exact vs cKDTree; indices are
    validated through the distances they realise (tie order is implementation defined)."""
    from oracle import oracle
    rng = np.random.default_rng(11)
    c = rng.uniform(-5, 5, (50, 3))
    pts = (c[rng.integers(0, 50, 60000)] + rng.normal(0, 0.2, (60000, 3))).astype(np.float32)
    pts[1000:1400] = pts[:400]                       # exact duplicates
    pts[2000:2040] = pts[0]                          # 41 copies of one location
    pts[3000:3200] = rng.uniform(-40, 40, (200, 3)).astype(np.float32)     # isolated points far from everything
    p = torch.as_tensor(pts, device=dev)
    p64 = torch.as_tensor(pts.astype(np.float64))
    for kw in (dict(k=16), dict(k=32, r=0.3), dict(k=5, r=0.05)):
        dist, idx = dc.nearest_neighbors(p, p, **kw)
        d_ref, i_ref = oracle.nearest_neighbors(p64, **kw)
        assert torch.equal(dist.cpu(), d_ref), kw
        idx = idx.cpu()
        assert torch.equal(idx >= 0, i_ref >= 0)
        df = p64[idx.clamp(min=0)] - p64[:, None, :]
        realised = (df * df).sum(dim=2).sqrt()
        fin = idx >= 0
        assert torch.allclose(realised[fin], d_ref[fin], rtol=1e-14, atol=0.0), kw    # (torch may contract / reorder the sum)
        srt = idx.sort(dim=1).values
        assert ((srt[:, 1:] != srt[:, :-1]) | (srt[:, 1:] < 0)).all(), 'a neighbour is listed twice'


def test_cell_table_is_lower_bound_of_sorted_keys(dc, dev):
    """cell_start[c] = first sorted position whose key is >= c (long empty runs, empty head / tail, one point)."""
    from depth_correction_b200.graph import SortedMap
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.normal(0, 0.05, (3000, 3)), rng.normal(0, 0.05, (2000, 3)) + [40.0, 3.0, -2.0],
                          [[-7.0, -7.0, -7.0]]]).astype(np.float32)
    for p, cell in ((pts, 0.11), (pts[:1], 0.5), (pts, 3.0)):
        smap = SortedMap(torch.as_tensor(p, device=dev), cell)
        ref = torch.searchsorted(smap.keys >> int(smap.spec.sub_bits), torch.arange(smap.n_cells + 1, device=dev, dtype=torch.int64))
        assert torch.equal(smap.cell_start.long(), ref)


def test_graph_roundtrip_and_transpose(dc, dev, golden):
    from depth_correction_b200.graph import Graph, SortedMap, search
    g = golden('nn')
    p = torch.as_tensor(g['points'], device=dev)
    nb = torch.as_tensor(g['knn16_r0.3'], device=dev)
    gr = Graph.from_padded(SortedMap(p, 0.3), nb)
    back = gr.neighbors()
    assert torch.equal(back, nb)            # valid entries keep their order, padding stays last
    t = gr.transposed().neighbors().cpu().numpy()
    ref = [[] for _ in range(len(nb))]
    for i, row in enumerate(g['knn16_r0.3']):
        for j in row:
            if j >= 0:
                ref[j].append(i)
    for j in range(len(nb)):
        assert sorted(x for x in t[j] if x >= 0) == ref[j]
    sym = search(p, None, r=0.15)
    assert sym.symmetric and sym.transposed() is sym
    assert torch.equal(sym.valid_counts().cpu(), torch.as_tensor((g['radius_r0.15'] >= 0).sum(1)))


# ------------------------------------------------------------------------------------------------
# staged DepthCloud API (unfused kernels) against the reference's update_all
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('tag,kw', [('r', dict(r=0.4)), ('kr', dict(k=12, r=0.5))])
def test_update_all_vs_reference(dc, dev, golden, tag, kw):
    g = golden('features_' + tag)
    cloud = dc.DepthCloud.from_points(torch.as_tensor(g['points'].astype(np.float64), device=dev))
    cloud.update_all(**kw)
    assert np.array_equal(cloud.neighbors.cpu().numpy(), g['neighbors'])
    assert str(cloud.weights.dtype) == str(g['weights_dtype']) and list(cloud.weights.shape) == list(g['weights_shape'])
    if 'distances' in g.files:
        # from_points normalises on the GPU here and on the CPU in the reference: points differ by an ulp
        ours, ref = cloud.distances.cpu().numpy(), g['distances']
        assert np.array_equal(np.isinf(ours), np.isinf(ref))
        assert np.max(np.abs(ours[np.isfinite(ref)] - ref[np.isfinite(ref)])) < 1e-14
    assert np.max(np.abs(cloud.mean.cpu().numpy() - g['mean'])) < 1e-12
    assert np.max(np.abs(cloud.cov.cpu().numpy() - g['cov'])) < 1e-13
    assert eig_close(cloud.eigvals.cpu().numpy(), g['eigvals'], 1e-9)
    ok = well_separated(g['eigvals'])
    V = cloud.eigvecs.cpu().numpy()
    dots = np.abs(np.einsum('nij,nij->nj', V, g['eigvecs']))
    assert np.all(np.abs(dots[ok] - 1) < 1e-7)
    # every decomposition must reconstruct its matrix, also in degenerate rows
    C = cloud.cov.cpu().numpy()
    rec = np.einsum('nij,nj,nkj->nik', V, cloud.eigvals.cpu().numpy(), V)
    assert np.max(np.abs(rec - C)) < 1e-12
    assert np.max(np.abs(cloud.normals.cpu().numpy()[ok] - g['normals'][ok])) < 1e-7
    assert np.max(np.abs(cloud.inc_angles.cpu().numpy()[ok] - g['inc_angles'][ok])) < 1e-6
    ratio = dc.filter_eigenvalue_ratios(cloud, [[0, 1, 0, 0.25], [1, 2, 0.25, 1.0]], only_mask=True).cpu().numpy()
    assert (ratio != g['mask_ratio']).mean() < 1e-3      # knife-edge rows may flip at 1e-16
    assert np.array_equal(dc.filter_valid_neighbors(cloud, min=5, only_mask=True).cpu().numpy(), g['mask_valid'])
    assert np.max(np.abs(cloud.vp_dispersion().cpu().numpy() - g['vp_dispersion'])) < 1e-12
    assert np.max(np.abs(cloud.dir_dispersion().cpu().numpy() - g['dir_dispersion'])) < 1e-12


def test_staged_autograd_matches_torch(dc, dev):
    """update_cov / update_eig backward kernels vs torch autograd of the same formulas."""
    from oracle import oracle
    rng = np.random.default_rng(3)
    pts = torch.as_tensor(rng.normal(0, 1, (300, 3)) * [1.0, 0.6, 0.05], device=dev).requires_grad_(True)
    _, nb = oracle.nearest_neighbors(pts.detach().cpu(), k=9)
    nb[::7, -2:] = -1
    cloud = dc.DepthCloud.from_points(pts.detach())
    cloud.points = pts
    cloud.neighbors = nb.to(dev)
    cloud.update_mean()
    cloud.update_cov()
    cloud.update_eig()
    coef = torch.as_tensor(rng.normal(0, 1, (300, 3)), device=dev)
    (cloud.eigvals * coef).sum().backward()
    g_ours = pts.grad.clone()
    p2 = pts.detach().cpu().clone().requires_grad_(True)
    f = oracle.neighborhood_features(p2, nb, eigvecs=False)
    (f['eigvals'] * coef.cpu()).sum().backward()
    assert rel_err_norm(g_ours.cpu().numpy(), p2.grad.numpy()) < 1e-9
    # mean + cov + eigenvector adjoint
    pts.grad = None
    cloud.update_mean()
    cloud.update_cov()
    cloud.update_eig()
    (cloud.mean.sum() + (cloud.cov ** 2).sum() + (cloud.eigvecs[:, :, 0] ** 2 * coef).sum()).backward()
    p3 = pts.detach().cpu().clone().requires_grad_(True)
    f = oracle.neighborhood_features(p3, nb)
    (f['mean'].sum() + (f['cov'] ** 2).sum() + (f['eigvecs'][:, :, 0] ** 2 * coef.cpu()).sum()).backward()
    assert rel_err_norm(pts.grad.cpu().numpy(), p3.grad.numpy()) < 1e-7


def test_eigh3_degenerate(dc, dev):
    from depth_correction_b200 import ops
    mats = torch.zeros((6, 3, 3), dtype=torch.float64, device=dev)
    mats[1] = torch.eye(3, device=dev) * 2.5                                  # triple
    mats[2] = torch.diag(torch.tensor([1.0, 1.0, 3.0], device=dev))           # double low
    mats[3] = torch.diag(torch.tensor([3.0, 1e-30, 3.0], device=dev))         # double high
    v = torch.tensor([1.0, 2.0, -1.0], dtype=torch.float64, device=dev)
    mats[4] = torch.outer(v, v)                                               # rank one
    mats[5] = torch.tensor([[4e-5, 1e-5, 0], [1e-5, 3e-2, 2e-3], [0, 2e-3, 5e-2]], device=dev)
    lam, V = ops.eigh3(mats)
    assert torch.isfinite(lam).all() and torch.isfinite(V).all()
    ref = torch.linalg.eigvalsh(mats.cpu())
    assert (lam.cpu() - ref).abs().max() < 1e-13 * 6
    rec = torch.einsum('nij,nj,nkj->nik', V, lam, V)
    assert (rec - mats).abs().max() < 1e-12
    eye = torch.einsum('nij,nik->njk', V, V)
    assert (eye - torch.eye(3, device=dev)).abs().max() < 1e-12
    bad = torch.full((1, 3, 3), float('nan'), dtype=torch.float64, device=dev)
    lam, V = ops.eigh3(bad)
    assert torch.isnan(lam).all()


def test_pose_compose_kernel(dc, dev, golden):
    from depth_correction_b200 import ops
    from oracle import oracle
    g = golden('misc')
    d = torch.as_tensor(g['xyz_axis_angle'], device=dev)
    S = d.shape[0]
    eye = torch.eye(4, dtype=torch.float64, device=dev).expand(S, 4, 4).contiguous()
    assert (ops.pose_compose(eye, d).cpu().numpy() - g['matrices']).__abs__().max() < 1e-15
    rng = np.random.default_rng(2)
    P = oracle.xyz_axis_angle_to_matrix(torch.as_tensor(rng.normal(0, 1, (S, 6))))
    G = torch.as_tensor(rng.normal(0, 1, (S, 4, 4)))
    for deltas in (torch.as_tensor(g['xyz_axis_angle']), torch.zeros((S, 6), dtype=torch.float64),
                   torch.as_tensor(g['xyz_axis_angle'][:1])):
        d_ref = deltas.clone().requires_grad_(True)
        (oracle.create_corrected_poses(P, d_ref) * G).sum().backward()
        d_gpu = deltas.clone().to(dev).requires_grad_(True)
        out = ops.pose_compose(P.to(dev), d_gpu)
        (out * G.to(dev)).sum().backward()
        assert (out.detach().cpu() - oracle.create_corrected_poses(P, deltas)).abs().max() < 1e-14
        assert rel_err_norm(d_gpu.grad.cpu().numpy(), d_ref.grad.numpy()) < 1e-12


# ------------------------------------------------------------------------------------------------
# kernels 2 + 3: the fused training step against the reference's goldens (float64 clouds)
# ------------------------------------------------------------------------------------------------
def _build_step(dc, dev, g, dtype, own_search):
    inp = step_inputs(g)
    S = len(inp['scans'])
    clouds = []
    for sc in inp['scans']:
        if dtype == torch.float64:
            c = dc.DepthCloud.from_points(torch.as_tensor(sc['points32'].astype(np.float64), device=dev))
        else:
            c = dc.DepthCloud.from_points(torch.as_tensor(sc['points32'], device=dev))
        c.inc_angles = sc['inc_angles'].to(dev).to(dtype)
        c.mask = sc['mask'].to(dev)
        clouds.append(c)
    k, r = int(g['nn_k']), float(g['nn_r'])
    cfg = dc.Config(nn_k=k, nn_r=r or None, pose_correction=str(g['pose_correction']))
    poses = inp['poses'].to(dev)
    if own_search:
        ns = dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg)
    else:
        nb = inp['neighbors'].to(dev)
        ns = (nb, (nb >= 0).float()[..., None])
        ns[1]._dc_is_mask = True
    Model = dc.ScaledPolynomial if inp['scaled'] else dc.Polynomial
    model = Model(w=inp['w'].flatten().tolist(), exponent=inp['exponent'].flatten().tolist(), device=dev)
    deltas = None
    if inp['pose_deltas'] is not None:
        deltas = inp['pose_deltas'].to(dev).requires_grad_(True)
        if deltas.shape[0] == 1:     # common / sequence correction: lists over sequences (train.py:225)
            poses_c = dc.create_corrected_poses([poses], [deltas], cfg)[0]
        else:
            poses_c = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
    else:
        poses_c = poses.clone().requires_grad_(True)
    if deltas is not None:
        poses_c.retain_grad()
    cloud = dc.global_cloud(clouds=clouds, model=model, poses=poses_c)
    feats = dc.compute_neighborhood_features(cloud=cloud, model=None, neighborhoods=ns, cfg=cfg)
    return inp, clouds, ns, model, deltas, poses_c, feats


def _run_loss(dc, inp, feats, dev, **extra):
    mask = None if inp['loss_mask'] is None else inp['loss_mask'].to(dev)
    kw = dict(sqrt=inp['sqrt'], reduction=dc.Reduction(inp['reduction']), **extra)
    if inp['loss'] == 'min_eigval_loss':
        return dc.min_eigval_loss(feats, mask=mask, normalization=inp['normalization'], **kw)
    return dc.trace_loss(feats, mask=mask, **kw)


@pytest.mark.parametrize('own_search', [True, False])
@pytest.mark.parametrize('tag', STEP_TAGS)
def test_fused_step_vs_reference_golden(dc, dev, golden, tag, own_search):
    g = golden('step_' + tag)
    inp, clouds, ns, model, deltas, poses_c, feats = _build_step(dc, dev, g, torch.float64, own_search)
    if own_search:
        nb = ns[0].cpu().numpy()
        assert nb.shape == g['neighbors'].shape and np.array_equal(nb, g['neighbors'])
    assert feats.fusable()
    if not own_search:
        feats._graph.transposed()      # force the gather-form backward; a fresh kNN graph starts with the scatter form
    loss, loss_cloud = _run_loss(dc, inp, feats, dev)
    loss.backward()
    assert own_search or feats._graph.symmetric or feats._graph._transposed is not None
    # sqrt variants amplify the +-1e-17 eigenvalue noise of rank-deficient neighbourhoods (sqrt'(x) ~ 1e9) in the
    # reference itself (its own CPU re-run differs from the golden by ~5e-7), so they get the north-star tolerance
    tol, gtol = (RTOL, RTOL) if inp['sqrt'] and inp['loss'] == 'min_eigval_loss' else (1e-9, 1e-8)
    assert rel_err(loss.item(), g['loss']) < tol
    assert rel_err_norm(loss_cloud.loss.cpu().numpy(), g['per_point']) < max(tol, 1e-7)
    assert rel_err_norm(model.w.grad.cpu().numpy(), g['w_grad']) < gtol
    assert rel_err_norm(poses_c.grad.cpu().numpy(), g['poses_grad']) < gtol
    if deltas is not None:
        assert rel_err_norm(deltas.grad.cpu().numpy(), g['pose_deltas_grad']) < gtol
    # lazily materialised features of the same cloud agree with the reference's update_all
    assert eig_close(feats.eigvals.detach().cpu().numpy(), g['eigvals'], 1e-8)
    assert np.max(np.abs(feats.points.detach().cpu().numpy() - g['points'])) < 1e-12
    assert np.max(np.abs(feats.cov.detach().cpu().numpy() - g['cov'])) < 1e-13


@pytest.mark.parametrize('tag', STEP_TAGS)
def test_fused_step_float32_storage_vs_oracle(dc, dev, golden, tag):
    """The bench configuration: float32 records; the oracle consumes the same float32 values in fp64."""
    from oracle import oracle
    g = golden('step_' + tag)
    inp, clouds, ns, model, deltas, poses_c, feats = _build_step(dc, dev, g, torch.float32, True)
    scans = [{'vps': c.vps.double().cpu().expand(len(c), 3), 'dirs': c.dirs.double().cpu(), 'depth': c.depth.double().cpu(),
              'inc_angles': c.inc_angles.double().cpu(), 'mask': c.mask.cpu()} for c in clouds]
    pts0, _ = oracle.global_points(scans, inp['poses'])
    _, nb_ref = oracle.nearest_neighbors(pts0, k=int(g['nn_k']) or None, r=float(g['nn_r']) or None)
    assert torch.equal(ns[0].cpu(), nb_ref)
    ref = oracle.map_consistency_step(scans, inp['poses'], nb_ref, inp['w'], inp['exponent'], pose_deltas=inp['pose_deltas'],
                                      loss_mask=inp['loss_mask'], loss=inp['loss'], scaled=inp['scaled'],
                                      normalization=inp['normalization'], sqrt=inp['sqrt'], reduction=inp['reduction'])
    loss, loss_cloud = _run_loss(dc, inp, feats, dev)
    loss.backward()
    assert rel_err(loss.item(), ref['loss'].item()) < RTOL
    assert rel_err_norm(loss_cloud.loss.cpu().numpy(), ref['per_point'].numpy()) < RTOL
    assert rel_err_norm(model.w.grad.cpu().numpy(), ref['w_grad'].numpy()) < RTOL
    assert rel_err_norm(poses_c.grad.cpu().numpy(), ref['poses_grad'].numpy()) < RTOL
    if deltas is not None:
        assert rel_err_norm(deltas.grad.cpu().numpy(), ref['pose_deltas_grad'].numpy()) < RTOL


def test_general_path_inliers_and_none_reduction(dc, dev, golden):
    """inlier_ratio / reduction none go through the per-point kernel output and per-point upstream gradient."""
    from oracle import oracle
    g = golden('step_scaled_mineig_norm_r')
    inp, clouds, ns, model, deltas, poses_c, feats = _build_step(dc, dev, g, torch.float64, False)
    mask = inp['loss_mask'].to(dev)
    loss, lc = dc.min_eigval_loss(feats, mask=mask, normalization=True, inlier_ratio=0.8)
    loss.backward()
    for sc in inp['scans']:
        sc.pop('points32')
    w = inp['w'].clone().requires_grad_(True)
    d = inp['pose_deltas'].clone().requires_grad_(True)
    pts, _ = oracle.global_points(inp['scans'], oracle.create_corrected_poses(inp['poses'], d), w, inp['exponent'], True)
    f = oracle.neighborhood_features(pts, inp['neighbors'], eigvecs=False)
    ref, _ = oracle.min_eigval_loss(f['eigvals'], inp['loss_mask'], normalization=True, inlier_ratio=0.8)
    ref.backward()
    assert rel_err(loss.item(), ref.item()) < 1e-9
    assert rel_err_norm(model.w.grad.cpu().numpy(), w.grad.numpy()) < 1e-8
    assert rel_err_norm(deltas.grad.cpu().numpy(), d.grad.numpy()) < 1e-8
    model.w.grad = None
    feats2 = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=poses_c.detach()),
                                              neighborhoods=ns, cfg=dc.Config(nn_r=0.4))
    pp, _ = dc.min_eigval_loss(feats2, mask=mask, normalization=True, reduction=dc.Reduction.NONE)
    assert pp.shape == (int(mask.sum()),)
    assert rel_err_norm(pp.detach().cpu().numpy(), g['per_point']) < 1e-9


def test_batch_of_clouds_and_invariances(dc, dev, golden):
    g = golden('step_scaled_mineig_norm_r')
    inp, clouds, ns, model, deltas, poses_c, feats = _build_step(dc, dev, g, torch.float64, False)
    mask = inp['loss_mask'].to(dev)
    single, _ = dc.min_eigval_loss(feats, mask=mask, normalization=True)
    feats_b = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=poses_c.detach()),
                                               neighborhoods=ns, cfg=dc.Config(nn_r=0.4))
    both, lcs = dc.min_eigval_loss([feats, feats_b], mask=[mask, mask], normalization=True)
    assert len(lcs) == 2 and abs(both.item() - single.item()) < 1e-15
    # rigid motion of the whole map leaves eigenvalue losses unchanged (graph is fixed)
    from oracle import oracle
    M = oracle.xyz_axis_angle_to_matrix(torch.tensor([[3.0, -2.0, 1.0, 0.3, -0.5, 0.2]], dtype=torch.float64))[0].to(dev)
    moved = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=M @ poses_c.detach()),
                                             neighborhoods=ns, cfg=dc.Config(nn_r=0.4))
    l2, _ = dc.min_eigval_loss(moved, mask=mask, normalization=True)
    assert abs(l2.item() - single.item()) < 1e-11 * abs(single.item()) + 1e-15
    # mask all false -> mean over nothing is NaN like torch's mean of an empty tensor
    none, _ = dc.min_eigval_loss(moved, mask=torch.zeros_like(mask), normalization=True)
    assert torch.isnan(none)


def test_local_feature_cloud_vs_reference(dc, dev, golden):
    g = golden('step_scaled_mineig_norm_r')
    cfg = dc.Config(nn_k=0, nn_r=0.4)
    for i in range(int(g['n_scans'])):
        c = dc.DepthCloud.from_points(torch.as_tensor(g['scan%d_points' % i].astype(np.float64), device=dev))
        c = dc.local_feature_cloud(c, cfg)
        flips = (c.mask.cpu().numpy() != g['scan%d_mask' % i]).mean()
        assert flips < 2e-3
        inc = c.inc_angles.cpu().numpy()
        ok = g['scan%d_mask' % i] & c.mask.cpu().numpy()
        assert np.max(np.abs(inc[ok] - g['scan%d_inc_angles' % i][ok])) < 1e-6


# ------------------------------------------------------------------------------------------------
# per-scan preprocessing filters
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_filter_grid_matches_reference(dc, dev, golden, dtype):
    """Same survivors in the same order as the reference's dict walk, for every keep mode / order."""
    g = golden('filters')
    tag = 'f32' if dtype == np.float32 else 'f64'
    x = torch.as_tensor(g['points'].astype(dtype), device=dev)
    for keep in ('first', 'last', 'random'):
        for po in (False, True):
            ind = dc.filter_grid(x, 0.2, only_mask=True, keep=keep, preserve_order=po, rng=np.random.default_rng(135))
            assert ind.dtype == torch.int64
            assert np.array_equal(ind.cpu().numpy(), g['grid_%s_%s_%d' % (tag, keep, int(po))]), (keep, po)
    rng = np.random.default_rng(7)            # a stateful generator advances exactly like the reference's
    x32 = torch.as_tensor(g['points'], device=dev)
    assert np.array_equal(dc.filter_grid(x32, 0.35, only_mask=True, keep='random', rng=rng).cpu().numpy(), g['grid_two_draws_a'])
    assert np.array_equal(dc.filter_grid(x32, 0.35, only_mask=True, keep='random', rng=rng).cpu().numpy(), g['grid_two_draws_b'])
    # DepthCloud in -> DepthCloud out (slicing keeps the source fields)
    cloud = dc.DepthCloud.from_points(x32)
    kept = dc.filter_grid(cloud, 0.2, keep='first', preserve_order=True)
    assert len(kept) == len(g['grid_f32_first_1'])


def test_filter_grid_full_scan_vs_oracle(dc, dev):
    """Full-resolution scan (131 k points, hundreds of points per voxel near the sensor), negative coordinates."""
    from oracle import oracle
    from depth_correction_b200.synthetic import make_sequence
    scans, _, _ = make_sequence('corridor', n_scans=1, pattern='os0-128', seed=2)
    pts = scans[0]['points']
    x = torch.as_tensor(pts, device=dev)
    for keep, po, res in (('random', False, 0.2), ('first', False, 0.1), ('last', True, 0.5)):
        ours = dc.filter_grid(x, res, only_mask=True, keep=keep, preserve_order=po, rng=np.random.default_rng(3))
        ref = oracle.filter_grid(pts, np.float32(res), keep=keep, preserve_order=po, rng=np.random.default_rng(3))
        assert np.array_equal(ours.cpu().numpy(), ref), (keep, po, res)
    with pytest.raises(ValueError):
        dc.filter_grid(torch.full((4, 3), float('nan'), device=dev), 0.2, only_mask=True)
    assert len(dc.filter_grid(torch.zeros((0, 3), device=dev), 0.2, only_mask=True)) == 0


def test_shadow_filter_matches_reference(dc, dev, golden):
    g = golden('filters')
    sp = torch.as_tensor(g['shadow_points'].astype(np.float64), device=dev)
    for tag, bounds in (('a', [0.0873, None]), ('b', [0.2, 2.8])):
        cloud = dc.DepthCloud.from_points(sp)
        cloud.update_dir_neighbors(angle=0.02)
        assert cloud.dir_neighbors.shape[1] == int(g['shadow_K_' + tag])
        mask = dc.filter_shadow_points(cloud, list(bounds), only_mask=True)
        assert np.array_equal(torch.nonzero(mask)[:, 0].cpu().numpy(), g['shadow_kept_' + tag]), tag
        assert len(dc.filter_shadow_points(cloud, list(bounds))) == len(g['shadow_kept_' + tag])
    # through local_feature_cloud (preproc.py:44-47)
    cfg = dc.Config(nn_k=0, nn_r=0.3, shadow_angle_bounds=[0.0873, None], shadow_neighborhood_angle=0.02,
                    eigenvalue_ratio_bounds=[])
    out = dc.local_feature_cloud(dc.DepthCloud.from_points(sp), cfg)
    assert len(out) == len(g['shadow_kept_a']) and out.eigvals is not None


def test_global_cloud_statistics_and_mask_match_reference(dc, dev, golden):
    """vp / dir dispersion, mean depth, mean viewpoint distance (depth_cloud.py:314-354) and global_cloud_mask
    (preproc.py:122-164) of a three-scan global cloud against the reference."""
    g = golden('stats')
    S = int(g['n_scans'])
    cfg = dc.Config(nn_k=0, nn_r=0.4, min_depth=0.0, grid_res=0.0, float_type='float64')
    clouds = [dc.local_feature_cloud(dc.DepthCloud.from_points(torch.as_tensor(g['scan%d_points' % i].astype(np.float64), device=dev)), cfg)
              for i in range(S)]
    poses = torch.as_tensor(g['poses'], device=dev)
    cloud = dc.global_cloud(clouds=clouds, model=None, poses=poses)
    cloud.update_all(r=0.4)
    for f in ('vp_dispersion', 'dir_dispersion', 'mean_depth', 'mean_vp_dist', 'vp_dispersion_to_depth2', 'vp_dist_to_depth'):
        ours = getattr(cloud, f)().cpu().numpy()
        assert ours.shape == g[f].shape, f
        assert np.allclose(ours, g[f], rtol=1e-9, atol=1e-13), (f, np.abs(ours - g[f]).max())
    mcfg = dc.Config(nn_k=0, nn_r=0.4, min_valid_neighbors=8, eigenvalue_bounds=[[0, None, 0.01]],
                     eigenvalue_ratio_bounds=[[0, 1, 0, 0.25], [1, 2, 0.25, 1.0]], dir_dispersion_bounds=[0.0, 0.02],
                     vp_dispersion_bounds=[0.05, float('inf')], vp_dispersion_to_depth2_bounds=[0.001, None])
    # lambda_0 / lambda_1 >= 0 fails for the rounding-noise lambda_0 < 0 of exactly planar (<= 3 point) neighbourhoods,
    # whose sign is solver noise in the reference (LAPACK) as much as here: compare where lambda_0 is resolved
    lev = torch.cat([c.eigvals for c in clouds]).cpu().numpy()
    resolved_local = np.abs(lev[:, 0]) > 1e-12 * lev[:, 2]
    assert np.array_equal(cloud.mask.cpu().numpy()[resolved_local], g['mask_start'][resolved_local])
    assert resolved_local.mean() > 0.9
    assert eig_close(cloud.eigvals.cpu().numpy(), g['eigvals'], 1e-9)
    assert np.array_equal(dc.filter_valid_neighbors(cloud, min=8, only_mask=True).cpu().numpy(), g['mask_valid'])
    assert np.array_equal(dc.filter_eigenvalues(cloud, mcfg.eigenvalue_bounds, only_mask=True).cpu().numpy(), g['mask_eig'])
    ratio = dc.filter_eigenvalue_ratios(cloud, mcfg.eigenvalue_ratio_bounds, only_mask=True).cpu().numpy()
    resolved = np.abs(g['eigvals'][:, 0]) > 1e-12 * g['eigvals'][:, 2]
    assert np.array_equal(ratio[resolved], g['mask_ratio'][resolved])
    mask = dc.global_cloud_mask(cloud, cloud.mask.clone(), mcfg).cpu().numpy()
    both = resolved & resolved_local
    assert np.array_equal(mask[both], g['mask'][both])
    assert both.mean() > 0.9 and g['mask'][both].sum() > 500


# ------------------------------------------------------------------------------------------------
# ICP-style losses
# ------------------------------------------------------------------------------------------------
def _our_icp(dc, dev, g, point_to_plane, use_masks, dtype=torch.float64):
    scans, poses, w, exponent, masks = icp_inputs(g)
    clouds = []
    for s in scans:
        c = dc.DepthCloud.from_points(s['points'].to(device=dev, dtype=dtype))
        c.inc_angles = s['inc_angles'].to(device=dev, dtype=dtype)
        c.mask = s['mask'].to(dev)
        c.normals = s['normals'].to(device=dev, dtype=dtype)
        clouds.append(c)
    model = dc.ScaledPolynomial(w=w.reshape(-1).tolist(), exponent=exponent.reshape(-1).tolist(), device=dev)
    poses_t = poses.to(dev).clone().requires_grad_(True)
    mk = [[(a.to(dev), b.to(dev)) for a, b in masks]] if use_masks else None
    loss, loss_clouds = dc.icp_loss([clouds], poses=[list(poses_t)], model=model, masks=mk, icp_point_to_plane=point_to_plane,
                                    icp_inlier_ratio=0.5)
    loss.backward()
    assert len(loss_clouds) == 1 and len(loss_clouds[0]) == sum(len(c) for c in clouds)
    return loss.item(), model.w.grad.cpu().numpy(), poses_t.grad.cpu().numpy()


@pytest.mark.parametrize('tag,p2pl,use_masks', [('plane', True, False), ('point', False, False), ('masked_plane', True, True)])
def test_icp_loss_matches_reference(dc, dev, golden, tag, p2pl, use_masks):
    """icp_loss forward + backward (model weights, poses) against the reference, which evaluates the residuals in
    float32 (loss.py:424-425; its own rounding is ~1e-6 relative), and against the fp64 oracle."""
    g = golden('icp')
    loss, gw, gp = _our_icp(dc, dev, g, p2pl, use_masks)
    assert abs(loss - float(g[tag + '_loss'])) <= RTOL * abs(float(g[tag + '_loss']))
    assert rel_err_norm(gw, g[tag + '_w_grad']) < 1e-4
    assert rel_err_norm(gp, g[tag + '_poses_grad']) < 1e-4
    from test_oracle import _oracle_icp
    lo, gwo, gpo = _oracle_icp(g, p2pl, use_masks)
    assert abs(loss - lo) <= RTOL * abs(lo)
    assert rel_err_norm(gw, gwo) < RTOL and rel_err_norm(gp, gpo) < RTOL
    # float32 clouds (the storage type of the B200 path)
    l32, gw32, gp32 = _our_icp(dc, dev, g, p2pl, use_masks, dtype=torch.float32)
    assert abs(l32 - lo) <= 1e-4 * abs(lo)


def test_icp_by_name_and_quantile(dc, dev):
    from depth_correction_b200.icp import nanquantile
    assert dc.loss_by_name('icp_loss') is dc.icp_loss
    rng = np.random.default_rng(3)
    x = rng.random(1001)
    x[::50] = np.nan
    xt = torch.as_tensor(x, device=dev)
    for q in (0.0, 0.25, 0.5, 0.9, 1.0):
        assert nanquantile(xt, q).item() == torch.nanquantile(torch.as_tensor(x), q).item()
    assert torch.isnan(nanquantile(torch.full((5,), float('nan'), device=dev, dtype=torch.float64), 0.5))


# ------------------------------------------------------------------------------------------------
# training driver
# ------------------------------------------------------------------------------------------------
def test_train_loop_follows_the_oracle(dc, dev, tmp_path):
    """train() (train.py:46-327 semantics): five Adam iterations on a three-scan sequence with an injected depth
    bias, model + per-pose corrections, first pose frozen; the loss trajectory and the final parameters follow the
    CPU oracle driven by the same optimiser, and the reference's checkpoint files appear."""
    from oracle import oracle
    from depth_correction_b200.synthetic import make_sequence
    scans_np, _, poses = make_sequence('corridor', n_scans=3, pattern='os0-32', seed=31, grid_res=0.2, step=0.8,
                                       pose_noise=(0.01, 0.005), bias_w=[-0.01], bias_exponent=[4.0], depth_clip=(1.0, 6.0))
    cfg = dc.Config(nn_k=0, nn_r=0.4, min_depth=0.0, max_depth=float('inf'), grid_res=0.0, float_type='float64',
                    model_kwargs={'w': [0.0, 0.0], 'exponent': [2.0, 4.0]}, pose_correction=dc.PoseCorrection.pose,
                    vp_dispersion_bounds=[], min_valid_neighbors=5, lr=1e-3, n_opt_iters=5, log_dir=str(tmp_path),
                    loss_kwargs={'sqrt': False, 'normalization': True})
    ds = [(torch.as_tensor(s['points'].astype(np.float64), device=dev), T) for s, T in zip(scans_np, poses)]
    to_cloud = lambda seq: [(dc.DepthCloud.from_points(p), T) for p, T in seq]
    seen = {}

    class Capture(dc.TrainCallbacks):
        def train_loss(self, iter, model, clouds, pose_deltas, poses, masks, loss):
            if iter == 0:
                seen['loss_mask'] = masks[0].cpu()
                seen['scan_masks'] = [c.mask.cpu() for c in clouds[0]._scans]

    best = dc.train(cfg, callbacks=Capture(cfg), train_datasets=[to_cloud(ds)], val_datasets=[to_cloud(ds[:2])])
    assert best is not None and os.path.exists(os.path.join(str(tmp_path), 'best.yaml'))
    assert os.path.exists(best.model_state_dict) and os.path.exists(best.train_pose_deltas)
    ours = np.array(best.loss_history)

    # oracle: same features and graph, same Adam; the masks are the driver's (their parity with the reference is
    # covered by test_global_cloud_statistics_and_mask_match_reference -- rank-deficient neighbourhoods may flip)
    scans = []
    for k, s in enumerate(scans_np):
        p64 = torch.as_tensor(s['points'].astype(np.float64))
        vps, dirs, depth = oracle.from_points(p64)
        _, nb = oracle.nearest_neighbors(p64, r=0.4)
        f = oracle.neighborhood_features(p64, nb, dirs=dirs)
        mask = oracle.eigenvalue_masks(f['eigvals'], (), [[0, 1, 0, 0.25], [1, 2, 0.25, 1.0]])
        assert (mask != seen['scan_masks'][k]).float().mean() < 0.01
        scans.append({'vps': vps, 'dirs': dirs, 'depth': depth, 'inc_angles': f['inc_angles'], 'mask': seen['scan_masks'][k]})
    poses_t = torch.as_tensor(poses)
    pts0, _ = oracle.global_points(scans, poses_t)
    _, nb = oracle.nearest_neighbors(pts0, r=0.4)
    f0 = oracle.neighborhood_features(pts0, nb)
    lmask = torch.cat([s['mask'] for s in scans]) & oracle.valid_neighbor_mask(nb, 5) \
        & oracle.eigenvalue_masks(f0['eigvals'], (), [[0, 1, 0, 0.25], [1, 2, 0.25, 1.0]])
    assert (lmask != seen['loss_mask']).float().mean() < 0.01
    lmask = seen['loss_mask']
    w = torch.zeros((1, 2), dtype=torch.float64, requires_grad=True)
    deltas = torch.zeros((3, 6), dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([{'params': [w], 'lr': 1e-3}, {'params': [deltas], 'lr': 1e-3}])
    ref_losses = []
    for it in range(5):
        out = oracle.map_consistency_step(scans, poses_t, nb, w.detach(), torch.tensor([[2.0, 4.0]], dtype=torch.float64),
                                          pose_deltas=deltas.detach(), loss_mask=lmask, loss='min_eigval_loss', normalization=True)
        ref_losses.append(float(out['loss']))
        opt.zero_grad()
        w.grad = out['w_grad'].reshape(1, 2).clone()
        deltas.grad = out['pose_deltas_grad'].clone()
        deltas.grad[0].zero_()
        opt.step()
    assert np.allclose(ours[:, 0], ref_losses, rtol=1e-6), (ours[:, 0], ref_losses)
    sd = torch.load(best.model_state_dict)
    assert ours[-1, 0] < ours[0, 0]                     # the loss goes down
    # parameters saved at the best iteration are the ones that produced its loss
    it_best = int(os.path.basename(best.model_state_dict).split('_')[0])
    assert 0 <= it_best < 5 and sd['w'].shape == (1, 2)


def test_scatter_f32_agrees_with_gather_form(dc, dev, monkeypatch):
    """The three backward forms (fp64 gather over the transposed graph, fp64 scatter, fp32 vector-reduction scatter)
    on a 0.5 M point map: fp64 forms agree to rounding, the fp32 form stays far inside the 1e-5 budget."""
    from depth_correction_b200 import fused
    from depth_correction_b200.synthetic import make_sequence
    scans_np, _, poses = make_sequence('corridor', n_scans=4, pattern='os0-128', seed=4)
    cfg = dc.Config(nn_k=16, nn_r=0.4, pose_correction=dc.PoseCorrection.pose)
    clouds = []
    for s in scans_np:
        c = dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev))
        c.inc_angles = torch.rand((len(c), 1), device=dev) * 1.2          # any fixed per-point constants will do
        clouds.append(c)
    poses_t = torch.as_tensor(poses, device=dev)
    ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)

    def grads(mode):
        monkeypatch.setattr(fused, 'BACKWARD_FORM', 'gather' if mode == 'gather' else 'scatter')
        monkeypatch.setenv('DC_SCATTER_F32', '1' if mode in ('f32', 'f32-two-kernels') else '0')
        monkeypatch.setenv('DC_FUSE_BWD', '0' if mode == 'f32-two-kernels' else '1')
        model = dc.ScaledPolynomial(w=[0.003, -0.002], exponent=[2, 4], device=dev)
        deltas = torch.full((len(clouds), 6), 1e-3, dtype=torch.float64, device=dev, requires_grad=True)
        pc = torch.stack(dc.create_corrected_poses(poses_t, deltas, cfg))
        cloud = dc.global_cloud(clouds=clouds, model=model, poses=pc)
        feats = dc.compute_neighborhood_features(cloud=cloud, neighborhoods=ns, cfg=cfg)
        loss, _ = dc.min_eigval_loss(feats, normalization=True)
        loss.backward()
        return model.w.grad.cpu().numpy().ravel(), deltas.grad.cpu().numpy()

    gw64, gd64 = grads('f64')
    gw32, gd32 = grads('f32')           # forward + scatter in ONE kernel (dc_step_forward_scatter)
    gw2k, gd2k = grads('f32-two-kernels')
    gwg, gdg = grads('gather')          # last: builds the transposed graph
    assert rel_err_norm(gw64, gwg) < 1e-11 and rel_err_norm(gd64, gdg) < 1e-11
    for gw, gd in ((gw32, gd32), (gw2k, gd2k)):
        assert rel_err_norm(gw, gwg) < 1e-6, rel_err_norm(gw, gwg)
        assert rel_err_norm(gd, gdg) < 1e-6, rel_err_norm(gd, gdg)


def test_knn_two_million_points_sampled_vs_ckdtree(dc, dev):
    """Bench-shaped map at 2 M points (16 full-resolution scans): a random sample of 20 000 rows of the kNN graph
    against cKDTree built on the whole map, plus size-independent properties of the full graph: every row holds k
    distinct indices or is padded, the search is idempotent, and a radius graph is symmetric."""
    from oracle import oracle
    from depth_correction_b200.graph import search
    from depth_correction_b200.synthetic import make_sequence
    scans, poses, _ = make_sequence('corridor', n_scans=16, pattern='os0-128', seed=0)
    pts = np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)])
    p = torch.as_tensor(pts, device=dev)
    g = search(p, None, k=32, r=0.4)
    nb = g.neighbors()
    dist = g.distances()
    rows = torch.as_tensor(np.random.default_rng(1).choice(len(pts), 20000, replace=False), device=dev)
    d_ref, i_ref = oracle.nearest_neighbors(torch.as_tensor(pts), torch.as_tensor(pts[rows.cpu().numpy()]), k=32, r=0.4)
    assert torch.equal(nb[rows].cpu(), i_ref)
    assert torch.equal(dist[rows].cpu(), d_ref)
    # properties of the whole graph
    srt = nb.sort(dim=1).values
    assert ((srt[:, 1:] != srt[:, :-1]) | (srt[:, 1:] < 0)).all()                   # no index twice in a row
    assert (nb[:, 0] == torch.arange(len(pts), device=dev)).all()                    # nearest neighbour of a point is itself
    assert (dist[:, 1:] >= dist[:, :-1]).all()                                       # rows sorted by distance
    assert torch.equal(search(p, None, k=32, r=0.4).neighbors(), nb)                 # idempotent
    sub = p[:400000]
    gr = search(sub, None, r=0.03)
    a = gr.neighbors()
    i = torch.arange(len(sub), device=dev)[:, None].expand_as(a)[a >= 0]
    j = a[a >= 0]
    fwd = torch.stack([i, j], 1)
    bwd = torch.stack([j, i], 1)
    key = lambda e: (e[:, 0] * len(sub) + e[:, 1]).sort().values
    assert torch.equal(key(fwd), key(bwd))                                           # j in N(i)  <=>  i in N(j)


def _edge_scene(n_per_scan, seed):
    """Two tiny scans of a noisy plane 3 m in front of the sensor (float32 values)."""
    rng = np.random.default_rng(seed)
    scans = []
    for s in range(2):
        y, z = rng.uniform(-0.5, 0.5, n_per_scan), rng.uniform(-0.5, 0.5, n_per_scan)
        x = 3.0 + 0.3 * y + 0.01 * rng.standard_normal(n_per_scan)
        scans.append(np.stack([x, y, z], 1).astype(np.float32))
    poses = np.stack([np.eye(4), np.eye(4)])
    poses[1, :3, 3] = [0.02, -0.01, 0.03]
    return scans, poses


@pytest.mark.parametrize('case', ['k1', 'tiny', 'mask_all_false', 'isolated'])
def test_fused_step_edge_cases_vs_oracle(dc, dev, case):
    """Edge cases of the fused step against the oracle: k = 1 (every neighbourhood is the point itself: zero
    covariance through the Bessel clamp), a cloud smaller than one slice, a loss mask that keeps nothing (mean over
    an empty set is NaN in the reference too), and points without any neighbour inside r."""
    from oracle import oracle
    n_per = 5 if case == 'tiny' else 300
    scans_np, poses = _edge_scene(n_per, 3)
    if case == 'isolated':
        scans_np[1][:40] += np.float32(50.0) * (1 + np.arange(40, dtype=np.float32))[:, None]     # far from everything
    kw = dict(k=1, r=None) if case == 'k1' else dict(k=None, r=0.25)
    cfg = dc.Config(nn_k=kw['k'] or 0, nn_r=kw['r'], pose_correction=dc.PoseCorrection.pose)
    rng = np.random.default_rng(8)
    inc = [rng.uniform(0.1, 1.2, (len(s), 1)) for s in scans_np]
    clouds, oscans = [], []
    for s, a in zip(scans_np, inc):
        c = dc.DepthCloud.from_points(torch.as_tensor(s.astype(np.float64), device=dev))
        c.inc_angles = torch.as_tensor(a, device=dev)
        clouds.append(c)
        vps, dirs, depth = oracle.from_points(torch.as_tensor(s.astype(np.float64)))
        oscans.append({'vps': vps, 'dirs': dirs, 'depth': depth, 'inc_angles': torch.as_tensor(a), 'mask': torch.ones(len(s), dtype=torch.bool)})
    n = sum(len(s) for s in scans_np)
    poses_t = torch.as_tensor(poses, device=dev)
    deltas = torch.as_tensor(rng.normal(0, 1e-3, (2, 6)), device=dev).requires_grad_(True)
    model = dc.ScaledPolynomial(w=[0.004, -0.003], exponent=[2, 4], device=dev)
    ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)
    pc = torch.stack(dc.create_corrected_poses(poses_t, deltas, cfg))
    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
    mask = torch.zeros(n, dtype=torch.bool, device=dev) if case == 'mask_all_false' else None
    loss, _ = dc.min_eigval_loss(feats, mask=mask, normalization=True)
    loss.backward()
    pts0, _ = oracle.global_points(oscans, torch.as_tensor(poses))
    _, nb = oracle.nearest_neighbors(pts0, k=kw['k'], r=kw['r'])
    assert torch.equal(ns[0].cpu(), nb)
    ref = oracle.map_consistency_step(oscans, torch.as_tensor(poses), nb, model.w.detach().cpu(), model.exponent.cpu(),
                                      pose_deltas=deltas.detach().cpu(), loss_mask=None if mask is None else mask.cpu(),
                                      loss='min_eigval_loss', normalization=True)
    if case == 'mask_all_false':
        assert torch.isnan(loss) and torch.isnan(ref['loss'])
        return
    assert abs(loss.item() - ref['loss'].item()) <= 1e-9 * abs(ref['loss'].item()) + 1e-18
    if case == 'k1':
        assert loss.item() == 0.0 and model.w.grad.abs().max() == 0 and deltas.grad.abs().max() == 0
        return
    assert rel_err_norm(model.w.grad.cpu().numpy(), ref['w_grad'].numpy()) < 1e-8
    assert rel_err_norm(deltas.grad.cpu().numpy(), ref['pose_deltas_grad'].numpy()) < 1e-8


def test_fused_step_one_million_points_vs_oracle(dc, dev, monkeypatch):
    """The fused step at map scale (9 full-resolution scans, 1.18 M points, float32 records) against the CPU oracle
    (cKDTree + torch fp64 autograd on the same values): loss, dL/dw and dL/dpose for the deterministic fp64 gather form
    and for the default form of large maps (float32 vector-reduction scatter)."""
    from oracle import oracle
    from depth_correction_b200.synthetic import make_sequence
    scans_np, _, poses = make_sequence('corridor', n_scans=9, pattern='os0-128', seed=12)
    rng = np.random.default_rng(2)
    cfg = dc.Config(nn_k=16, nn_r=0.4, pose_correction=dc.PoseCorrection.pose)
    clouds, oscans = [], []
    for s in scans_np:
        inc = rng.uniform(0.05, 1.3, (len(s['points']), 1)).astype(np.float32)
        msk = rng.random(len(s['points'])) < 0.9
        c = dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev))
        c.inc_angles = torch.as_tensor(inc, device=dev)
        c.mask = torch.as_tensor(msk, device=dev)
        clouds.append(c)
        # the oracle consumes the float32 records the kernels see (dirs / depth as computed in float32), up-cast
        oscans.append({'vps': c.vps.double().cpu(), 'dirs': c.dirs.double().cpu(), 'depth': c.depth.double().cpu(),
                       'inc_angles': torch.as_tensor(inc.astype(np.float64)), 'mask': torch.as_tensor(msk)})
    n = sum(len(c) for c in clouds)
    assert n >= (1 << 20)
    poses_t = torch.as_tensor(poses, device=dev)
    d0 = torch.as_tensor(rng.normal(0, 2e-3, (len(clouds), 6)), device=dev)
    ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)
    pts0, _ = oracle.global_points(oscans, torch.as_tensor(poses))
    _, nb = oracle.nearest_neighbors(pts0, k=16, r=0.4)
    assert torch.equal(ns[0].cpu(), nb)
    ref = oracle.map_consistency_step(oscans, torch.as_tensor(poses), nb, torch.tensor([[0.004, -0.003]], dtype=torch.float64),
                                      torch.tensor([[2.0, 4.0]], dtype=torch.float64), pose_deltas=d0.cpu(),
                                      loss='min_eigval_loss', normalization=True)
    from depth_correction_b200 import fused
    for form, tol in (('gather', 1e-9), ('auto', 1e-6)):
        monkeypatch.setattr(fused, 'BACKWARD_FORM', form)
        model = dc.ScaledPolynomial(w=[0.004, -0.003], exponent=[2, 4], device=dev)
        deltas = d0.clone().requires_grad_(True)
        pc = torch.stack(dc.create_corrected_poses(poses_t, deltas, cfg))
        feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
        loss, _ = dc.min_eigval_loss(feats, normalization=True)
        loss.backward()
        assert abs(loss.item() - ref['loss'].item()) <= 1e-10 * abs(ref['loss'].item()), form
        assert rel_err_norm(model.w.grad.cpu().numpy(), ref['w_grad'].numpy()) < tol, (form, rel_err_norm(model.w.grad.cpu().numpy(), ref['w_grad'].numpy()))
        assert rel_err_norm(deltas.grad.cpu().numpy(), ref['pose_deltas_grad'].numpy()) < tol, form


def test_container_and_pipeline_helpers(dc, dev):
    """Smaller pieces of the reference API that the loops above do not touch directly: pose parametrisation round
    trip, filtered_cloud (depth + seeded voxel filter, preproc.py:25-32) against the oracle, slicing / concatenation
    with neighbour index shifting, neighbour collection."""
    from oracle import oracle
    from depth_correction_b200.synthetic import make_sequence
    rng = np.random.default_rng(4)
    xyzaa = torch.as_tensor(np.concatenate([rng.normal(0, 1, (12, 3)), rng.normal(0, 0.7, (12, 3))], 1), device=dev)
    back = dc.matrix_to_xyz_axis_angle(dc.xyz_axis_angle_to_matrix(xyzaa))
    assert torch.allclose(back, xyzaa, atol=1e-12)
    # filtered_cloud == filter_depth then filter_grid(keep='random', rng = default_rng(cfg.random_seed))
    scans, _, _ = make_sequence('corridor', n_scans=1, pattern='os0-32', seed=13)
    pts = scans[0]['points']
    cfg = dc.Config(min_depth=2.0, max_depth=10.0, grid_res=0.3, random_seed=99)
    out = dc.filtered_cloud(dc.DepthCloud.from_points(torch.as_tensor(pts, device=dev)), cfg)
    depth = np.linalg.norm(pts.astype(np.float32), axis=1)
    kept = np.nonzero((depth >= np.float32(2.0)) & (depth <= np.float32(10.0)))[0]
    # the voxel filter sees the points rebuilt from (vps, dirs, depth) in float32, like the reference
    c_ref = dc.DepthCloud.from_points(torch.as_tensor(pts[kept], device=dev))
    ind = oracle.filter_grid(c_ref.to_points().cpu().numpy(), np.float32(0.3), keep='random', rng=np.random.default_rng(99))
    assert len(out) == len(ind)
    assert torch.equal(out.to_points().cpu(), c_ref.to_points().cpu()[torch.as_tensor(ind)])
    # slicing, concatenation, neighbour collection
    a = dc.DepthCloud.from_points(torch.as_tensor(pts[:500], device=dev))
    b = dc.DepthCloud.from_points(torch.as_tensor(pts[500:900], device=dev))
    a.update_all(r=0.5)
    b.update_all(r=0.5)
    both = a + b
    assert len(both) == 900 and both.neighbors.shape[0] == 900
    nb_b = both.neighbors[500:]
    assert torch.equal(nb_b[nb_b >= 0] - 500, b.neighbors[b.neighbors >= 0]) and (both.neighbors[:500][:, :a.neighbors.shape[1]] == a.neighbors).all()
    sel = torch.arange(10, 20, device=dev)
    idx = a.collect_neighbors(sel)
    expect = torch.unique(a.neighbors[sel][a.neighbors[sel] >= 0])
    assert torch.equal(idx, expect) and len(a.filter_with_neighbors(sel)) == len(expect)
    sub = a[torch.arange(0, 500, 7, device=dev)]
    assert len(sub) == 72 and sub.neighbors is None and sub.eigvals.shape == (72, 3)
