"""GPU tests added in round 2: the two independent kNN kernels against each other, tie-break invariance, the street
(HDL-64) geometry against the oracle, the one-launch feature masks, and the snapshot semantics of the fused step's
caches."""
import numpy as np
import pytest
import torch

from helpers import rel_err_norm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dc():
    import depth_correction_b200 as dc
    return dc


@pytest.fixture(scope='module')
def dev():
    return torch.device('cuda:0')


def _rows_sorted(g):
    k, n = g.k, g.n_rows
    ns = (n + 31) // 32
    return g.ell_idx[:ns * 32 * k].view(ns, k, 32).permute(0, 2, 1).reshape(ns * 32, k)[:n].sort(dim=1).values


def _street_points(n_scans, seed=0):
    from depth_correction_b200.synthetic import make_sequence
    scans, poses, _ = make_sequence('street', n_scans=n_scans, pattern='hdl-64', seed=seed, depth_clip=(5.0, 80.0))
    return scans, poses


@pytest.mark.parametrize('case', ['clustered', 'lattice', 'street', 'cross'])
def test_knn_cell_kernel_equals_thread_kernel(dc, dev, monkeypatch, case):
    """dc_knn (one query per thread, fp64), dc_knn_cells (one warp per cell, fp32 classification + fp64 re-check) and
    dc_knn_recorded (one query per thread, one distance pass) are three implementations of the same selection:
    identical neighbour SETS, row by row."""
    from depth_correction_b200.graph import search
    rng = np.random.default_rng(5)
    query = None
    if case == 'clustered':
        c = rng.uniform(-10, 10, (200, 3))
        pts = (c[rng.integers(0, 200, 120000)] + rng.normal(0, 0.3, (120000, 3))).astype(np.float32)
        pts[500:700] = pts[:200]                            # duplicates
        pts[900:1000] = rng.uniform(-30, 30, (100, 3)).astype(np.float32)   # isolated points
        kws = (dict(k=16, r=0.25), dict(k=8, r=None), dict(k=1, r=None), dict(k=64, r=0.5))
    elif case == 'lattice':
        pts = np.stack(np.meshgrid(np.arange(24), np.arange(24), np.arange(8), indexing='ij'), -1).reshape(-1, 3).astype(np.float32) * 0.25
        kws = (dict(k=7, r=None), dict(k=27, r=0.5))      # exact ties at every k-th place
    elif case == 'street':
        scans, poses = _street_points(4)
        pts = np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32)
        kws = (dict(k=32, r=0.4), dict(k=12, r=None))
    else:
        c = rng.uniform(-5, 5, (50, 3))
        pts = (c[rng.integers(0, 50, 60000)] + rng.normal(0, 0.2, (60000, 3))).astype(np.float32)
        query = torch.as_tensor((pts[::5] + rng.normal(0, 0.05, pts[::5].shape)).astype(np.float32), device=dev)
        kws = (dict(k=4, r=None), dict(k=1, r=0.2))
    p = torch.as_tensor(pts, device=dev)
    for kw in kws:
        monkeypatch.setenv('DC_KNN', 'thread')
        g = search(p, query, **kw)
        a, a_raw = _rows_sorted(g), g.ell_idx.clone()
        monkeypatch.setenv('DC_KNN', 'cells')
        b = _rows_sorted(search(p, query, **kw))
        assert torch.equal(a, b), (case, kw, int((a != b).any(dim=1).sum()))
        # dc_knn_recorded (the default: one distance pass, emit from the recorded bins) walks the rows in the order of
        # dc_knn: the lists are equal entry by entry, not only as sets
        monkeypatch.setenv('DC_KNN', 'record')
        assert torch.equal(search(p, query, **kw).ell_idx, a_raw), (case, kw)


def test_knn_cell_model_changes_speed_not_results(dc, dev, monkeypatch):
    """The cell size of a kNN search comes from a cost model evaluated on a sample of queries (graph._knn_cell_from_sample):
    it may differ from the occupancy estimate, the neighbour lists may not; the map carries the KNN_PAD records the first
    pass of dc_knn_recorded may read behind its last row."""
    from depth_correction_b200 import graph
    scans, poses = _street_points(3)
    pts = np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32)
    p = torch.as_tensor(pts, device=dev)
    assert len(p) >= graph.KNN_MODEL_MIN_POINTS
    cells, rows = {}, {}
    for mode in ('occ', 'model'):
        monkeypatch.setenv('DC_KNN_CELL', mode)
        graph.clear_cell_hints()
        g = graph.search(p, k=32, r=0.4)
        cells[mode] = g.map.cell
        rows[mode] = g.neighbors()
        assert g.map.P.untyped_storage().nbytes() >= (len(p) + graph.KNN_PAD) * 32
        # a second search of the same map takes the remembered cell: no estimate at all
        assert graph.search(p, k=32, r=0.4).map.cell == g.map.cell
    graph.clear_cell_hints()
    assert torch.equal(rows['occ'], rows['model'])
    # the model looks at cells within 2^(-8/6) .. 2 of a first estimate that is itself within 0.77 .. 1.3 of the occupancy cell
    assert cells['occ'] / 3.4 <= cells['model'] <= 2.7 * cells['occ']
    tree_d, tree_i = __import__('scipy.spatial', fromlist=['cKDTree']).cKDTree(pts.astype(np.float64)).query(
        pts[:2000].astype(np.float64), k=32, distance_upper_bound=0.4)
    ref = np.where(np.isfinite(tree_d), tree_i, -1)
    got = rows['model'][:2000].cpu().numpy()
    assert np.array_equal(np.sort(got, axis=1), np.sort(ref, axis=1))


def test_knn_ties_do_not_depend_on_the_cell_size(dc, dev):
    """Exact ties at the k-th place are broken by the ORIGINAL index, so the index matrix is the same for every cell
    size (round 1 broke them by the position in the cell-sorted map)."""
    from depth_correction_b200.graph import search
    pts = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(6), indexing='ij'), -1).reshape(-1, 3).astype(np.float32) * 0.25
    rng = np.random.default_rng(1)
    pts = pts[rng.permutation(len(pts))]                  # original order unrelated to position
    p = torch.as_tensor(pts, device=dev)
    ref = None
    for cell in (0.11, 0.26, 0.7, 1.9):
        g = search(p, None, k=5, r=None, cell=cell)
        nb = g.neighbors()
        if ref is None:
            ref = nb
        assert torch.equal(nb, ref), cell
    # and the chosen tie members are the smallest original indices among the equidistant candidates
    d = g.distances()
    row = 0
    x = torch.as_tensor(pts.astype(np.float64), device=dev)
    dist_all = (x - x[row]).norm(dim=1)
    kth = d[row, -1]
    cand = torch.nonzero(dist_all == kth)[:, 0]
    n_below = int((dist_all < kth).sum())
    chosen = ref[row][n_below:]
    assert torch.equal(chosen.sort().values, cand.sort().values[:5 - n_below])


def test_street_geometry_step_vs_oracle(dc, dev):
    """The KITTI-360-shaped workload (HDL-64, depth clip 5-80 m, rings far apart at range: many rank-deficient
    neighbourhoods, isolated points): neighbour indices identical to cKDTree, loss and gradients of the fused step
    against the oracle on the same float32 records."""
    from oracle import oracle
    scans_np, poses = _street_points(5, seed=2)
    rng = np.random.default_rng(3)
    cfg = dc.Config(nn_k=16, nn_r=0.4, pose_correction=dc.PoseCorrection.pose)
    clouds, oscans = [], []
    for s in scans_np:
        inc = rng.uniform(0.05, 1.3, (len(s['points']), 1)).astype(np.float32)
        msk = rng.random(len(s['points'])) < 0.9
        c = dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev))
        c.inc_angles = torch.as_tensor(inc, device=dev)
        c.mask = torch.as_tensor(msk, device=dev)
        clouds.append(c)
        oscans.append({'vps': c.vps.double().cpu(), 'dirs': c.dirs.double().cpu(), 'depth': c.depth.double().cpu(),
                       'inc_angles': torch.as_tensor(inc.astype(np.float64)), 'mask': torch.as_tensor(msk)})
    poses_t = torch.as_tensor(poses, device=dev)
    d0 = torch.as_tensor(rng.normal(0, 2e-3, (len(clouds), 6)), device=dev)
    ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)
    pts0, _ = oracle.global_points(oscans, torch.as_tensor(poses))
    _, nb = oracle.nearest_neighbors(pts0, k=16, r=0.4)
    assert torch.equal(ns[0].cpu(), nb)
    assert int((nb >= 0).sum(dim=1).min()) <= 3, 'the sample should contain nearly isolated points'
    ref = oracle.map_consistency_step(oscans, torch.as_tensor(poses), nb, torch.tensor([[0.004, -0.003]], dtype=torch.float64),
                                      torch.tensor([[2.0, 4.0]], dtype=torch.float64), pose_deltas=d0.cpu(),
                                      loss='min_eigval_loss', normalization=True)
    model = dc.ScaledPolynomial(w=[0.004, -0.003], exponent=[2, 4], device=dev)
    deltas = d0.clone().requires_grad_(True)
    pc = torch.stack(dc.create_corrected_poses(poses_t, deltas, cfg))
    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc), neighborhoods=ns, cfg=cfg)
    loss, _ = dc.min_eigval_loss(feats, normalization=True)
    loss.backward()
    assert abs(loss.item() - ref['loss'].item()) <= 1e-9 * abs(ref['loss'].item())
    assert rel_err_norm(model.w.grad.cpu().numpy(), ref['w_grad'].numpy()) < 1e-6
    assert rel_err_norm(deltas.grad.cpu().numpy(), ref['pose_deltas_grad'].numpy()) < 1e-6


def test_feature_mask_kernel_equals_torch_comparisons(dc, dev):
    """dc_feature_mask (all bounds of a configuration in one launch) against the reference's torch comparisons
    (filters.py:85-113, 184-254) in both dtypes, with NaN eigenvalues, infinite / missing sides and a starting mask."""
    rng = np.random.default_rng(9)
    n = 5000
    for dt in (torch.float32, torch.float64):
        ev = np.sort(rng.random((n, 3)) ** 3, axis=1)
        ev[::97] = np.nan
        ev[5::131, 0] = 0.0
        eig = torch.as_tensor(ev, device=dev).to(dt)
        nbr = torch.as_tensor(rng.integers(-1, 50, (n, 12)), device=dev)
        cloud = dc.DepthCloud.from_points(torch.as_tensor(rng.normal(size=(n, 3)), device=dev).to(dt))
        cloud.eigvals = eig
        cloud.neighbors = nbr
        start = torch.as_tensor(rng.random(n) < 0.8, device=dev)
        eb = [[0, None, 0.02], [2, 0.05, float('inf')]]
        rb = [[0, 1, 0, 0.25], [1, 2, 0.25, 1.0], [0, 2, float('-inf'), None]]
        got = dc.feature_mask(cloud, eigenvalue_bounds=eb, eigenvalue_ratio_bounds=rb, min_valid_neighbors=5, mask=start)
        want = start & ((nbr >= 0).sum(dim=1) >= 5)
        want = want & (eig[:, 0] <= 0.02) & (eig[:, 2] >= 0.05)
        r01, r12 = eig[:, 0] / eig[:, 1], eig[:, 1] / eig[:, 2]
        want = want & (r01 >= 0) & (r01 <= 0.25) & (r12 >= 0.25) & (r12 <= 1.0)
        assert got.dtype == torch.bool and torch.equal(got, want), dt
        assert torch.equal(dc.filter_eigenvalue_ratios(cloud, rb, only_mask=True), (r01 >= 0) & (r01 <= 0.25) & (r12 >= 0.25) & (r12 <= 1.0))
        assert torch.equal(dc.filter_eigenvalues(cloud, [], only_mask=True), torch.ones(n, dtype=torch.bool, device=dev))
        assert torch.equal(dc.within_bounds(eig[:, 1], min=0.1, max=0.6), (eig[:, 1] >= 0.1) & (eig[:, 1] <= 0.6))
        assert start.sum() > got.sum() > 0          # the starting mask is not modified in place
        assert torch.equal(start, torch.as_tensor(np.asarray(start.cpu()), device=dev))


def test_incidence_angle_of_a_ray_along_the_normal_is_not_nan(dc, dev):
    """|dirs . n| = 1 + rounding must give 0, not NaN (a NaN incidence angle poisons the corrected depth of the point
    and the loss of its whole neighbourhood: 151 NaN terms on the 57 M point street map)."""
    from depth_correction_b200 import ops
    rng = np.random.default_rng(0)
    v = rng.normal(size=(4096, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v32 = torch.as_tensor(v.astype(np.float32), device=dev)
    eigvecs = torch.zeros((len(v), 3, 3), dtype=torch.float32, device=dev)
    eigvecs[:, :, 0] = v32
    normals, inc = ops.normals_and_angles(v32, eigvecs)
    assert not inc.isnan().any()
    assert float(inc.abs().max()) < 1e-3


def test_loss_mask_and_scan_records_are_snapshots(dc, dev):
    """(a) a NEW mask tensor that recycles the address of a freed one must be uploaded again; (b) an in-place
    edit of a scan's depth must be seen by the next fused step (the packed records are rebuilt)."""
    from depth_correction_b200.synthetic import make_sequence
    scans_np, _, poses = make_sequence('corridor', n_scans=3, pattern='os0-32', seed=8, grid_res=0.15)
    cfg = dc.Config(nn_k=0, nn_r=0.4, min_depth=0.0, grid_res=0.0, pose_correction=dc.PoseCorrection.none)
    clouds = [dc.local_feature_cloud(dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev)), cfg) for s in scans_np]
    clouds = [dc.DepthCloud(vps=c.vps, dirs=c.dirs, depth=c.depth.clone(), inc_angles=c.inc_angles, mask=c.mask) for c in clouds]
    poses_t = torch.as_tensor(poses, device=dev)
    model = dc.ScaledPolynomial(w=[0.002, -0.001], exponent=[2, 4], device=dev)
    ns = dc.establish_neighborhoods(clouds=clouds, poses=poses_t, cfg=cfg)
    n = sum(len(c) for c in clouds)

    def loss_with(mask):
        feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=poses_t), neighborhoods=ns, cfg=cfg)
        return dc.min_eigval_loss(feats, mask=mask, normalization=True)[0].item()

    rng = np.random.default_rng(2)
    m1 = torch.as_tensor(rng.random(n) < 0.5, device=dev)
    ref1 = loss_with(m1.clone())
    m2_host = rng.random(n) < 0.5
    l1 = loss_with(m1)
    addr = m1.data_ptr()
    del m1
    m2 = torch.as_tensor(m2_host, device=dev)             # usually lands on the address just freed
    l2 = loss_with(m2)
    ref2 = loss_with(torch.as_tensor(m2_host, device=dev).clone())
    assert l1 == ref1 and l2 == ref2 and l1 != l2, (l1, l2, ref1, ref2, addr == m2.data_ptr())
    # (b) in-place edit of a scan
    base = loss_with(None)
    clouds[1].depth.mul_(1.01)
    edited = loss_with(None)
    fresh_clouds = [dc.DepthCloud(vps=c.vps, dirs=c.dirs, depth=c.depth.clone(), inc_angles=c.inc_angles, mask=c.mask) for c in clouds]
    feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=fresh_clouds, model=model, poses=poses_t), neighborhoods=ns, cfg=cfg)
    fresh = dc.min_eigval_loss(feats, normalization=True)[0].item()
    assert edited == fresh and edited != base


@pytest.mark.parametrize('kw', [dict(nn_k=0, nn_r=0.4), dict(nn_k=12, nn_r=0.5)])
def test_batched_local_features_equal_the_per_scan_loop(dc, dev, kw):
    """local_feature_clouds (one stacked search + one neighbourhood pass for all scans) against
    [local_feature_cloud(c) for c in scans] (preproc.py:35-64 per scan): same neighbourhoods -> eigenvalues, normals and
    incidence angles agree to rounding; masks agree except on rank-deficient neighbourhoods."""
    from depth_correction_b200.synthetic import make_sequence
    scans_np, _, _ = make_sequence('fee', n_scans=5, pattern='os0-128', seed=6, rings=48, azimuths=384, depth_clip=(1.0, 20.0))
    cfg = dc.Config(min_depth=0.0, grid_res=0.0, **kw)
    pts = [torch.as_tensor(s['points'], device=dev) for s in scans_np]
    ref = [dc.local_feature_cloud(dc.DepthCloud.from_points(p), cfg) for p in pts]
    got = dc.local_feature_clouds([dc.DepthCloud.from_points(p) for p in pts], cfg)
    assert len(got) == len(ref)
    n_all = n_diff = 0
    for a, b in zip(got, ref):
        assert a.eigvals.shape == b.eigvals.shape and a.inc_angles.shape == b.inc_angles.shape and a.mask.dtype == torch.bool
        ea, eb = a.eigvals.double(), b.eigvals.double()
        scale = eb[:, 2:3].clamp_min(1e-12)
        assert float(((ea - eb).abs() / scale).max()) < 2e-5                     # relative to the largest eigenvalue (float32 storage)
        planar = (eb[:, 0] < 0.05 * eb[:, 1]) & (eb[:, 1] > 1e-4) & ((b.neighbors >= 0).sum(dim=1) >= 6)      # well-defined normal
        assert planar.float().mean() > 0.3
        assert float((a.inc_angles - b.inc_angles).abs()[planar].max()) < 2e-3
        cos = (a.normals * b.normals).sum(dim=1)[planar]
        assert float(cos.min()) > 1.0 - 1e-5
        assert float((a.mean - b.mean).abs().max()) < 1e-4
        n_all += a.mask.numel()
        n_diff += int((a.mask != b.mask).sum())
    assert n_diff <= 2e-3 * n_all, (n_diff, n_all)


@pytest.mark.parametrize('kw', [dict(nn_k=0, nn_r=0.25), dict(nn_k=10, nn_r=0.4)])
def test_captured_iteration_follows_the_eager_loop(dc, dev, kw):
    """The optimisation iteration of scripts/model_poses_learning:121-135 recorded into a CUDA graph
    (depth_correction_b200/capture.py) takes the same trajectory as the eager loop: same losses, same parameters."""
    from depth_correction_b200.synthetic import make_sequence
    cfg = dc.Config(min_depth=1.0, max_depth=15.0, grid_res=0.1, loss='min_eigval_loss',
                    pose_correction=dc.PoseCorrection.pose, **kw)
    scans, _, poses_init = make_sequence('fee', n_scans=5, pattern='os0-32', seed=9, pose_noise=(0.01, 0.005),
                                         bias_w=[-0.01], bias_exponent=[4.0])
    poses = torch.as_tensor(poses_init, device=dev)

    def setup():
        clouds = dc.local_feature_clouds([dc.filtered_cloud(dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev)), cfg)
                                          for s in scans], cfg)
        model = dc.ScaledPolynomial(w=[0.0, 0.0], exponent=[2, 4], device=dev)
        deltas = torch.zeros((len(clouds), 6), dtype=torch.float64, device=dev, requires_grad=True)
        opt = torch.optim.Adam([{'params': deltas, 'lr': 1e-3}, {'params': model.parameters(), 'lr': 1e-3}], capturable=True)
        ns = dc.establish_neighborhoods(clouds=clouds, poses=poses, cfg=cfg)

        def iteration():
            pc = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
            feats = dc.compute_neighborhood_features(cloud=dc.global_cloud(clouds=clouds, model=model, poses=pc),
                                                     neighborhoods=ns, cfg=cfg)
            loss, _ = dc.min_eigval_loss(feats)
            opt.zero_grad()
            loss.backward()
            deltas.grad[0].zero_()                       # first pose fixed (train.py:281-284)
            opt.step()
            return loss
        return iteration, model, deltas

    it_e, model_e, deltas_e = setup()
    eager = [float(it_e()) for _ in range(9)]
    it_c, model_c, deltas_c = setup()
    step = dc.CapturedIteration(it_c, warmup=3)
    assert step.library_launches >= 5
    captured = [float(x) for x in step.warmup_outputs] + [float(step().clone()) for _ in range(6)]
    assert eager[-1] < eager[0]
    np.testing.assert_allclose(captured, eager, rtol=1e-9, atol=0)
    assert rel_err_norm(model_c.w.detach().cpu(), model_e.w.detach().cpu()) < 1e-8
    assert rel_err_norm(deltas_c.detach().cpu(), deltas_e.detach().cpu()) < 1e-8
    assert step.replays == 6


def test_radius_lists_longer_than_the_staging_tile_vs_ckdtree(dc, dev):
    """Radius mode on a dense map (lists of several hundred entries, lanes of a warp with very different counts): the fill
    pass stages 64 rows per lane in shared memory and flushes row by row (dc_nn.cu radius_fill_kernel); every flush
    boundary has to land in the right row.  Reference layout against cKDTree.query_ball_point (nearest_neighbors.py:50-73)."""
    from oracle import oracle
    scans, poses = _street_points(2)
    pts = np.concatenate([s['points'].astype(np.float64) @ T[:3, :3].T + T[:3, 3] for s, T in zip(scans, poses)]).astype(np.float32)
    rng = np.random.default_rng(0)
    pts = pts[rng.permutation(len(pts))[:60000]]
    p = torch.as_tensor(pts, device=dev)
    _, idx = dc.nearest_neighbors(p, p, r=0.6)
    _, ref = oracle.nearest_neighbors(torch.as_tensor(pts.astype(np.float64)), r=0.6)
    assert ref.shape[1] > 200                         # several flushes per slice
    assert torch.equal(idx.cpu(), ref)
    # cross query with a ragged last slice
    q = torch.as_tensor(pts[:1000 + 13] + np.float32(0.01), device=dev)
    _, idx = dc.nearest_neighbors(p, q, r=0.5)
    _, ref = oracle.nearest_neighbors(torch.as_tensor(pts.astype(np.float64)), torch.as_tensor((pts[:1013] + np.float32(0.01)).astype(np.float64)), r=0.5)
    assert torch.equal(idx.cpu(), ref)
