"""Pipeline glue of the hot path (preproc.py:25-243 of the reference, ball neighbourhoods only).

`global_cloud` returns a *lazy* GlobalCloud: it remembers the per-scan clouds, the model and the
poses instead of eagerly running model -> transform -> concatenate (preproc.py:80-119).  When such
a cloud reaches `min_eigval_loss` / `trace_loss` with a fixed neighbourhood graph attached by
`compute_neighborhood_features`, the whole chain runs in the fused sm_100a kernels (fused.py);
any attribute the reference would have computed (vps, dirs, depth, points, mean, cov, eigvals, ...)
is still available and is materialised on first access through the staged kernels.
Plane neighbourhoods (RANSAC planes, preproc.py:218-243) are out of scope.
"""
import numpy as np
import torch

from . import _lib as _L
from .config import NeighborhoodType
from .depth_cloud import DepthCloud
from .filters import (feature_mask, filter_depth, filter_eigenvalue_ratios, filter_eigenvalues, filter_shadow_points,
                      filter_valid_neighbors, within_bounds)
from .fused import StepState, model_kind_of, scan_table
from .graph import Graph, SortedMap, search
from .transform import xyz_axis_angle_to_matrix

__all__ = [
    'compute_neighborhood_features',
    'establish_neighborhoods',
    'filtered_cloud',
    'global_cloud',
    'global_cloud_mask', 'local_feature_clouds',
    'GlobalCloud',
    'local_feature_cloud',
    'Neighborhoods',
    'offset_cloud',
]

_FEATURE_FIELDS = ('points', 'mean', 'cov', 'eigvals', 'eigvecs', 'normals', 'inc_angles', 'trace')
_SOURCE_FIELDS = ('vps', 'dirs', 'depth', 'mask')


class Neighborhoods(object):
    """(neighbors, weights) pair as returned by establish_neighborhoods (preproc.py:185), backed by the
    sorted-space graph.  Unpacks like the reference's tuple; the padded tensors are built on demand."""

    def __init__(self, graph):
        self.graph = graph
        self._pair = None

    def _materialize(self):
        if self._pair is None:
            nb = self.graph.neighbors()
            nb._dc_graph = self.graph
            w = (nb >= 0).float()[..., None]
            w._dc_is_mask = True
            self._pair = (nb, w)
        return self._pair

    def __iter__(self):
        return iter(self._materialize())

    def __getitem__(self, i):
        return self._materialize()[i]

    def __len__(self):
        return 2


class GlobalCloud(DepthCloud):
    """Lazy global cloud = concatenation of model-corrected, pose-transformed scans."""

    def __init__(self, clouds, model, poses):
        # deliberately no DepthCloud.__init__: source and feature fields appear on first access
        self._scans = list(clouds)
        self._model = model
        self._poses = poses
        self._graph = None
        self._neighbors = None
        self._weights = None
        self._distances = None
        self._distances_stale = False
        self._features_pending = False
        self._scale = None
        self.neighbor_points = None
        self.dir_neighbors = self.dir_neighbor_weights = self.dir_distances = None
        self.loss = None

    def __getattr__(self, name):
        # only reached when `name` is not in __dict__
        if name in _SOURCE_FIELDS:
            self._materialize_sources()
            return self.__dict__[name]
        if name in _FEATURE_FIELDS:
            if '_scans' not in self.__dict__:
                raise AttributeError(name)
            self._materialize_features()
            return self.__dict__.get(name)
        raise AttributeError(name)

    def size(self):
        return sum(len(c) for c in self._scans)

    def device(self):
        return self._scans[0].depth.device

    def poses_tensor(self):
        p = self._poses
        if isinstance(p, (list, tuple)):
            p = torch.stack(list(p))
        return p

    def _materialize_sources(self):
        """Reference semantics, staged: model(cloud).transform(pose), concatenated (preproc.py:108-119)."""
        poses = self.poses_tensor()
        parts = []
        for i, cloud in enumerate(self._scans):
            if self._model is not None:
                cloud = self._model(cloud)
            parts.append(cloud.transform(poses[i]))
        dc = DepthCloud.concatenate(parts, dependent=True)
        for f in ('vps', 'dirs', 'depth', 'mask', 'normals'):
            if f not in self.__dict__:
                self.__dict__[f] = getattr(dc, f)

    def _materialize_features(self):
        for f in _FEATURE_FIELDS:
            self.__dict__.setdefault(f, None)
        if self._graph is None and self._neighbors is None:
            self.update_points()
            return
        self._features_pending = False
        DepthCloud.update_all(self, scale=self._scale, keep_neighbors=True)

    def copy(self):
        dc = GlobalCloud(self._scans, self._model, self._poses)
        dc.__dict__.update(self.__dict__)
        return dc

    # ---- fused path ---------------------------------------------------------------------------
    def fusable(self):
        if self._graph is None or self._scale is not None or self._kernel_weights() is not None:
            return False
        if not self._scans or not self._scans[0].depth.is_cuda:
            return False
        if self._graph.map.n != self.size() or not self._graph.self_query:
            return False
        try:
            model_kind_of(self._model)
        except NotImplementedError:
            return False
        # source fields overridden by the caller would not be seen by the packed records
        return not any(f in self.__dict__ for f in ('vps', 'dirs', 'depth'))

    def step_state(self):
        """The packed scan records are a snapshot of the scans' tensors: the cache entry is valid only for the very
        same tensor objects at the same version (in-place edits bump `_version`), and it keeps strong references to
        them so that neither `id()` nor device addresses can be recycled while it lives."""
        tensors = []
        for c in self._scans:
            tensors += [c.depth, c.dirs, c.vps, c.inc_angles, c.mask]
        key = tuple(None if t is None else (t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)
        cache = self._graph._step_cache
        st = cache.get(key)
        if st is None or len(st._keyed) != len(tensors) or any(a is not b for a, b in zip(st._keyed, tensors)):
            cache.clear()
            st = StepState(self._graph, self._scans)
            st._keyed = tensors
            cache[key] = st
        return st


def filtered_cloud(cloud, cfg):
    """Depth and voxel-grid filters of preproc.py:25-32."""
    from .filters_grid import filter_grid
    if ((cfg.min_depth is not None and cfg.min_depth > 0.0)
            or (cfg.max_depth is not None and cfg.max_depth < float('inf'))):
        cloud = filter_depth(cloud, min=cfg.min_depth, max=cfg.max_depth, log=cfg.log_filters)
    if cfg.grid_res > 0.0:
        rng = np.random.default_rng(cfg.random_seed)
        cloud = filter_grid(cloud, grid_res=cfg.grid_res, keep='random', log=cfg.log_filters, rng=rng)
    return cloud


def local_feature_cloud(cloud, cfg):
    """Per-scan features: neighbours + mean/cov/eig/normals/incidence angles + planarity mask (preproc.py:35-64)."""
    if isinstance(cloud, np.ndarray):
        if cloud.dtype.names:
            cloud = DepthCloud.from_structured_array(cloud, dtype=cfg.numpy_float_type(), device=cfg.device)
        else:
            cloud = DepthCloud.from_points(cloud, dtype=cfg.numpy_float_type(), device=cfg.device)
    assert isinstance(cloud, DepthCloud)
    if getattr(cfg, 'shadow_angle_bounds', None):
        cloud.update_dir_neighbors(angle=cfg.shadow_neighborhood_angle)
        cloud = filter_shadow_points(cloud, list(cfg.shadow_angle_bounds), log=cfg.log_filters)
    cloud.update_all(k=cfg.nn_k, r=cfg.nn_r)
    if cfg.eigenvalue_bounds or cfg.eigenvalue_ratio_bounds:
        # every bound of the configuration in one launch (the reference: one compare + one AND per bound, :53-62)
        cloud.mask = feature_mask(cloud, eigenvalue_bounds=cfg.eigenvalue_bounds,
                                  eigenvalue_ratio_bounds=cfg.eigenvalue_ratio_bounds, mask=cloud.mask)
        if cfg.log_filters:
            filter_eigenvalues(cloud, cfg.eigenvalue_bounds, only_mask=True, log=True)
            filter_eigenvalue_ratios(cloud, cfg.eigenvalue_ratio_bounds, only_mask=True, log=True)
    return cloud


LOCAL_FEATURES_CHUNK = 1 << 23


def local_feature_clouds(clouds, cfg):
    """[local_feature_cloud(c, cfg) for c in clouds] (the setup loop of train.py:97-104 / scripts/model_poses_learning:85-86)
    in a handful of launches instead of ~15 per scan: ONE stacked neighbour search (every scan in its own band of cell
    layers: dc_cell_keys_stacked), ONE pass over the neighbourhoods (dc_step_forward: mean, covariance, eigen in fp64),
    one kernel for normals / incidence angles (dc_local_features_finish) and one for the masks (dc_feature_mask).

    Sets points, mean, eigvals, normals, inc_angles and mask on every cloud (cov, eigvecs and the per-scan neighbour
    lists, which only the staged API reads, are not kept).  Same neighbour sets as the per-scan search; eigenvalues agree
    to rounding (the fused pass treats |lambda0| <= 1e-14 lambda2 as exactly 0, so masks can differ on rank-deficient
    neighbourhoods only).  Falls back to the per-scan loop for host clouds, a shadow filter, or kNN without a radius."""
    from . import _lib as L
    clouds = list(clouds)
    batchable = (len(clouds) > 1 and all(isinstance(c, DepthCloud) and c.depth.is_cuda for c in clouds) and cfg.nn_r
                 and not getattr(cfg, 'shadow_angle_bounds', None) and len({c.depth.dtype for c in clouds}) == 1
                 and cfg.nn_type == NeighborhoodType.ball)
    if not batchable:
        return [local_feature_cloud(c, cfg) for c in clouds]
    total = sum(len(c) for c in clouds)
    if total > LOCAL_FEATURES_CHUNK:
        # bound the temporaries (graph + fp64 records + stash of ALL scans at once): groups of up to ~8 M points
        out, group, count = [], [], 0
        for c in clouds:
            if group and count + len(c) > LOCAL_FEATURES_CHUNK:
                out += local_feature_clouds(group, cfg) if len(group) > 1 else [local_feature_cloud(group[0], cfg)]
                group, count = [], 0
            group.append(c)
            count += len(c)
        out += local_feature_clouds(group, cfg) if len(group) > 1 else [local_feature_cloud(group[0], cfg)]
        return out
    dev, dt = clouds[0].depth.device, clouds[0].depth.dtype
    st = L.stream()
    sizes = [len(c) for c in clouds]
    n = sum(sizes)
    pts = torch.cat([c.get_points().detach() for c in clouds]).contiguous()
    dirs = torch.cat([c.dirs.detach() for c in clouds]).contiguous()
    first_host = [0]
    for m in sizes:
        first_host.append(first_host[-1] + m)
    first = L.upload(first_host, torch.int64, dev)
    graph = search(pts, None, k=cfg.nn_k, r=cfg.nn_r, stack_first=first)
    smap = graph.map
    meta = torch.full((n,), 3, dtype=torch.int32, device=dev)
    stash = torch.empty((n, 8), dtype=torch.float64, device=dev)
    eig_sorted = torch.empty((n, 3), dtype=torch.float64, device=dev)
    L.call('dc_step_forward', L.ptr(smap.P), L.ptr(meta), n, L.ptr(graph.slice_ptr), L.ptr(graph.ell_idx), L.LOSS_MIN_EIGVAL,
           L.FLAG_RAW, None, L.ptr(stash), L.ptr(eig_sorted), None, None, 0, st)
    eigvals = torch.empty((n, 3), dtype=dt, device=dev)
    mean = torch.empty((n, 3), dtype=dt, device=dev)
    normals = torch.empty((n, 3), dtype=dt, device=dev)
    inc = torch.empty((n, 1), dtype=dt, device=dev)
    L.call('dc_local_features_finish', L.ptr(stash), L.ptr(eig_sorted), L.ptr(smap.order), L.ptr(dirs), L.dtype_code(dt), n, 0,
           L.ptr(eigvals), L.ptr(mean), L.ptr(normals), L.ptr(inc), st)
    mask = None
    if cfg.eigenvalue_bounds or cfg.eigenvalue_ratio_bounds:
        start = None
        if any(c.mask is not None for c in clouds):
            start = torch.cat([c.mask if c.mask is not None else torch.ones(m, dtype=torch.bool, device=dev) for c, m in zip(clouds, sizes)])
        holder = DepthCloud(dirs=dirs[:1], depth=inc[:1])          # (feature_mask only reads .eigvals of its argument)
        holder.eigvals = eigvals
        mask = feature_mask(holder, eigenvalue_bounds=cfg.eigenvalue_bounds, eigenvalue_ratio_bounds=cfg.eigenvalue_ratio_bounds, mask=start)
    for s, c in enumerate(clouds):
        a, b = first_host[s], first_host[s + 1]
        c.points = pts[a:b]
        c.mean, c.eigvals, c.normals, c.inc_angles = mean[a:b], eigvals[a:b], normals[a:b], inc[a:b]
        if mask is not None:
            c.mask = mask[a:b]
    return clouds


def offset_cloud(clouds, model):
    corrected = [model(c) if model is not None else c for c in clouds]
    return DepthCloud.concatenate(corrected, fields=DepthCloud.source_fields + ['eigvals'])


def global_cloud(clouds=None, model=None, poses=None, pose_corrections=None, dataset=None):
    """Global cloud with corrected depth (preproc.py:80-119) -- lazy, see GlobalCloud."""
    if dataset is not None:
        assert clouds is None
        assert poses is None
        clouds, poses = zip(*dataset)
        clouds = [DepthCloud.from_structured_array(c, dtype=np.float32, device='cuda') for c in clouds]
        poses = torch.as_tensor(np.array(poses), device='cuda')
    assert clouds is not None
    assert poses is not None
    if pose_corrections is not None:
        if isinstance(poses, (list, tuple)):
            poses = torch.stack(list(poses))
        if pose_corrections.shape[-1] == 6:
            pose_corrections = xyz_axis_angle_to_matrix(pose_corrections)
        poses = poses @ pose_corrections
    return GlobalCloud(clouds, model, poses)


def global_cloud_mask(cloud, mask, cfg):
    """Mask of points used by the loss (preproc.py:122-164)."""
    # neighbour count, eigenvalue and eigenvalue-ratio bounds: one launch (the reference: :130-142, a compare and an
    # AND per bound)
    mask = feature_mask(cloud, eigenvalue_bounds=cfg.eigenvalue_bounds, eigenvalue_ratio_bounds=cfg.eigenvalue_ratio_bounds,
                        min_valid_neighbors=cfg.min_valid_neighbors, mask=mask)
    if cfg.log_filters:
        if cfg.min_valid_neighbors:
            filter_valid_neighbors(cloud, min=cfg.min_valid_neighbors, only_mask=True, log=True)
        filter_eigenvalues(cloud, bounds=cfg.eigenvalue_bounds, only_mask=True, log=True)
        filter_eigenvalue_ratios(cloud, bounds=cfg.eigenvalue_ratio_bounds, only_mask=True, log=True)
    if cfg.dir_dispersion_bounds:
        mask = mask & within_bounds(cloud.dir_dispersion(), bounds=cfg.dir_dispersion_bounds)
    if cfg.vp_dispersion_bounds:
        mask = mask & within_bounds(cloud.vp_dispersion(), bounds=cfg.vp_dispersion_bounds)
    if cfg.vp_dispersion_to_depth2_bounds:
        mask = mask & within_bounds(cloud.vp_dispersion_to_depth2(), bounds=cfg.vp_dispersion_to_depth2_bounds)
    return mask


def establish_neighborhoods(dataset=None, clouds=None, poses=None, cloud=None, cfg=None):
    """Neighbourhood graph of the initial global cloud (preproc.py:168-191): kernel 1 on the GPU.
    Returns a Neighborhoods pair (unpacks to (neighbors int64 [N,K], weights float32 [N,K,1]))."""
    if cloud is None:
        cloud = global_cloud(clouds=clouds, poses=poses, dataset=dataset)
    assert cloud is not None
    if cfg.nn_type != NeighborhoodType.ball:
        raise NotImplementedError('plane neighbourhoods (RANSAC) are out of scope of the B200 hot path')
    pts = _initial_map_points(cloud)
    graph = search(pts, None, k=cfg.nn_k, r=cfg.nn_r)
    return Neighborhoods(graph)


def _initial_map_points(cloud):
    """Points of the (uncorrected) global cloud the graph is built on.  For a lazy cloud on the GPU every scan is
    moved to the map frame by one dc_world_points launch (fp64 from the stored records); any other cloud goes
    through the staged to_points()."""
    scans = getattr(cloud, '_scans', None)
    if not (isinstance(cloud, GlobalCloud) and scans and cloud._model is None and scans[0].depth.is_cuda
            and not any(f in cloud.__dict__ for f in ('vps', 'dirs', 'depth'))):
        return cloud.to_points().detach()
    dev = scans[0].depth.device
    dt = scans[0].depth.dtype
    poses = cloud.poses_tensor().detach().to(device=dev, dtype=torch.float64).reshape(len(scans), 16).contiguous()
    n = cloud.size()
    out = torch.empty((n, 3), dtype=torch.float64, device=dev)
    tbl, first, _keep = scan_table(scans, dt)
    _L.call('dc_world_points_batched', _L.ptr(tbl), _L.ptr(first), len(scans), n, _L.dtype_code(dt), _L.ptr(poses),
            _L.ptr(out), _L.stream())
    return out


def compute_neighborhood_features(dataset=None, clouds=None, poses=None, model=None, pose_corrections=None, cloud=None,
                                  neighborhoods=None, cfg=None):
    """Attach a fixed graph to the (lazy) global cloud (preproc.py:195-217, ball branch).

    The reference runs update_all(keep_neighbors=True) here; we defer it: the losses run the fused
    kernels, and reading any feature attribute computes them through the staged kernels."""
    if cfg.nn_type != NeighborhoodType.ball:
        raise NotImplementedError('plane neighbourhoods (RANSAC) are out of scope of the B200 hot path')
    if neighborhoods is None:
        neighborhoods = establish_neighborhoods(dataset=dataset, cloud=cloud, cfg=cfg)
    if cloud is None:
        cloud = global_cloud(clouds=clouds, model=model, poses=poses, pose_corrections=pose_corrections, dataset=dataset)
    assert neighborhoods is not None
    if isinstance(neighborhoods, Neighborhoods):
        graph = neighborhoods.graph
        weights = None
    else:
        nb, weights = neighborhoods
        graph = getattr(nb, '_dc_graph', None)
        if graph is None:
            # a graph from elsewhere (e.g. the reference's cKDTree): import it once and remember it on the tensor
            ref_pts = cloud.copy().to_points().detach()   # on a copy: keeps the lazy cloud unmaterialised
            cell = getattr(cfg, 'nn_r', None) or 0.5
            graph = Graph.from_padded(SortedMap(ref_pts, cell), nb.to(ref_pts.device))
            nb._dc_graph = graph
        if weights is not None and getattr(weights, '_dc_is_mask', False):
            weights = None
    if isinstance(cloud, GlobalCloud):
        for f in _FEATURE_FIELDS:
            cloud.__dict__.pop(f, None)
        cloud._graph = graph
        cloud._neighbors = None
        cloud._weights = weights
        cloud._scale = cfg.nn_scale
        cloud._features_pending = True
        return cloud
    cloud._graph = graph
    cloud._neighbors = None
    cloud._weights = weights
    cloud.update_all(scale=cfg.nn_scale, keep_neighbors=True)
    return cloud
