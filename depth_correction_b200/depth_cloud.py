"""Drop-in for depth_correction.depth_cloud.DepthCloud (depth_cloud.py:18-741), B200-native.

Same field bag and method names; the heavy methods call the sm_100a kernels of libdcb200.so:

    update_neighbors           -> kernel 1 (grid search)                  depth_cloud.py:210-215
    update_mean / update_cov   -> dc_features (+ backward)                depth_cloud.py:291-295,366-369
    update_eig                 -> dc_eigh3 (+ backward), on the GPU       depth_cloud.py:376-399
    update_normals / update_incidence_angles -> dc_normals_angles         depth_cloud.py:401-424

Like the reference, updates rebind attributes to fresh tensors (no in-place writes into
caller-visible tensors), `copy()` is shallow and slicing keeps only `sliced_fields`.
`neighbors`, `weights` and `distances` are materialised lazily from the internal sorted-space
graph, so a training loop that never reads them never pays for the padded int64 matrix.
ROS / open3d / matplotlib helpers of the reference (visualisation, meshes, messages) are out of scope.
"""
import numpy as np
from numpy.lib.recfunctions import structured_to_unstructured
import torch

from .nearest_neighbors import ball_angle_to_distance, nearest_neighbors
from .graph import search
from . import ops
from . import _lib as L
from .utils import covs, trace

__all__ = ['DepthCloud']


class DepthCloud(object):
    """Point cloud constructed from viewpoints, directions, and depths."""
    source_fields = ['vps', 'dirs', 'depth']
    sliced_fields = (source_fields
                     + ['points', 'mean', 'cov', 'eigvals', 'eigvecs',
                        'normals', 'inc_angles', 'trace',
                        'loss', 'mask'])
    not_sliced_fields = ['neighbors', 'weights', 'distances', 'neighbor_points',
                         'dir_neighbors', 'dir_neighbor_weights', 'dir_distances']
    all_fields = sliced_fields + not_sliced_fields

    def __init__(self, vps=None, dirs=None, depth=None,
                 points=None, mean=None, cov=None, eigvals=None, eigvecs=None,
                 normals=None, inc_angles=None, trace=None,
                 loss=None, mask=None,
                 neighbors=None, distances=None, neighbor_points=None, weights=None,
                 dir_neighbors=None, dir_neighbor_weights=None, dir_distances=None):
        if vps is None:
            vps = torch.zeros((1, 3))
        assert isinstance(vps, torch.Tensor)
        assert vps.shape[-1] == 3
        assert isinstance(dirs, torch.Tensor)
        assert dirs.shape[-1] == 3
        assert dirs.shape == vps.shape or vps.shape == (1, 3)
        assert isinstance(depth, torch.Tensor)
        assert depth.shape[-1] == 1
        assert depth.shape[:-1] == dirs.shape[:-1]

        self.vps = vps
        self.dirs = dirs
        self.depth = depth
        self.dir_neighbors = dir_neighbors
        self.dir_neighbor_weights = dir_neighbor_weights
        self.dir_distances = dir_distances
        self.points = points
        self._graph = None            # internal sorted-space graph (depth_correction_b200.graph.Graph)
        self._neighbors = None
        self._weights = None
        self._distances = None
        self._distances_stale = False
        self.neighbor_points = neighbor_points
        self.mean = mean
        self.cov = cov
        self.eigvals = eigvals
        self.eigvecs = eigvecs
        self.normals = normals
        self.inc_angles = inc_angles
        self.trace = trace
        self.loss = loss
        self.mask = mask
        if neighbors is not None:
            self.neighbors = neighbors
        if weights is not None:
            self.weights = weights
        if distances is not None:
            self.distances = distances

    # ---- lazily materialised graph views ------------------------------------------------------
    @property
    def neighbors(self):
        if self._neighbors is None and self._graph is not None:
            self._neighbors = self._graph.neighbors()
            self._neighbors._dc_graph = self._graph
        return self._neighbors

    @neighbors.setter
    def neighbors(self, value):
        self._neighbors = value
        self._graph = getattr(value, '_dc_graph', None) if value is not None else None
        self._weights = None
        self._distances = None

    @property
    def weights(self):
        if self._weights is None and self.neighbors is not None:
            self._weights = self.valid_neighbor_mask().float()[..., None]
            self._weights._dc_is_mask = True
        return self._weights

    @weights.setter
    def weights(self, value):
        self._weights = value

    @property
    def distances(self):
        if self._distances_stale:
            x = self.get_points()
            self._distances = torch.linalg.norm(x.unsqueeze(dim=1) - x[self.neighbors], dim=-1)
            self._distances_stale = False
        elif self._distances is None and self._graph is not None and self._graph.mode == 'knn':
            self._distances = self._graph.distances()
        return self._distances

    @distances.setter
    def distances(self, value):
        self._distances = value
        self._distances_stale = False

    def _kernel_weights(self):
        """None when the weights are just the valid-neighbour mask (the kernels derive it from the indices)."""
        w = self._weights
        if w is None or getattr(w, '_dc_is_mask', False):
            return None
        return w

    # ---- container ----------------------------------------------------------------------------
    def copy(self):
        """Create shallow copy of the cloud."""
        dc = DepthCloud(self.vps, self.dirs, self.depth)
        for f in DepthCloud.sliced_fields + ['neighbor_points', 'dir_neighbors', 'dir_neighbor_weights', 'dir_distances']:
            setattr(dc, f, getattr(self, f))
        dc._graph, dc._neighbors, dc._weights = self._graph, self._neighbors, self._weights
        dc._distances, dc._distances_stale = self._distances, self._distances_stale
        return dc

    def clone(self):
        """Create deep copy of the cloud (gradients are still propagated if detach is not called)."""
        kwargs = {}
        for f in DepthCloud.all_fields:
            x = getattr(self, f)
            if x is not None:
                kwargs[f] = x.clone()
        dc = DepthCloud(**kwargs)
        dc._graph = self._graph
        return dc

    def size(self):
        return self.dirs.shape[0]

    def __len__(self):
        return self.size()

    def to_points(self):
        return self.vps + self.depth * self.dirs

    def update_points(self):
        self.points = self.to_points()
        self.neighbor_points = None

    def get_points(self):
        if self.points is None:
            self.update_points()
        return self.points

    def transform(self, T):
        assert isinstance(T, torch.Tensor)
        assert T.shape == (4, 4)
        T = T.to(dtype=self.vps.dtype, device=self.vps.device)
        R = T[:3, :3]
        t = T[:3, 3:]
        vps = torch.matmul(self.vps, R.t()) + t.t()
        dirs = torch.matmul(self.dirs, R.t())
        kwargs = {'mask': self.mask}
        if self.normals is not None:
            kwargs['normals'] = torch.matmul(self.normals, R.t())
        return DepthCloud(vps, dirs, self.depth, **kwargs)

    def __getitem__(self, item):
        kwargs = {}
        if isinstance(item, list) and len(item) > 0 and isinstance(item[0], str):
            for f in item:
                kwargs[f] = getattr(self, f)
        else:
            for f in DepthCloud.sliced_fields:
                x = getattr(self, f)
                if x is not None:
                    if f == 'vps' and x.shape[0] == 1 and self.dirs.shape[0] != 1:
                        kwargs[f] = x
                    else:
                        kwargs[f] = x[item]
        return DepthCloud(**kwargs)

    def __add__(self, other):
        return DepthCloud.concatenate([self, other], dependent=True)

    def collect_neighbors(self, item):
        assert self.neighbors is not None
        idx = self.neighbors[item].unique()
        return idx[idx >= 0]

    def filter_with_neighbors(self, item):
        return self[self.collect_neighbors(item)]

    # ---- neighbourhoods -----------------------------------------------------------------------
    def update_distances(self):
        assert self.neighbors is not None or self._graph is not None
        self._distances_stale = True     # evaluated on first read (the losses never read it)

    def valid_neighbor_mask(self):
        assert self.neighbors is not None
        return self.neighbors >= 0

    def num_valid_neighbors(self):
        """neighbors >= 0 summed over rows, without materialising the padded matrix."""
        if self._graph is not None:
            return self._graph.valid_counts()
        return self.valid_neighbor_mask().sum(dim=-1)

    def update_neighbors(self, k=None, r=None):
        assert self.points is not None
        assert k or r
        pts = self.get_points()
        if not pts.is_cuda:
            raise RuntimeError('DepthCloud.update_neighbors needs a CUDA cloud; there is no CPU fallback')
        self._graph = search(pts, None, k=k, r=r)
        self._neighbors = None
        self._weights = None
        self._distances = None
        self._distances_stale = False
        self.neighbor_points = None

    def update_dir_neighbors(self, k=None, r=None, angle=None):
        assert self.dirs is not None
        if angle is not None:
            assert r is None
            r = ball_angle_to_distance(torch.as_tensor(angle)).item()
        self.dir_distances, self.dir_neighbors = nearest_neighbors(self.dirs, self.dirs, k=k, r=r)
        self.dir_neighbor_weights = (self.dir_neighbors >= 0).float()

    def compute_neighbor_points(self):
        return self.get_points()[self.neighbors]

    def update_neighbor_points(self):
        self.neighbor_points = self.compute_neighbor_points()

    def get_neighbor_points(self):
        if self.neighbor_points is None:
            self.update_neighbor_points()
        return self.neighbor_points

    # ---- features -----------------------------------------------------------------------------
    def update_mean(self, invalid=0.0):
        self.mean = ops.neighborhood_mean_cov(self.get_points(), self.neighbors, self._kernel_weights(), mean=True, cov=False)

    def update_weights(self, scale=None):
        assert self.mean is not None
        weights = self.valid_neighbor_mask().float()[..., None]
        if scale is not None:
            dist = (self.get_points() - self.mean).norm(dim=1, keepdim=True)
            weights = weights * torch.exp(-(dist / scale) ** 2).unsqueeze(-1).to(weights.dtype)
        else:
            weights._dc_is_mask = True
        self._weights = weights

    def update_cov(self, correction=1, invalid=0.0):
        self.cov = ops.neighborhood_mean_cov(self.get_points(), self.neighbors, self._kernel_weights(), mean=False, cov=True)

    def compute_eig(self):
        assert self.cov is not None
        return ops.eigh3(self.cov)

    def update_eig(self):
        self.eigvals, self.eigvecs = self.compute_eig()

    def orient_normals(self):
        assert isinstance(self.dirs, torch.Tensor)
        assert isinstance(self.normals, torch.Tensor)
        cos = (self.dirs * self.normals).sum(dim=-1)
        self.normals = - torch.sign(cos)[..., None] * self.normals

    def update_normals(self):
        assert self.eigvecs is not None
        if self.eigvecs.requires_grad or self.dirs.requires_grad or not self.eigvecs.is_cuda:
            self.normals = self.eigvecs[..., 0]
            self.orient_normals()
        else:
            self.normals, _ = ops.normals_and_angles(self.dirs, self.eigvecs, want_angles=False)

    def update_incidence_angles(self, use_normal_sign=False):
        assert self.dirs is not None
        assert self.normals is not None
        if use_normal_sign:
            inc_angles = torch.arccos(-(self.dirs * self.normals).sum(dim=-1)).unsqueeze(-1)
        else:
            inc_angles = torch.arccos((self.dirs * self.normals).sum(dim=-1).abs()).unsqueeze(-1)
        self.inc_angles = inc_angles

    def update_features(self, scale=None):
        pts = self.get_points()
        fused_ok = (scale is None and self._kernel_weights() is None and pts.is_cuda and not pts.requires_grad)
        if fused_ok:
            # one gather for mean and covariance, eigen-solve and normals / angles without autograd bookkeeping
            self.mean, self.cov = ops.neighborhood_mean_cov(pts, self.neighbors, None, mean=True, cov=True)
            w = self.weights     # keep the reference's side effect: weights = valid mask, float32 [N,K,1]
            self.update_eig()
            self.normals, self.inc_angles = ops.normals_and_angles(self.dirs, self.eigvecs)
            return
        self.update_mean()
        self.update_weights(scale=scale)
        self.update_cov()
        self.update_eig()
        self.update_normals()
        self.update_incidence_angles()

    def update_all(self, k=None, r=None, scale=None, keep_neighbors=False):
        self.update_points()
        if keep_neighbors:
            self.update_distances()
        else:
            self.update_neighbors(k=k, r=r)
        self.update_features(scale=scale)

    # ---- neighbourhood statistics used by global_cloud_mask (depth_cloud.py:314-354) ------------
    def vp_dispersion(self):
        assert self.vps is not None and self.neighbors is not None
        vps = self.vps.expand(self.size(), 3).contiguous()
        return trace(ops.neighborhood_mean_cov(vps, self.neighbors, self._kernel_weights(), mean=False, cov=True))

    def dir_dispersion(self):
        assert self.dirs is not None and self.neighbors is not None
        return trace(ops.neighborhood_mean_cov(self.dirs.contiguous(), self.neighbors, self._kernel_weights(), mean=False, cov=True))

    def _neighbor_stats(self, want_depth, want_vp):
        """One pass over the neighbour lists (dc_neighbor_stats) instead of [N,K(,3)] temporaries."""
        n = self.size()
        depth = self.depth.detach().reshape(-1).contiguous()
        if not depth.is_cuda:
            raise RuntimeError('neighbourhood statistics need a CUDA cloud; there is no CPU fallback')
        nb = self.neighbors.contiguous()
        w = self._kernel_weights()
        w = None if w is None else w.detach().reshape(nb.shape).to(torch.float32).contiguous()
        vps = self.vps.detach().to(depth.dtype).expand(n, 3).contiguous() if want_vp else None
        md = torch.empty(n, dtype=depth.dtype, device=depth.device) if want_depth else None
        mv = torch.empty(n, dtype=depth.dtype, device=depth.device) if want_vp else None
        L.call('dc_neighbor_stats', L.ptr(depth), L.ptr(vps), L.dtype_code(depth.dtype), L.ptr(nb), L.ptr(w), n, nb.shape[1],
               L.ptr(md), L.ptr(mv), L.stream())
        return md, mv

    def mean_depth(self):
        assert self.neighbors is not None
        return self._neighbor_stats(True, False)[0]

    def mean_vp_dist(self):
        assert self.vps is not None and self.neighbors is not None
        return self._neighbor_stats(False, True)[1]

    def vp_dispersion_to_depth2(self):
        return self.vp_dispersion() / self.mean_depth() ** 2

    def vp_dist_to_depth(self, mode='mean'):
        return self.mean_vp_dist() / self.mean_depth()

    # ---- I/O ----------------------------------------------------------------------------------
    def to_structured_array(self, colors=None):
        """depth_cloud.py:508-533: x, y, z, vp_*[, normal_*, inc_angle, loss, mask, r, g, b] (float32; mask uint8).
        Built column by column (numpy.lib.recfunctions.merge_arrays, which the reference uses, cannot merge an
        unsigned field in numpy >= 2: its default fill value -1 overflows uint8)."""
        cols = []

        def add(x, names, dtype=np.float32):
            a = np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x, dtype=dtype).reshape(self.size(), -1)
            assert a.shape[1] == len(names)
            cols.extend((name, dtype, a[:, i]) for i, name in enumerate(names))

        add(self.get_points(), ['x', 'y', 'z'])
        add(self.vps.expand(self.size(), 3), ['vp_%s' % f for f in 'xyz'])
        if self.normals is not None:
            add(self.normals, ['normal_%s' % f for f in 'xyz'])
        if self.inc_angles is not None:
            add(self.inc_angles, ['inc_angle'])
        if self.loss is not None:
            add(self.loss, ['loss'])
        if self.mask is not None:
            add(self.mask, ['mask'], np.uint8)
        if colors is not None:
            add(colors, ['r', 'g', 'b'])
        out = np.empty(self.size(), dtype=[(name, dtype) for name, dtype, _ in cols])
        for name, _, values in cols:
            out[name] = values
        return out

    @staticmethod
    def concatenate(clouds, fields=None, dependent=False):
        if not fields:
            fields = DepthCloud.all_fields if dependent else DepthCloud.source_fields
        else:
            assert not dependent
        kwargs = {}
        for f in fields:
            xs = [getattr(dc, f) for dc in clouds]
            valid = [x is not None for x in xs]
            if all(valid):
                if f in ('dir_neighbors', 'neighbors'):
                    # shift indices by the number of points in preceding clouds (out of place; the
                    # reference shifts in place, depth_cloud.py:555-559, and corrupts -1 padding)
                    sizes = [len(cloud) for cloud in clouds]
                    shift = [0] + list(np.cumsum(sizes[:-1]))
                    xs = [torch.where(x >= 0, x + int(s), x) for x, s in zip(xs, shift)]
                    width = max(x.shape[1] for x in xs)
                    xs = [torch.nn.functional.pad(x, (0, width - x.shape[1]), value=-1) for x in xs]
                elif f == 'vps':
                    xs = [x.expand(len(dc), 3) for x, dc in zip(xs, clouds)]
                elif f in ('weights', 'distances', 'dir_neighbor_weights', 'dir_distances', 'neighbor_points'):
                    widths = set(x.shape[1] for x in xs)
                    if len(widths) > 1:
                        continue
                kwargs[f] = torch.cat(xs)
            elif any(valid):
                print('Field %s not available for %i of %i clouds.' % (f, sum(valid), len(clouds)))
        return DepthCloud(**kwargs)

    @staticmethod
    def from_structured_array(arr, dtype=None, device=None):
        """Create depth cloud from a structured array with x, y, z[, vp_*, normal_*] fields."""
        assert isinstance(arr, np.ndarray)
        pts = structured_to_unstructured(arr[['x', 'y', 'z']], dtype=dtype)
        vps = normals = None
        if 'vp_x' in arr.dtype.names:
            vps = structured_to_unstructured(arr[['vp_x', 'vp_y', 'vp_z']], dtype=dtype)
        if 'normal_x' in arr.dtype.names:
            normals = structured_to_unstructured(arr[['normal_x', 'normal_y', 'normal_z']], dtype=dtype)
        return DepthCloud.from_points(pts, vps=vps, normals=normals, device=device)

    @staticmethod
    def from_points(pts, vps=None, normals=None, dtype=None, device=None):
        """Create depth cloud from points and viewpoints (depth_cloud.py:592-638)."""
        try:
            if pts.dtype.names:
                return DepthCloud.from_structured_array(pts)
        except AttributeError:
            pass
        if isinstance(dtype, type) and issubclass(dtype, np.generic):
            dtype = getattr(torch, np.dtype(dtype).name)
        pts = torch.as_tensor(pts, dtype=dtype, device=device)
        assert isinstance(pts, torch.Tensor)
        if vps is not None:
            vps = torch.as_tensor(vps, dtype=dtype, device=device)
            assert vps.shape == pts.shape
        if pts.is_cuda and pts.dim() == 2 and pts.shape[1] == 3 and pts.dtype in (torch.float32, torch.float64) \
                and not pts.requires_grad and (vps is None or not vps.requires_grad):
            # one kernel, no host synchronisation (dc_from_points)
            n = pts.shape[0]
            src = pts.contiguous()
            dirs = torch.empty_like(src)
            depth = torch.empty((n, 1), dtype=src.dtype, device=src.device)
            vps_out = torch.empty_like(src)
            L.call('dc_from_points', L.ptr(src), L.ptr(vps.contiguous()) if vps is not None else None,
                   L.dtype_code(src.dtype), n, L.ptr(dirs), L.ptr(depth), L.ptr(vps_out), L.stream())
            kwargs = {'vps': vps_out, 'depth': depth, 'dirs': dirs}
            if normals is not None:
                kwargs['normals'] = torch.as_tensor(normals, dtype=dtype, device=device)
            return DepthCloud(**kwargs)
        # host-side container construction (CPU tensors, or points that carry autograd history)
        if vps is None:
            vps = torch.zeros_like(pts)
        dirs = pts - vps
        depth = dirs.norm(dim=-1, keepdim=True)
        valid = depth[:, 0] > 0.0
        dirs[valid] = dirs[valid] / depth[valid]
        kwargs = {'vps': vps, 'depth': depth, 'dirs': dirs}
        if normals is not None:
            kwargs['normals'] = torch.as_tensor(normals, dtype=dtype, device=device)
        depth_cloud = DepthCloud(**kwargs)
        if device:
            depth_cloud = depth_cloud.to(device=device)
        return depth_cloud

    def to(self, device=None, dtype=None, float_type=None, int_type=None):
        kwargs = {}
        for f in DepthCloud.all_fields:
            x = getattr(self, '_' + f if f in ('neighbors', 'weights', 'distances') else f)
            if x is not None:
                if (float_type and x.dtype.is_floating_point) \
                        or (int_type and not x.dtype.is_floating_point) \
                        or (dtype and dtype.is_floating_point == x.dtype.is_floating_point):
                    x_type = dtype or float_type or int_type
                else:
                    x_type = None
                kwargs[f] = x.to(device=device, dtype=x_type)
        dc = DepthCloud(**kwargs)
        same_device = device is None or torch.device(device) == self.dirs.device
        if same_device and dc._graph is None:
            dc._graph = self._graph
        return dc

    def cpu(self):
        return self.to(torch.device('cpu'))

    def gpu(self):
        return self.to(torch.device('cuda:0'))

    def device(self):
        return self.depth.device

    def type(self, dtype=None):
        if dtype is None:
            assert self.vps.dtype == self.dirs.dtype == self.depth.dtype
            return self.vps.dtype
        for f in DepthCloud.all_fields:
            x = getattr(self, '_' + f if f in ('neighbors', 'weights', 'distances') else f)
            if x is not None and dtype.is_floating_point == x.dtype.is_floating_point:
                setattr(self, f, x.type(dtype))
        return self

    def float(self):
        return self.type(torch.float32)

    def double(self):
        return self.type(torch.float64)

    def detach(self):
        for f in DepthCloud.all_fields:
            x = getattr(self, '_' + f if f in ('neighbors', 'weights', 'distances') else f)
            if x is not None:
                g = getattr(x, '_dc_graph', None)
                x = x.detach()
                if g is not None:
                    x._dc_graph = g
                setattr(self, f, x)
        return self
