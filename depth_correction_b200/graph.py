"""Cell-sorted map and neighbourhood graph (host-side orchestration of kernel 1).

`SortedMap` = the point set sorted by uniform-grid cell (the replacement of the cKDTree index built
at nearest_neighbors.py:46); `Graph` = a neighbourhood graph in sorted space, sliced-ELL int32
(see include/dc_b200.h), convertible to/from the reference layout (int64 [N,K], -1 padded,
nearest_neighbors.py:69-78).  All heavy work is done by libdcb200.so; torch is used for device
memory and tiny glue (permutation inverse, scalar read-backs at setup time).
"""
import ctypes
import math
import os

import torch

from . import _lib as L

__all__ = ['SortedMap', 'Graph', 'search']

KNN_OCC_DEFAULT = '0.3'                # mean points per occupied cell / k the kNN cell size aims at
KNN_SAMPLE = 8192                     # queries searched to model the cost of the kNN kernel against the cell size
KNN_ROW_COST = 3.0                    # overhead of one row of cells, in candidates (fitted: tools/knn_cell_sweep.py)
KNN_TABLE_COST = 0.55                 # one entry of the dense cell table, in candidates of one query (0.83 ms per 2^30 cells)
KNN_MODEL_MIN_POINTS = 1 << 17        # smaller maps are launch bound: the occupancy estimate is good enough
KNN_PAD = 3                           # readable records dc_knn_recorded expects behind the n records of the map
DENSE_TABLE_MAX_CELLS = 1 << 30      # 4 GB of int32 cell starts at most (a 100 M point, 760 m corridor needs 3e8 cells)


def _as_points(x):
    assert isinstance(x, torch.Tensor)
    x = x.detach().reshape(-1, x.shape[-1])
    assert x.shape[-1] == 3, 'the B200 neighbour search is specialised for 3-D points'
    return x.contiguous()


class SortedMap(object):
    """Points sorted by grid cell: P (fp64 32-byte records), keys, order / inv_order, optional dense cell table."""

    @staticmethod
    def bounds_of(points, also_cover=None):
        """(lo[3], hi[3]) of the finite point set (one reduction kernel + one small read-back)."""
        dev = points.device
        st = L.stream()
        out = torch.empty((2, 6), dtype=torch.float64, device=dev)
        bad = torch.zeros(2, dtype=torch.int32, device=dev)
        L.call('dc_bounds', L.ptr(points), L.dtype_code(points.dtype), points.shape[0], L.ptr(out[0]), L.ptr(bad[0:1]), st)
        have_q = also_cover is not None and also_cover.shape[0] > 0
        if have_q:
            q = _as_points(also_cover)
            L.call('dc_bounds', L.ptr(q), L.dtype_code(q.dtype), q.shape[0], L.ptr(out[1]), L.ptr(bad[1:2]), st)
        b = out.cpu()
        if int(bad.cpu()[0]) > 0:
            raise ValueError('points must be finite (cKDTree raises as well)')
        lo, hi = b[0, :3], b[0, 3:]
        if have_q:
            lo, hi = torch.minimum(lo, b[1, :3]), torch.maximum(hi, b[1, 3:])
        if points.shape[0] == 0:
            lo, hi = torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64)
        return lo.tolist(), hi.tolist()

    @staticmethod
    def make_spec(lo, hi, cell):
        spec = L.GridSpec()
        spec.cell = float(cell)
        ext = []
        for a in range(3):
            spec.origin[a] = lo[a] - 1e-3 * cell
            spec.dims[a] = int(math.floor((hi[a] - spec.origin[a]) / cell)) + 1
            ext.append(hi[a] - lo[a])
        # Fastest key digit = LONGEST extent: surfaces of a mapped corridor / street run along the trajectory, so a
        # row of cells along that axis is one long contiguous run of the sorted map and 32 consecutive queries
        # share (almost) the same block of candidate cells (what the segment kNN kernel stages once per warp).
        # DC_AXIS_ORDER=short restores shortest-first.
        if os.environ.get('DC_AXIS_ORDER', 'long')[0] == 's':
            axes = sorted(range(3), key=lambda a: (ext[a], a))
        else:
            axes = sorted(range(3), key=lambda a: (-ext[a], a))
        for i, a in enumerate(axes):
            spec.axis[i] = a
        # points of one cell in Morton order of their 4x4x4 sub-cell (6 extra key bits): consecutive queries of a warp
        # are neighbours in space.  DC_SUB_ORDER=0 keeps the order of the caller inside a cell.
        spec.sub_bits = 6 if os.environ.get('DC_SUB_ORDER', '1') == '1' else 0
        n_cells = int(spec.dims[0]) * int(spec.dims[1]) * int(spec.dims[2])
        if n_cells >= (1 << 56):
            raise OverflowError('search grid has too many cells; increase the cell size')
        return spec, axes, n_cells

    @staticmethod
    def occupancy_of(points, lo, hi, cell):
        """Mean number of points per occupied cell for a candidate cell size (keys + key-only sort)."""
        n = points.shape[0]
        if n == 0:
            return 0.0
        spec, _, n_cells = SortedMap.make_spec(lo, hi, cell)
        dev = points.device
        st = L.stream()
        keys = L.scratch('keys', n, torch.int64, dev)
        ids = L.scratch('ids', n, torch.int32, dev)
        skeys = L.scratch('skeys', n, torch.int64, dev)
        L.call('dc_cell_keys', L.ptr(points), L.dtype_code(points.dtype), n, ctypes.byref(spec), L.ptr(keys), L.ptr(ids), st)
        sb = int(spec.sub_bits)
        L.call_with_temp('dc_sort_keys', dev, L.ptr(keys), L.ptr(skeys), n, sb, sb + max(1, int(n_cells - 1).bit_length()), after=(st,))
        n_occ = int(((skeys[1:] >> sb) != (skeys[:-1] >> sb)).sum().item()) + 1
        return n / n_occ

    def __init__(self, points, cell, also_cover=None, bounds=None, stack=None):
        """stack = (first int64 [S+1] on the device, S, guard): S clouds stored one after the other in `points`, searched
        at once but never paired with each other (dc_cell_keys_stacked)."""
        points = _as_points(points)
        dev = points.device
        self.device = dev
        self.n = points.shape[0]
        self.dtype = points.dtype
        code = L.dtype_code(points.dtype)
        n = self.n
        st = L.stream()
        lo, hi = bounds if bounds is not None else SortedMap.bounds_of(points, also_cover)
        self.cell = float(cell)
        self.spec, self.axes, self.n_cells = SortedMap.make_spec(lo, hi, self.cell)
        self.stack = None
        if stack is not None:
            first, n_clouds, guard = stack
            slow = self.axes[2]
            band = int(self.spec.dims[slow])
            period = band + int(guard)
            self.spec.dims[slow] = n_clouds * period
            self.n_cells = int(self.spec.dims[0]) * int(self.spec.dims[1]) * int(self.spec.dims[2])
            if self.n_cells >= (1 << 56):
                raise OverflowError('stacked search grid has too many cells; increase the cell size')
            self.stack = (first, int(n_clouds), period, int(guard))
        self.key_bits = max(1, int(self.n_cells - 1).bit_length()) + int(self.spec.sub_bits)

        keys = L.scratch('keys', n, torch.int64, dev)     # uint64 bit patterns (< 2^62)
        ids = L.scratch('ids', n, torch.int32, dev)
        self.keys = torch.empty(n, dtype=torch.int64, device=dev)
        self.order = torch.empty(n, dtype=torch.int32, device=dev)
        # KNN_PAD records behind the map: the first pass of dc_knn_recorded reads four consecutive records per step without
        # clamping the last ones to the end of the row (include/dc_b200.h); their content is never used
        buf = torch.empty((n + KNN_PAD, 4), dtype=torch.float64, device=dev)
        buf[n:].zero_()
        self.P = buf[:n]
        if n > 0:
            if self.stack is None:
                L.call('dc_cell_keys', L.ptr(points), code, n, ctypes.byref(self.spec), L.ptr(keys), L.ptr(ids), st)
            else:
                first, n_clouds, period, guard = self.stack
                L.call('dc_cell_keys_stacked', L.ptr(points), code, n, ctypes.byref(self.spec), L.ptr(first), n_clouds, period,
                       guard, L.ptr(keys), L.ptr(ids), st)
            L.call_with_temp('dc_sort_pairs', dev, L.ptr(keys), L.ptr(self.keys), L.ptr(ids), L.ptr(self.order), n,
                             self.key_bits, after=(st,))
        self.inv_order = torch.empty(n, dtype=torch.int32, device=dev)
        if n > 0:
            L.call('dc_gather_points', L.ptr(points), code, L.ptr(self.order), n, L.ptr(self.P), L.ptr(self.inv_order), st)
        self.cell_start = None
        if 0 < self.n_cells <= DENSE_TABLE_MAX_CELLS and n > 0:
            self.cell_start = torch.empty(self.n_cells + 1, dtype=torch.int32, device=dev)
            L.call('dc_cell_table', L.ptr(self.keys), n, self.n_cells, int(self.spec.sub_bits), L.ptr(self.cell_start), st)

    def sort_queries(self, query):
        """Sort a query set by the same grid -> (Q records, qkeys, q_order)."""
        query = _as_points(query)
        nq = query.shape[0]
        dev = self.device
        st = L.stream()
        code = L.dtype_code(query.dtype)
        keys = torch.empty(nq, dtype=torch.int64, device=dev)
        ids = torch.empty(nq, dtype=torch.int32, device=dev)
        qkeys = torch.empty(nq, dtype=torch.int64, device=dev)
        qorder = torch.empty(nq, dtype=torch.int32, device=dev)
        Q = torch.empty((nq, 4), dtype=torch.float64, device=dev)
        if nq > 0:
            L.call('dc_cell_keys', L.ptr(query), code, nq, ctypes.byref(self.spec), L.ptr(keys), L.ptr(ids), st)
            L.call_with_temp('dc_sort_pairs', dev, L.ptr(keys), L.ptr(qkeys), L.ptr(ids), L.ptr(qorder), nq,
                             self.key_bits, after=(st,))
            L.call('dc_gather_points', L.ptr(query), code, L.ptr(qorder), nq, L.ptr(Q), None, st)
        return Q, qkeys, qorder

    def occupancy(self):
        """Mean number of points per occupied cell."""
        if self.n == 0:
            return 0.0
        sb = int(self.spec.sub_bits)
        n_occ = int(((self.keys[1:] >> sb) != (self.keys[:-1] >> sb)).sum().item()) + 1
        return self.n / n_occ


def _n_slices(n):
    return (n + L.SLICE - 1) // L.SLICE


class Graph(object):
    """Neighbourhood graph in sorted space (sliced-ELL) over a SortedMap.

    rows = queries (sorted by the map's grid), columns = map points.  `symmetric` graphs (self-query
    radius graphs) are their own transpose; otherwise `transposed()` builds the reverse lists the
    backward kernel gathers over.
    """

    def __init__(self, smap, slice_ptr, ell_idx, n_rows, width, q_order=None, ell_d2=None, mode='radius',
                 symmetric=False, k=None, r=None):
        self.map = smap
        self.slice_ptr = slice_ptr
        self.ell_idx = ell_idx
        self.n_rows = n_rows
        self._width = width           # K of the reference layout (max row length / k); None = computed on demand
        self._in_degree = None
        self._Q = None                # sorted query records of a cross query (self query: the map's own records)
        self.q_order = q_order if q_order is not None else smap.order
        self.self_query = q_order is None
        self.ell_d2 = ell_d2
        self.mode = mode
        self.symmetric = symmetric
        self.k, self.r = k, r
        self._transposed = None
        self._rows_sorted = False
        self._step_cache = {}

    @property
    def width(self):
        if self._width is None:
            deg = self._in_degree if self._in_degree is not None else self.degrees()
            self._width = int(deg.max().item()) if self.n_rows > 0 else 0
        return self._width

    # ---- reference layout views ---------------------------------------------------------------
    def _sort_knn_rows(self):
        """kNN rows come out of the selection kernel unordered; the reference layout is distance-sorted."""
        if self.mode == 'knn' and not self._rows_sorted and self.n_rows > 0:
            if self.ell_d2 is None:
                # the search does not store distances; recompute them (bit-identical) for the export
                self.ell_d2 = torch.empty(self.ell_idx.numel(), dtype=torch.float64, device=self.map.device)
                Q = self.map.P if self._Q is None else self._Q
                L.call('dc_knn_distances', L.ptr(self.map.P), L.ptr(Q), self.width, L.ptr(self.ell_idx), self.n_rows,
                       L.ptr(self.ell_d2), L.stream())
            L.call('dc_knn_sort_rows', L.ptr(self.map.P), self.map.n, self.width, L.ptr(self.ell_idx), L.ptr(self.ell_d2), self.n_rows, L.stream())
        self._rows_sorted = True

    def neighbors(self):
        """int64 [n_rows, K] in original order, -1 = missing (what nearest_neighbors() returns)."""
        self._sort_knn_rows()
        dev = self.map.device
        K = self.width
        out = torch.empty((self.n_rows, K), dtype=torch.int64, device=dev)
        st = L.stream()
        if self.n_rows > 0 and K > 0:
            L.call('dc_ell_to_padded', L.ptr(self.slice_ptr), L.ptr(self.ell_idx), self.n_rows, L.ptr(self.map.order),
                   L.ptr(self.q_order), K, L.ptr(out), st)
            if self.mode == 'radius':
                # query_ball_point on many points returns index-sorted rows
                L.call_with_temp('dc_sort_rows', dev, L.ptr(out), self.n_rows, K, after=(st,))
        return out

    def distances(self):
        """fp64 [n_rows, k] (inf = missing) for kNN graphs, None for radius graphs (nearest_neighbors.py:51)."""
        if self.mode != 'knn':
            return None
        self._sort_knn_rows()
        out = torch.empty((self.n_rows, self.width), dtype=torch.float64, device=self.map.device)
        if self.n_rows > 0:
            L.call('dc_ell_to_dist', self.width, L.ptr(self.ell_d2), L.ptr(self.ell_idx), self.n_rows,
                   L.ptr(self.q_order), L.ptr(out), L.stream())
        return out

    def degrees(self):
        """Valid neighbours per row, sorted space, int32."""
        deg = torch.empty(self.n_rows, dtype=torch.int32, device=self.map.device)
        if self.n_rows > 0:
            L.call('dc_graph_degrees', L.ptr(self.slice_ptr), L.ptr(self.ell_idx), self.n_rows, L.ptr(deg), L.stream())
        return deg

    def valid_counts(self):
        """Valid neighbours per row in ORIGINAL order (filter_valid_neighbors, filters.py:184-193)."""
        out = torch.empty(self.n_rows, dtype=torch.int64, device=self.map.device)
        out[self.q_order.long()] = self.degrees().long()
        return out

    # ---- transpose ----------------------------------------------------------------------------
    def transposed(self):
        if self.symmetric:
            return self
        if self._transposed is not None:
            return self._transposed
        assert self.self_query, 'transpose is only defined for self-query graphs'
        dev = self.map.device
        st = L.stream()
        n = self.n_rows
        # one (dst << 32 | src) pair per ELL slot (padding: dst = n, sorts last); the radix sort is stable, so
        # sorting the dst bits alone gives every transposed row in a deterministic order, in 3-4 passes
        m = self.ell_idx.numel()
        pairs = L.scratch('pairs', m, torch.int64, dev)
        pairs_sorted = L.scratch('pairs_sorted', m, torch.int64, dev)
        L.call('dc_graph_edges', L.ptr(self.slice_ptr), L.ptr(self.ell_idx), n, None, L.ptr(pairs), st)
        bits = 32 + max(1, int(n).bit_length())
        L.call_with_temp('dc_sort_keys', dev, L.ptr(pairs), L.ptr(pairs_sorted), m, 32, bits, after=(st,))
        indeg = torch.empty(n, dtype=torch.int32, device=dev)
        sw = torch.zeros(_n_slices(n), dtype=torch.int32, device=dev)
        L.call('dc_transpose_widths', L.ptr(pairs_sorted), m, n, L.ptr(indeg), L.ptr(sw), st)
        sp = torch.empty(_n_slices(n) + 1, dtype=torch.int64, device=dev)
        L.call_with_temp('dc_ell_offsets', dev, L.ptr(sw), _n_slices(n), L.ptr(sp), after=(st,))
        total = int(sp[-1].item())           # the one host read-back of the transpose
        idx_t = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        L.call('dc_transpose_fill', L.ptr(pairs_sorted), m, n, L.ptr(sp), L.ptr(idx_t), st)
        self._transposed = Graph(self.map, sp, idx_t, n, None, mode='transposed', symmetric=False)
        self._transposed._in_degree = indeg
        # no back reference: a reference cycle would leave multi-GB graphs to the cyclic garbage collector,
        # whose timing makes the caching allocator fall back to cudaMalloc at random
        return self._transposed

    # ---- import -------------------------------------------------------------------------------
    @staticmethod
    def from_padded(smap, neighbors):
        """Build the sorted-space graph from a reference-layout int64 [N,K] tensor (self-query)."""
        assert neighbors.dim() == 2 and neighbors.shape[0] == smap.n
        neighbors = neighbors.contiguous().to(torch.int64)
        n, K = neighbors.shape
        dev = smap.device
        st = L.stream()
        sw = torch.zeros(_n_slices(n), dtype=torch.int32, device=dev)
        L.call('dc_padded_to_ell', L.ptr(neighbors), n, K, L.ptr(smap.order), L.ptr(smap.inv_order), L.ptr(sw), None, None, st)
        sp = torch.empty(_n_slices(n) + 1, dtype=torch.int64, device=dev)
        L.call_with_temp('dc_ell_offsets', dev, L.ptr(sw), _n_slices(n), L.ptr(sp), after=(st,))
        total = int(sp[-1].item())
        idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        L.call('dc_padded_to_ell', L.ptr(neighbors), n, K, L.ptr(smap.order), L.ptr(smap.inv_order), None, L.ptr(sp), L.ptr(idx), st)
        return Graph(smap, sp, idx, n, K, mode='imported', symmetric=False)


# (k, r, dtype, device) -> (n, cell, occupancy target) of the last estimate: a map that is searched again (training loops
# rebuild the graph of the same scans after pose updates) starts from the previous cell size and usually needs one
# confirming pass instead of three
_cell_hint = {}


def _knn_cell_size(points, k, r, bounds, use_hint=True):
    """Cell edge for kNN search: ~DC_KNN_OCC * k points per occupied cell (estimated from key-only sorts), refined on
    maps of >= KNN_MODEL_MIN_POINTS points by a cost model of the kernel evaluated on a sample of queries
    (_knn_cell_from_sample; DC_KNN_CELL=occ keeps the occupancy estimate).

    The estimate of the previous search is reused (no sort at all) only for a map of the same size (+-5 %) AND the
    same extent (+-10 % per axis): an unrelated map of similar size but different scale gets its own estimate."""
    lo, hi = bounds
    occ_env = os.environ.get('DC_KNN_OCC', KNN_OCC_DEFAULT)
    hint_key = (int(k), float(r) if r else None, points.dtype, str(points.device))
    hint = _cell_hint.get(hint_key) if use_hint else None
    ext = [max(h - l, 1e-12) for l, h in zip(lo, hi)]
    same_extent = hint is not None and all(0.9 * a <= b <= 1.1 * a for a, b in zip(hint[3], ext))
    if hint is not None and same_extent and 0.8 * hint[0] <= points.shape[0] <= 1.25 * hint[0] and hint[2] == occ_env:
        c0 = hint[1]
        if 0.95 * hint[0] <= points.shape[0] <= 1.05 * hint[0]:
            return c0           # same map: the cell size only steers speed, never the result
    elif r:
        c0 = float(r) * (1.0 + 1e-6)
    else:
        c0 = max(max(ext) / 256.0, 1e-9)
    # a disc of radius c on a surface with `occ` points per c^2 holds pi * occ points; the mean occupancy is dominated
    # by sparse cells (a typical QUERY sits in a cell 2x as full)
    target = max(float(occ_env) * k, 2.0)
    use_model = points.shape[0] >= KNN_MODEL_MIN_POINTS and os.environ.get('DC_KNN_CELL', 'model') == 'model'
    band = (0.6, 1.7) if use_model else (0.9, 1.11)       # the sample model refines a rough estimate: fewer sorts
    for _ in range(5):
        occ = SortedMap.occupancy_of(points, lo, hi, c0)
        if band[0] * target <= occ <= band[1] * target:
            break
        c0 = c0 * min(max(math.sqrt(target / occ), 1.0 / 8.0), 8.0)
        if r and c0 > r:
            c0 = float(r) * (1.0 + 1e-6)      # a hair above r: one ring always covers r
            break
    if use_model:
        c0 = _knn_cell_from_sample(points, int(k), r, bounds, c0)
    _cell_hint[hint_key] = (points.shape[0], c0, occ_env, ext)
    return c0


def _grid_cells(bounds, cell):
    """Number of cells of the search grid SortedMap.make_spec lays over `bounds` (no overflow check)."""
    n = 1
    for lo, hi in zip(*bounds):
        n *= int(math.floor((hi - (lo - 1e-3 * cell)) / cell)) + 1
    return n


def _ring_sequence(max_ring):
    """Rings the kNN kernels try in turn (dc_knn.cu: 1, 2, 3, 4, 8, 16, ... capped at max_ring)."""
    seq, rho = [], 1
    while True:
        seq.append(min(rho, max_ring))
        if rho >= max_ring:
            return seq
        rho = rho + 1 if rho < 4 else rho * 2


def _knn_cell_from_sample(points, k, r, bounds, c0):
    """Cell edge that minimises a cost model of the kNN kernel on a SAMPLE of the map's own queries.

    The mean occupancy the first estimate aims at says little about a map whose density varies by orders of magnitude
    (lidar: ~1 / range^2): the best cell of the 64-scan corridor holds d_k (the distance of the k-th neighbour) of 75 %
    of the queries, the best cell of a street map of HDL-64 scans only 50 % (tools/knn_cell_sweep.py).  So KNN_SAMPLE
    random points are searched exactly (cell c0), which gives each one's d_k and, through it, the surface density
    sigma = k / (pi d_k^2) around it; a query then costs, for every ring rho the kernel has to try,
    (2 rho + 1)^2 rows of sigma c^2 candidates + KNN_ROW_COST candidates' worth of row overhead, and it is finished by
    the first ring with d_k < rho c.  (The kernel's reach is rho c + the distance of the query to the nearest face of its
    cell, but a WARP goes on to the next ring as soon as one of its 32 queries does, and one of them always sits next to a
    face: fitted against the measured sweeps of seven maps, profiles/r2_knn_cell_sweep.log, the plain rho c criterion with
    three candidates per row picks a cell within 2 % of the best of every sweep; crediting the mean face distance c / 8, or
    its exact expectation, up to 17 % off.)  The cell minimising the mean cost over candidate sizes c0 / 2.5 ... 2 c0 is
    returned."""
    n = points.shape[0]
    dev = points.device
    sample = points[torch.randint(0, n, (KNN_SAMPLE,), device=dev, generator=_sample_generator(dev))]
    g = search(points, sample, k=k, r=r, cell=c0)
    d = g.distances()                                     # fp64 [m, k], inf = fewer than k within r
    # the sorted map of this search is the one the caller builds next when c0 stands: keep it for search()
    _sample_map[0] = (points.data_ptr(), points._version, tuple(points.shape), points.dtype, c0, bounds, g.map)
    del g
    return _knn_cell_of_distances(d, n, r, bounds, c0)


def _knn_cell_of_distances(d, n, r, bounds, c0):
    """The cost model of _knn_cell_from_sample on the sorted neighbour distances d (fp64 [m, k], inf = missing) of m sample
    queries of a map of n points (pure torch: runs on any device, tests/test_host_graph.py)."""
    dev = d.device
    fin = torch.isfinite(d)
    n_valid = fin.sum(dim=1).clamp_(min=1).double()
    dk = d[:, -1]
    have_k = fin[:, -1]
    if float(have_k.double().mean().item()) < 0.1:
        return c0                                          # hardly any query has k neighbours within r: nothing to model
    far = torch.where(fin, d, torch.zeros_like(d)).amax(dim=1)
    rad = torch.where(have_k, dk, torch.full_like(dk, float(r)) if r else far).clamp_(min=1e-12)
    sigma = n_valid / (math.pi * rad * rad)               # points per unit area of the surface around the query
    cand = [c0 * 2.0 ** (e / 12.0) for e in range(-16, 13)]          # c0 / 2.5 ... 2 c0, 6 % apart
    if r:
        cand = [c for c in cand if c <= float(r) * (1.0 + 1e-6)]
    # a grid too fine for the dense cell table (binary searches for every row of cells: 6x slower on the 57 M point street
    # map) is only considered when no candidate fits
    fits = [c for c in cand if _grid_cells(bounds, c) <= DENSE_TABLE_MAX_CELLS]
    if fits:
        cand = fits
        c0 = min(cand, key=lambda c: abs(c - c0))          # c0 itself may not fit
    if not cand:
        return c0
    cells = torch.tensor(cand, dtype=torch.float64, device=dev)
    cost = torch.zeros((d.shape[0], cells.numel()), dtype=torch.float64, device=dev)
    done = torch.zeros_like(cost, dtype=torch.bool)
    unit = sigma[:, None] * (cells * cells)[None, :] + KNN_ROW_COST
    max_ring = (torch.ceil(float(r) / cells) if r else torch.full_like(cells, 64.0))[None, :]
    for rho in _ring_sequence(64):
        rho_c = torch.clamp(torch.full_like(max_ring, float(rho)), max=max_ring)      # the kernel caps the ring at r / cell
        last = rho_c >= max_ring
        cost += torch.where(done, torch.zeros_like(cost), (2.0 * rho_c + 1.0) ** 2 * unit)
        done |= (dk[:, None] < rho_c * cells[None, :]) | last
        if bool(done.all()):
            break
    # + the dense cell table the search builds (n_cells + 1 int32 starts: a fill and a scan at HBM speed), per query and in
    # the model's unit (one candidate ~ 1.8 ps of kernel time per query of the map, one table entry ~ 1 ps)
    table = torch.tensor([_grid_cells(bounds, c) for c in cand], dtype=torch.float64, device=dev)
    table = torch.where(table <= DENSE_TABLE_MAX_CELLS, table, torch.zeros_like(table))
    mean = cost.mean(dim=0) + KNN_TABLE_COST * table / float(n)
    best = int(torch.argmin(mean).item())
    i0 = int(torch.argmin((cells - c0).abs()).item())
    # c0 stands (and the map of the sample search is reused) only when the minimum is c0 or one of its direct neighbours
    # (6 % apart): the model's curve is flatter than the kernel's, "within 5 % of the model's minimum" kept cells 12-20 %
    # slower than the best of the sweep on mid-size street maps
    return c0 if abs(best - i0) <= 1 else float(cells[best].item())


_sample_map = [None]                   # (data_ptr, version, shape, dtype, cell, bounds, SortedMap) of the last sample search


def _take_sample_map(points, cell, bounds):
    """The map the cell-size estimate has just built for these very points, cell and bounds (used once), else None."""
    e, _sample_map[0] = _sample_map[0], None
    if e is not None and e[:6] == (points.data_ptr(), points._version, tuple(points.shape), points.dtype, cell, bounds):
        return e[6]
    return None


_sample_generators = {}


def _sample_generator(dev):
    """Own seeded generator for the query sample: the cell size (speed only) does not consume the global RNG stream."""
    key = str(dev)
    if key not in _sample_generators:
        _sample_generators[key] = torch.Generator(device=dev)
    _sample_generators[key].manual_seed(12345)
    return _sample_generators[key]


def clear_cell_hints():
    """Forget the cell sizes remembered from earlier searches (the next search estimates its own: "cold" search)."""
    _cell_hint.clear()


def search(points, query=None, k=None, r=None, cell=None, stack_first=None):
    """Neighbour search -> Graph.  Modes follow nearest_neighbors.py:47-53:
    k (and optionally r as a strict upper bound) -> kNN graph; r only -> radius graph (<= r).

    stack_first (int64 [S+1] on the device): `points` holds S clouds one after the other (cloud s = rows
    stack_first[s] .. stack_first[s+1]); each is searched on its own (self query, needs r), all in the same launches."""
    assert k or r
    points = _as_points(points)
    n = points.shape[0]
    dev = points.device
    self_query = query is None or query is points
    bounds = SortedMap.bounds_of(points, None if self_query else query)
    stack = None
    if stack_first is not None:
        assert self_query and r, 'the stacked search is a self query with a radius (the guard band between clouds)'
        n_clouds = int(stack_first.numel()) - 1
    if cell is None:
        if k:
            if stack_first is not None and n > 0:
                # occupancy of ONE cloud (the clouds overlap in their common frame)
                n0 = max(n // n_clouds, 1)
                cell = _knn_cell_size(points[:n0], int(k), r, bounds, use_hint=False)
            else:
                cell = _knn_cell_size(points, int(k), r, bounds) if n > 0 else 1.0
        else:
            cell = float(r) * (1.0 + 1e-6)   # a hair above r: one ring of cells is always enough
    if stack_first is not None:
        stack = (stack_first, n_clouds, int(math.ceil(float(r) / cell)))
    smap = _take_sample_map(points, cell, bounds) if (stack is None and self_query) else None
    _sample_map[0] = None
    if smap is None:
        smap = SortedMap(points, cell, bounds=bounds, stack=stack)
    st = L.stream()
    if self_query:
        Q, qkeys, qorder, nq = smap.P, smap.keys, None, n
    else:
        Q, qkeys, qorder = smap.sort_queries(query)
        nq = Q.shape[0]
    ns = _n_slices(nq)
    spec = ctypes.byref(smap.spec)
    if k:
        k = int(k)
        sp = torch.arange(ns + 1, dtype=torch.int64, device=dev) * (L.SLICE * k)
        idx = torch.empty(max(ns * L.SLICE * k, 1), dtype=torch.int32, device=dev)
        # DC_KNN selects among three implementations that return identical rows (tests/test_gpu_round2.py): 'record'
        # (default: one distance pass, the emit pass replays the recorded bins), 'thread' (dc_knn: re-walks the rows;
        # also the one that can return distances), 'cells' (one warp per occupied cell)
        knn_impl = os.environ.get('DC_KNN', 'record')
        if knn_impl == 'record':
            L.call_with_temp('dc_knn_recorded', dev, L.ptr(smap.P), L.ptr(smap.keys), n, L.ptr(Q), L.ptr(qkeys), nq, spec,
                             L.ptr(smap.cell_start), k, float(r) if r else 0.0, L.ptr(idx), after=(st,))
        elif k <= 128 and knn_impl == 'cells':
            # second, independent implementation (one warp per occupied cell, fp32 classification + fp64 re-check):
            # bit-identical rows, not faster on lidar maps (profiles/r2_knn_cell_kernel.md) -- opt-in, used as a cross-check
            L.call_with_temp('dc_knn_cells', dev, L.ptr(smap.P), L.ptr(smap.keys), n, L.ptr(Q), L.ptr(qkeys), nq, spec,
                             L.ptr(smap.cell_start), k, float(r) if r else 0.0, L.ptr(idx), after=(st,))
        else:
            L.call('dc_knn', L.ptr(smap.P), L.ptr(smap.keys), n, L.ptr(Q), L.ptr(qkeys), nq, spec, L.ptr(smap.cell_start),
                   k, float(r) if r else 0.0, L.ptr(idx), None, st)
        g = Graph(smap, sp, idx, nq, k, q_order=qorder, ell_d2=None, mode='knn', symmetric=False, k=k, r=r)
        g._Q = None if self_query else Q
        return g
    counts = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)
    sw = torch.zeros(max(ns, 1), dtype=torch.int32, device=dev)
    L.call('dc_radius_count', L.ptr(smap.P), L.ptr(smap.keys), n, L.ptr(Q), L.ptr(qkeys), nq, spec,
           L.ptr(smap.cell_start), float(r), L.ptr(counts), L.ptr(sw), st)
    sp = torch.empty(ns + 1, dtype=torch.int64, device=dev)
    L.call_with_temp('dc_ell_offsets', dev, L.ptr(sw), ns, L.ptr(sp), after=(st,))
    total = int(sp[-1].item())
    idx = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
    L.call('dc_radius_fill', L.ptr(smap.P), L.ptr(smap.keys), n, L.ptr(Q), L.ptr(qkeys), nq, spec,
           L.ptr(smap.cell_start), float(r), L.ptr(sp), L.ptr(idx), st)
    g = Graph(smap, sp, idx, nq, None, q_order=qorder, mode='radius', symmetric=self_query, k=None, r=r)
    g._in_degree = counts[:nq]       # row lengths; the padded width K = max is only needed for export
    return g
