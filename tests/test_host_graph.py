"""Host-side logic of the neighbour search that needs no GPU: the ring schedule the cell-size model mirrors, the grid size
it limits, and the model itself on synthetic neighbour distances (graph._knn_cell_of_distances is pure torch)."""
import math

import numpy as np
import torch

from depth_correction_b200 import graph


def test_ring_sequence_follows_the_kernels():
    # dc_knn.cu: rho = 1; rho = rho < 4 ? rho + 1 : rho * 2; capped at max_ring, which is also the last ring tried
    assert graph._ring_sequence(1) == [1]
    assert graph._ring_sequence(3) == [1, 2, 3]
    assert graph._ring_sequence(13) == [1, 2, 3, 4, 8, 13]
    assert graph._ring_sequence(64) == [1, 2, 3, 4, 8, 16, 32, 64]


def test_grid_cells_equals_make_spec():
    rng = np.random.default_rng(0)
    for _ in range(50):
        lo = rng.uniform(-100, 100, 3)
        hi = lo + rng.uniform(0.0, 300, 3)
        cell = float(rng.uniform(0.02, 2.0))
        _, _, n_cells = graph.SortedMap.make_spec(lo.tolist(), hi.tolist(), cell)
        assert graph._grid_cells((lo.tolist(), hi.tolist()), cell) == n_cells


def _surface_distances(dk, k):
    """Sorted neighbour distances of queries on a surface of uniform density whose k-th neighbour lies at dk."""
    j = torch.arange(1, k + 1, dtype=torch.float64)
    return dk[:, None] * torch.sqrt(j / k)[None, :]


def test_cell_model_on_a_homogeneous_surface():
    k = 32
    dk = torch.full((4096,), 0.030, dtype=torch.float64) * torch.exp(0.05 * torch.randn(4096, dtype=torch.float64, generator=torch.Generator().manual_seed(1)))
    d = _surface_distances(dk, k)
    bounds = ([0.0, 0.0, 0.0], [60.0, 3.0, 3.0])
    for c0 in (0.02, 0.033, 0.06):
        cell = graph._knn_cell_of_distances(d, 8_000_000, 0.4, bounds, c0)
        # ring 1 must hold the k-th neighbour of most queries (d_k < 1.125 cell) without tripling the candidates
        assert 0.027 <= cell <= 0.045, (c0, cell)
    # a start inside 5 % of the optimum is kept as it is (the map built for the sample search is reused)
    best = graph._knn_cell_of_distances(d, 8_000_000, 0.4, bounds, 0.033)
    assert graph._knn_cell_of_distances(d, 8_000_000, 0.4, bounds, best) == best


def test_cell_model_keeps_the_grid_inside_the_dense_table():
    k = 32
    dk = torch.full((2048,), 0.030, dtype=torch.float64)
    d = _surface_distances(dk, k)
    bounds = ([0.0, -6.0, 0.0], [760.0, 6.0, 4.5])           # the 600-scan street map: 2^30 cells at 0.0337 m
    cell = graph._knn_cell_of_distances(d, 57_000_000, 0.4, bounds, 0.032)
    assert graph._grid_cells(bounds, cell) <= graph.DENSE_TABLE_MAX_CELLS
    assert cell < 0.05
    # no candidate fits (an absurdly large box): the model still answers, unconstrained
    huge = ([0.0, 0.0, 0.0], [1e5, 1e5, 1e3])
    cell = graph._knn_cell_of_distances(d, 57_000_000, 0.4, huge, 0.032)
    assert 0.02 <= cell <= 0.07


def test_cell_model_without_neighbours_or_radius():
    k = 8
    d = torch.full((512, k), math.inf, dtype=torch.float64)
    d[:, 0] = 0.0                                             # every query only finds itself within r
    assert graph._knn_cell_of_distances(d, 1_000_000, 0.4, ([0.0] * 3, [10.0] * 3), 0.1) == 0.1
    # no radius: the far tail of d_k (isolated points) must not be able to produce a non-finite cost
    dk = torch.cat([torch.full((1000,), 0.05, dtype=torch.float64), torch.full((24,), 30.0, dtype=torch.float64)])
    cell = graph._knn_cell_of_distances(_surface_distances(dk, k), 1_000_000, None, ([0.0] * 3, [100.0] * 3), 0.08)
    assert math.isfinite(cell) and 0.03 <= cell <= 0.16
