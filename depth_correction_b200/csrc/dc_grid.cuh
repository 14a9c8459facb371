// Device-side view of the uniform search grid (axes permuted so that axis 0 is the fastest
// varying key digit and axis 2 -- the longest extent of the map -- the slowest, which keeps a
// slab perpendicular to the trajectory contiguous in sorted order).
#pragma once
#include "dc_common.cuh"

struct dc_grid {
  double org[3];     // origin in permuted axes
  double cell, inv_cell;
  int d[3];          // dims in permuted axes (d[0] fastest)
  int ax[3];         // permuted axis a reads xyz component ax[a]
  int sub_bits;      // low key bits holding the sub-cell Morton code (0 or 6); cell = key >> sub_bits
  int64_t n_cells;
};

int dc_make_grid(const dc_grid_spec* spec, dc_grid* g);

__device__ __forceinline__ int dc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// cell coordinates (permuted axes), clamped into the grid so that out-of-box queries still work
__device__ __forceinline__ void dc_cell_coords(const dc_grid& g, const double p[3], int& c0, int& c1, int& c2) {
  c0 = dc_clampi((int)floor((p[g.ax[0]] - g.org[0]) * g.inv_cell), 0, g.d[0] - 1);
  c1 = dc_clampi((int)floor((p[g.ax[1]] - g.org[1]) * g.inv_cell), 0, g.d[1] - 1);
  c2 = dc_clampi((int)floor((p[g.ax[2]] - g.org[2]) * g.inv_cell), 0, g.d[2] - 1);
}

__device__ __forceinline__ uint64_t dc_cell_key(const dc_grid& g, int c0, int c1, int c2) {
  return ((uint64_t)c2 * (uint64_t)g.d[1] + (uint64_t)c1) * (uint64_t)g.d[0] + (uint64_t)c0;
}

__device__ __forceinline__ void dc_key_coords(const dc_grid& g, uint64_t key, int& c0, int& c1, int& c2) {
  key >>= g.sub_bits;
  c0 = (int)(key % (uint64_t)g.d[0]);
  const uint64_t t = key / (uint64_t)g.d[0];
  c1 = (int)(t % (uint64_t)g.d[1]);
  c2 = (int)(t / (uint64_t)g.d[1]);
}

// first position in sorted keys[0..n) whose cell (keys[pos] >> sub_bits) is >= cell
__device__ __forceinline__ int64_t dc_lower_bound(const uint64_t* __restrict__ keys, int64_t n, uint64_t cell, int sub_bits) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((__ldg(keys + mid) >> sub_bits) < cell) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Sorted-position range [lo, hi) of all points in cells (c0lo..c0hi, c1, c2): a row of cells along the
// fastest axis is one contiguous run of the sorted map.
__device__ __forceinline__ void dc_row_range(const dc_grid& g, const uint64_t* __restrict__ keys, int64_t n,
                                             const int32_t* __restrict__ cell_start, int c0lo, int c0hi, int c1, int c2,
                                             int& lo, int& hi) {
  if (c1 < 0 || c1 >= g.d[1] || c2 < 0 || c2 >= g.d[2]) { lo = hi = 0; return; }
  c0lo = c0lo < 0 ? 0 : c0lo;
  c0hi = c0hi >= g.d[0] ? g.d[0] - 1 : c0hi;
  if (c0lo > c0hi) { lo = hi = 0; return; }
  const uint64_t base = ((uint64_t)c2 * (uint64_t)g.d[1] + (uint64_t)c1) * (uint64_t)g.d[0];
  const uint64_t k0 = base + (uint64_t)c0lo, k1 = base + (uint64_t)c0hi + 1;
  if (cell_start) {
    lo = __ldg(cell_start + k0);
    hi = __ldg(cell_start + k1);
  } else {
    lo = (int)dc_lower_bound(keys, n, k0, g.sub_bits);
    hi = (int)dc_lower_bound(keys, n, k1, g.sub_bits);
  }
}

// The same range without branches around the loads (cell table only): an invalid row reads entry 0 twice and yields
// lo = hi = 0.  Lets a caller issue the lookups of several rows back to back instead of one dependent pair per row.
__device__ __forceinline__ void dc_row_range_nb(const dc_grid& g, const int32_t* __restrict__ cell_start, int c0lo, int c0hi,
                                                int c1, int c2, int& lo, int& hi) {
  c0lo = c0lo < 0 ? 0 : c0lo;
  c0hi = c0hi >= g.d[0] ? g.d[0] - 1 : c0hi;
  const bool ok = c1 >= 0 && c1 < g.d[1] && c2 >= 0 && c2 < g.d[2] && c0lo <= c0hi;
  // a dense cell table exists for at most 2^30 cells (graph.py: DENSE_TABLE_MAX_CELLS): 32-bit index arithmetic
  const unsigned int base = ((unsigned int)c2 * (unsigned int)g.d[1] + (unsigned int)c1) * (unsigned int)g.d[0];
  const unsigned int k0 = ok ? base + (unsigned int)c0lo : 0u, k1 = ok ? base + (unsigned int)c0hi + 1u : 0u;
  lo = __ldg(cell_start + k0);
  hi = __ldg(cell_start + k1);
}

// Squared distance exactly as cKDTree's p=2 kernel accumulates it for m=3: ((dx*dx + dy*dy) + dz*dz)
// with every product and sum rounded separately (no FMA contraction), so that `<= r*r` / `< r*r`
// decide identically for boundary points.
__device__ __forceinline__ double dc_dist2(const dc_point& a, const dc_point& b) {
  const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
