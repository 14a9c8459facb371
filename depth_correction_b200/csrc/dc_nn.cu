// Kernel 1: radius / kNN / kNN-within-r neighbour search on the cell-sorted map, plus graph
// format conversions (sliced-ELL <-> the reference's padded int64 [N,K]) and the transposed graph.
// Replaces cKDTree.query / query_ball_point and the Python padding loop of
// nearest_neighbors.py:46-73.
#include <cstdlib>
#include <cstring>
#include <cub/cub.cuh>
#include "dc_common.cuh"
#include "dc_grid.cuh"

#define NN_THREADS 128

// ---------------------------------------------------------------------------------------------
// Radius mode.  One thread per query (queries are cell-sorted, so a warp shares its candidate rows
// and the candidate records stay in L1).  FILL=false counts, FILL=true writes sorted-space indices
// in ascending order (rows of cells are visited in increasing key order).
// ---------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(NN_THREADS)
radius_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
              const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
              const int32_t* __restrict__ cell_start, double r2, int rings, int32_t* __restrict__ counts,
              int32_t* __restrict__ slice_width, const int64_t* __restrict__ slice_ptr, int32_t* __restrict__ ell_idx) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  int64_t base = 0;
  int width = 0;
  if (FILL && q < nq) {
    base = slice_ptr[q >> 5];
    width = (int)((slice_ptr[(q >> 5) + 1] - base) >> 5);
  }
  if (q < nq) {
    const dc_point pq = dc_ld_point(Q + q);
    int c0, c1, c2;
    dc_key_coords(g, qkeys[q], c0, c1, c2);
    for (int e2 = -rings; e2 <= rings; ++e2) {
      for (int e1 = -rings; e1 <= rings; ++e1) {
        int lo, hi;
        dc_row_range(g, pkeys, n, cell_start, c0 - rings, c0 + rings, c1 + e1, c2 + e2, lo, hi);
        // four candidate loads in flight (the scan is latency bound); the tail of a row reuses the body with
        // clamped addresses
        for (int j = lo; j < hi; j += 4) {
          const int last = hi - 1;
          const int j1 = j + 1 < last ? j + 1 : last, j2 = j + 2 < last ? j + 2 : last, j3 = j + 3 < last ? j + 3 : last;
          const dc_point p0 = dc_ld_point(P + j), p1 = dc_ld_point(P + j1);
          const dc_point p2 = dc_ld_point(P + j2), p3 = dc_ld_point(P + j3);
          const bool in0 = dc_dist2(p0, pq) <= r2, in1 = j + 1 < hi && dc_dist2(p1, pq) <= r2;
          const bool in2 = j + 2 < hi && dc_dist2(p2, pq) <= r2, in3 = j + 3 < hi && dc_dist2(p3, pq) <= r2;
          if (in0) { if (FILL) ell_idx[base + (int64_t)cnt * DC_SLICE + lane] = j; ++cnt; }
          if (in1) { if (FILL) ell_idx[base + (int64_t)cnt * DC_SLICE + lane] = j + 1; ++cnt; }
          if (in2) { if (FILL) ell_idx[base + (int64_t)cnt * DC_SLICE + lane] = j + 2; ++cnt; }
          if (in3) { if (FILL) ell_idx[base + (int64_t)cnt * DC_SLICE + lane] = j + 3; ++cnt; }
        }
      }
    }
    if (!FILL) counts[q] = cnt;
  }
  if (!FILL) {
    int m = cnt;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && q < nq) slice_width[q >> 5] = m;
  } else if (q < nq) {
    for (int c = cnt; c < width; ++c) ell_idx[base + (int64_t)c * DC_SLICE + lane] = -1;
  } else if ((q >> 5) <= ((nq - 1) >> 5) && nq > 0) {
    // tail lanes of the last (partial) slice: pad the whole column
    base = slice_ptr[q >> 5];
    width = (int)((slice_ptr[(q >> 5) + 1] - base) >> 5);
    for (int c = 0; c < width; ++c) ell_idx[base + (int64_t)c * DC_SLICE + lane] = -1;
  }
}

// Fill pass with coalesced list stores.  The sliced-ELL layout keeps the c-th neighbour of the 32 queries of a slice
// in one 128-byte line, so a lane that stores its own matches as it finds them touches a different line per store (one
// LSU wavefront each: the stores, not the distance tests, were 75 % of the fill pass).  Here every lane parks its
// matches in a private shared-memory column (RF_ROWS deep; bank = lane, no conflicts, no synchronisation: a lane only
// reads what it wrote) and the warp flushes row by row: for row R every lane that holds its R-th neighbour writes it,
// one partially-masked 128-byte store per row.  The loops over candidates are made warp-uniform (trip count = the
// longest row range of the warp) so that the flush can be taken by all lanes together.
#define RF_ROWS 64

__global__ void __launch_bounds__(NN_THREADS)
radius_fill_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                   const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
                   const int32_t* __restrict__ cell_start, double r2, int rings, const int64_t* __restrict__ slice_ptr,
                   int32_t* __restrict__ ell_idx) {
  __shared__ int32_t s_tile[NN_THREADS / 32][RF_ROWS * 32];
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if ((q >> 5) > ((nq - 1) >> 5)) return;                       // whole warp beyond the last slice
  int32_t* tile = s_tile[threadIdx.x >> 5] + lane;
  const int64_t base = slice_ptr[q >> 5];
  const int width = (int)((slice_ptr[(q >> 5) + 1] - base) >> 5);
  int32_t* out = ell_idx + base + lane;
  const bool live = q < nq;                                     // lanes past the end of the last slice only pad
  int cnt = 0, done = 0;                                        // matches found / already written (per lane)
  dc_point pq = {0.0, 0.0, 0.0, 0};
  int c0 = 0, c1 = 0, c2 = 0;
  if (live) {
    pq = dc_ld_point(Q + q);
    dc_key_coords(g, qkeys[q], c0, c1, c2);
  }
  for (int e2 = -rings; e2 <= rings; ++e2) {
    for (int e1 = -rings; e1 <= rings; ++e1) {
      int lo = 0, hi = 0;
      if (live) dc_row_range(g, pkeys, n, cell_start, c0 - rings, c0 + rings, c1 + e1, c2 + e2, lo, hi);
      const int trips = __reduce_max_sync(0xffffffffu, (hi - lo + 3) >> 2);
      for (int it = 0; it < trips; ++it) {
        const int j = lo + 4 * it;
        if (j < hi) {
          const int last = hi - 1;
          const int j1 = j + 1 < last ? j + 1 : last, j2 = j + 2 < last ? j + 2 : last, j3 = j + 3 < last ? j + 3 : last;
          const dc_point p0 = dc_ld_point(P + j), p1 = dc_ld_point(P + j1);
          const dc_point p2 = dc_ld_point(P + j2), p3 = dc_ld_point(P + j3);
          const bool in0 = dc_dist2(p0, pq) <= r2, in1 = j + 1 < hi && dc_dist2(p1, pq) <= r2;
          const bool in2 = j + 2 < hi && dc_dist2(p2, pq) <= r2, in3 = j + 3 < hi && dc_dist2(p3, pq) <= r2;
          if (in0) { tile[(cnt - done) * 32] = j; ++cnt; }
          if (in1) { tile[(cnt - done) * 32] = j + 1; ++cnt; }
          if (in2) { tile[(cnt - done) * 32] = j + 2; ++cnt; }
          if (in3) { tile[(cnt - done) * 32] = j + 3; ++cnt; }
        }
        if (__any_sync(0xffffffffu, cnt - done > RF_ROWS - 4)) {
          const int r_lo = __reduce_min_sync(0xffffffffu, done), r_hi = __reduce_max_sync(0xffffffffu, cnt);
          for (int R = r_lo; R < r_hi; ++R)
            if (R >= done && R < cnt) out[(int64_t)R * DC_SLICE] = tile[(R - done) * 32];
          done = cnt;
        }
      }
    }
  }
  // remaining matches, then the -1 padding, still row by row
  const int r_lo = __reduce_min_sync(0xffffffffu, done);
  for (int R = r_lo; R < width; ++R)
    if (R >= done) out[(int64_t)R * DC_SLICE] = R < cnt ? tile[(R - done) * 32] : -1;
}

static int radius_launch(bool fill, const void* P, const uint64_t* pkeys, int64_t n, const void* Q,
                         const uint64_t* qkeys, int64_t nq, const dc_grid_spec* spec, const int32_t* cell_start,
                         double r, int32_t* counts, int32_t* slice_width, const int64_t* slice_ptr, int32_t* ell_idx,
                         void* stream) {
  if (nq <= 0) return DC_OK;
  if (!(r > 0.0)) return dc_set_error(DC_ERR_ARG, "radius search: r must be positive");
  dc_grid g;
  int rc = dc_make_grid(spec, &g);
  if (rc) return rc;
  const int rings = (int)ceil(r / g.cell);
  // cKDTree compares the squared distance against r*r (distance_upper_bound ** p)
  const double r2 = r * r;
  const int blocks = dc_blocks(((nq + 31) / 32) * 32, NN_THREADS);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool direct_stores = getenv("DC_RADIUS_FILL") && !strcmp(getenv("DC_RADIUS_FILL"), "direct");   // A/B switch
  if (fill && !direct_stores)
    radius_fill_kernel<<<blocks, NN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g, cell_start,
                                                       r2, rings, slice_ptr, ell_idx);
  else if (fill)
    radius_kernel<true><<<blocks, NN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g,
                                                        cell_start, r2, rings, counts, slice_width, slice_ptr, ell_idx);
  else
    radius_kernel<false><<<blocks, NN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g,
                                                         cell_start, r2, rings, counts, slice_width, slice_ptr, ell_idx);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_radius_count(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys,
                               int64_t nq, const dc_grid_spec* spec, const int32_t* cell_start, double r,
                               int32_t* counts, int32_t* slice_width, void* stream) {
  return radius_launch(false, P, pkeys, n, Q, qkeys, nq, spec, cell_start, r, counts, slice_width, nullptr, nullptr, stream);
}

extern "C" int dc_radius_fill(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys,
                              int64_t nq, const dc_grid_spec* spec, const int32_t* cell_start, double r,
                              const int64_t* slice_ptr, int32_t* ell_idx, void* stream) {
  return radius_launch(true, P, pkeys, n, Q, qkeys, nq, spec, cell_start, r, nullptr, nullptr, slice_ptr, ell_idx, stream);
}

// (kNN / kNN within r: dc_knn.cu)

// ---------------------------------------------------------------------------------------------
// Export / import between sliced-ELL (sorted space, int32) and the reference layout
// (original order, padded int64 [N,K], -1 = missing; nearest_neighbors.py:69-78).
// ---------------------------------------------------------------------------------------------
__global__ void ell_to_padded_kernel(const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ ell_idx,
                                     int64_t nq, const int32_t* __restrict__ order_p,
                                     const int32_t* __restrict__ order_q, int K, int64_t* __restrict__ out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= nq) return;
  const int lane = (int)(row & 31);
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  int64_t* o = out + (int64_t)order_q[row] * K;
  for (int c = 0; c < K; ++c) {
    int64_t v = -1;
    if (c < width) {
      const int j = ell_idx[base + (int64_t)c * DC_SLICE + lane];
      if (j >= 0) v = order_p[j];
    }
    o[c] = v;
  }
}

extern "C" int dc_ell_to_padded(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t nq, const int32_t* order_p,
                                const int32_t* order_q, int K, int64_t* out, void* stream) {
  if (nq <= 0 || K <= 0) return DC_OK;
  ell_to_padded_kernel<<<dc_blocks(nq, 128), 128, 0, (cudaStream_t)stream>>>(slice_ptr, ell_idx, nq, order_p, order_q, K, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

__global__ void ell_to_dist_kernel(int k, const double* __restrict__ ell_d2, const int32_t* __restrict__ ell_idx,
                                   int64_t nq, const int32_t* __restrict__ order_q, double* __restrict__ out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= nq) return;
  const int lane = (int)(row & 31);
  const int64_t base = (row >> 5) * (int64_t)k * DC_SLICE;
  double* o = out + (int64_t)order_q[row] * k;
  for (int c = 0; c < k; ++c) {
    const int64_t e = base + (int64_t)c * DC_SLICE + lane;
    o[c] = ell_idx[e] >= 0 ? sqrt(ell_d2[e]) : INFINITY;
  }
}

extern "C" int dc_ell_to_dist(int k, const double* ell_d2, const int32_t* ell_idx, int64_t nq, const int32_t* order_q,
                              double* out, void* stream) {
  if (nq <= 0) return DC_OK;
  ell_to_dist_kernel<<<dc_blocks(nq, 128), 128, 0, (cudaStream_t)stream>>>(k, ell_d2, ell_idx, nq, order_q, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

struct dc_row_offset {
  int64_t K;
  __host__ __device__ int64_t operator()(int64_t i) const { return i * K; }
};

extern "C" int dc_sort_rows(int64_t* rows, int64_t n, int K, void* temp, size_t* temp_bytes, void* stream) {
  // -1 reinterpreted as uint64 is the maximum key, so padding stays at the end of every row
  typedef dc_row_offset Off;
  if ((double)n * (double)K > 2.0e9) return dc_set_error(DC_ERR_OVERFLOW, "dc_sort_rows: more than 2^31 entries");
  cub::CountingInputIterator<int64_t> cnt(0);
  cub::TransformInputIterator<int64_t, Off, cub::CountingInputIterator<int64_t>> begin(cnt, Off{K});
  cub::TransformInputIterator<int64_t, Off, cub::CountingInputIterator<int64_t>> end(cnt + 1, Off{K});
  uint64_t* keys = (uint64_t*)rows;
  // in-place through the double-buffer-free API needs a second buffer: temp holds [cub temp | copy]
  size_t cub_bytes = 0;
  DC_CUDA_CHECK(cub::DeviceSegmentedSort::SortKeys(nullptr, cub_bytes, keys, keys, (int)(n * K), (int)n, begin, end,
                                                   (cudaStream_t)stream));
  cub_bytes = (cub_bytes + 255) & ~(size_t)255;
  const size_t copy_bytes = (size_t)n * K * sizeof(uint64_t);
  if (!temp) { *temp_bytes = cub_bytes + copy_bytes + 256; return DC_OK; }
  if (*temp_bytes < cub_bytes + copy_bytes) return dc_set_error(DC_ERR_ARG, "dc_sort_rows: temp too small");
  uint64_t* copy = (uint64_t*)((char*)temp + cub_bytes);
  DC_CUDA_CHECK(cudaMemcpyAsync(copy, keys, copy_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  DC_CUDA_CHECK(cub::DeviceSegmentedSort::SortKeys(temp, cub_bytes, copy, keys, (int)(n * K), (int)n, begin, end,
                                                   (cudaStream_t)stream));
  return DC_OK;
}

// pass 1 (ell_idx == NULL): slice widths; pass 2: fill.  Rows of `neighbors` are in original order.
__global__ void padded_to_ell_kernel(const int64_t* __restrict__ neighbors, int64_t n, int K,
                                     const int32_t* __restrict__ order, const int32_t* __restrict__ inv_order,
                                     int32_t* __restrict__ slice_width, const int64_t* __restrict__ slice_ptr,
                                     int32_t* __restrict__ ell_idx) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // sorted-space row
  const int lane = threadIdx.x & 31;
  const bool live = row < n;
  const int64_t* src = live ? neighbors + (int64_t)order[row] * K : nullptr;
  if (!ell_idx) {
    int cnt = 0;
    if (live) for (int c = 0; c < K; ++c) cnt += (src[c] >= 0 && src[c] < n);
    for (int o = 16; o > 0; o >>= 1) cnt = max(cnt, __shfl_xor_sync(0xffffffffu, cnt, o));
    if (lane == 0 && live) slice_width[row >> 5] = cnt;
    return;
  }
  if ((row >> 5) > ((n - 1) >> 5)) return;
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  int cnt = 0;
  if (live)
    for (int c = 0; c < K; ++c) {
      const int64_t v = src[c];
      if (v >= 0 && v < n) ell_idx[base + (int64_t)(cnt++) * DC_SLICE + lane] = inv_order[v];
    }
  for (int c = cnt; c < width; ++c) ell_idx[base + (int64_t)c * DC_SLICE + lane] = -1;
}

extern "C" int dc_padded_to_ell(const int64_t* neighbors, int64_t n, int K, const int32_t* order,
                                const int32_t* inv_order, int32_t* slice_width, const int64_t* slice_ptr,
                                int32_t* ell_idx, void* stream) {
  if (n <= 0) return DC_OK;
  const int blocks = dc_blocks(((n + 31) / 32) * 32, 128);
  padded_to_ell_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(neighbors, n, K, order, inv_order, slice_width, slice_ptr, ell_idx);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Transposed graph: row j of the transpose lists every i with j in N(i) (ascending i).
// ---------------------------------------------------------------------------------------------
__global__ void graph_degrees_kernel(const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ ell_idx,
                                     int64_t n_rows, int32_t* __restrict__ out_degree) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  const int lane = (int)(row & 31);
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  int cnt = 0;
  for (int c = 0; c < width; ++c) cnt += ell_idx[base + (int64_t)c * DC_SLICE + lane] >= 0;
  out_degree[row] = cnt;
}

extern "C" int dc_graph_degrees(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t n_rows, int32_t* out_degree,
                                void* stream) {
  if (n_rows <= 0) return DC_OK;
  graph_degrees_kernel<<<dc_blocks(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(slice_ptr, ell_idx, n_rows, out_degree);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

__global__ void graph_edges_kernel(const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ ell_idx,
                                   int64_t n_rows, const int64_t* __restrict__ edge_offset, uint64_t* __restrict__ pairs) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ((row >> 5) > ((n_rows - 1) >> 5)) return;
  const int lane = (int)(row & 31);
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  if (!edge_offset) {
    // slot mode: one pair per ELL slot (coalesced, no prefix sum); padding gets dst = n_rows, which sorts last
    const uint64_t pad = ((uint64_t)(uint32_t)n_rows << 32) | 0xffffffffull;
    for (int c = 0; c < width; ++c) {
      const int64_t e = base + (int64_t)c * DC_SLICE + lane;
      const int j = row < n_rows ? ell_idx[e] : -1;
      pairs[e] = j >= 0 ? (((uint64_t)(uint32_t)j << 32) | (uint64_t)(uint32_t)row) : pad;
    }
    return;
  }
  if (row >= n_rows) return;
  int64_t e = edge_offset[row];
  for (int c = 0; c < width; ++c) {
    const int j = ell_idx[base + (int64_t)c * DC_SLICE + lane];
    if (j >= 0) pairs[e++] = ((uint64_t)(uint32_t)j << 32) | (uint64_t)(uint32_t)row;
  }
}

extern "C" int dc_graph_edges(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t n_rows,
                              const int64_t* edge_offset, uint64_t* pairs, void* stream) {
  if (n_rows <= 0) return DC_OK;
  const int blocks = dc_blocks(((n_rows + 31) / 32) * 32, 128);
  graph_edges_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(slice_ptr, ell_idx, n_rows, edge_offset, pairs);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

__global__ void transpose_kernel(const uint64_t* __restrict__ pairs, int64_t n_edges, int64_t n_cols,
                                 int32_t* __restrict__ in_degree, int32_t* __restrict__ slice_width,
                                 const int64_t* __restrict__ slice_ptr_t, int32_t* __restrict__ ell_idx_t) {
  const int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = col < n_cols;
  int64_t lo = 0, hi = 0;
  if (live) {
    lo = dc_lower_bound(pairs, n_edges, (uint64_t)col << 32, 0);
    hi = dc_lower_bound(pairs, n_edges, (uint64_t)(col + 1) << 32, 0);
  }
  const int deg = (int)(hi - lo);
  if (!ell_idx_t) {
    if (live) in_degree[col] = deg;
    int m = deg;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && live) slice_width[col >> 5] = m;
    return;
  }
  if (n_cols <= 0 || (col >> 5) > ((n_cols - 1) >> 5)) return;
  const int64_t base = slice_ptr_t[col >> 5];
  const int width = (int)((slice_ptr_t[(col >> 5) + 1] - base) >> 5);
  for (int c = 0; c < width; ++c)
    ell_idx_t[base + (int64_t)c * DC_SLICE + lane] = c < deg ? (int32_t)(uint32_t)(pairs[lo + c] & 0xffffffffu) : -1;
}

extern "C" int dc_transpose_widths(const uint64_t* pairs_sorted, int64_t n_edges, int64_t n_cols, int32_t* in_degree,
                                   int32_t* slice_width, void* stream) {
  if (n_cols <= 0) return DC_OK;
  const int blocks = dc_blocks(((n_cols + 31) / 32) * 32, 128);
  transpose_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pairs_sorted, n_edges, n_cols, in_degree, slice_width, nullptr, nullptr);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_transpose_fill(const uint64_t* pairs_sorted, int64_t n_edges, int64_t n_cols,
                                 const int64_t* slice_ptr_t, int32_t* ell_idx_t, void* stream) {
  if (n_cols <= 0) return DC_OK;
  const int blocks = dc_blocks(((n_cols + 31) / 32) * 32, 128);
  transpose_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pairs_sorted, n_edges, n_cols, nullptr, nullptr, slice_ptr_t, ell_idx_t);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
