// Uniform-grid construction for the neighbour search (kernel 1, setup part): bounds, cell keys,
// radix sort (cub), gather into 32-byte fp64 records, optional dense cell table.
// Replaces the cKDTree build of nearest_neighbors.py:46.
#include <cub/cub.cuh>
#include <string.h>
#include <stdio.h>
#include "dc_common.cuh"
#include "dc_grid.cuh"

static thread_local char g_err[512] = "";

int dc_set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int dc_set_cuda_error(cudaError_t e, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
  return DC_ERR_CUDA;
}

extern "C" const char* dc_last_error(void) { return g_err; }
extern "C" int dc_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void bounds_kernel(const T* __restrict__ pts, int64_t n, double* out6, int32_t* bad_count) {
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int bad = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = (double)pts[3 * i], y = (double)pts[3 * i + 1], z = (double)pts[3 * i + 2];
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      mn[0] = fmin(mn[0], x); mn[1] = fmin(mn[1], y); mn[2] = fmin(mn[2], z);
      mx[0] = fmax(mx[0], x); mx[1] = fmax(mx[1], y); mx[2] = fmax(mx[2], z);
    } else {
      ++bad;
    }
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int a = 0; a < 3; ++a) {
    double v = BR(tmp).Reduce(mn[a], cub::Min());
    __syncthreads();
    double u = BR(tmp).Reduce(mx[a], cub::Max());
    __syncthreads();
    if (threadIdx.x == 0) {
      // fp64 atomic min/max through the ordered-integer trick is overkill here: one CAS loop per block
      unsigned long long* pm = (unsigned long long*)&out6[a];
      unsigned long long old = *pm, assumed;
      do {
        assumed = old;
        if (__longlong_as_double(assumed) <= v) break;
        old = atomicCAS(pm, assumed, __double_as_longlong(v));
      } while (assumed != old);
      unsigned long long* px = (unsigned long long*)&out6[3 + a];
      old = *px;
      do {
        assumed = old;
        if (__longlong_as_double(assumed) >= u) break;
        old = atomicCAS(px, assumed, __double_as_longlong(u));
      } while (assumed != old);
    }
  }
  typedef cub::BlockReduce<int, 256> BRI;
  __shared__ typename BRI::TempStorage tmpi;
  int b = BRI(tmpi).Sum(bad);
  if (threadIdx.x == 0 && b) atomicAdd(bad_count, b);
}

__global__ void bounds_init_kernel(double* out6, int32_t* bad_count) {
  if (threadIdx.x < 3) out6[threadIdx.x] = INFINITY;
  else if (threadIdx.x < 6) out6[threadIdx.x] = -INFINITY;
  if (threadIdx.x == 6) *bad_count = 0;
}

extern "C" int dc_bounds(const void* pts, int dtype, int64_t n, double* out6, int32_t* bad_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  bounds_init_kernel<<<1, 32, 0, st>>>(out6, bad_count);
  DC_LAUNCH_CHECK();
  if (n <= 0) return DC_OK;
  int blocks = dc_blocks(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dtype == DC_F32) bounds_kernel<float><<<blocks, 256, 0, st>>>((const float*)pts, n, out6, bad_count);
  else bounds_kernel<double><<<blocks, 256, 0, st>>>((const double*)pts, n, out6, bad_count);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void cell_keys_kernel(const T* __restrict__ pts, int64_t n, dc_grid g, uint64_t* keys, int32_t* ids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double p[3] = {(double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2]};
  int c0, c1, c2;
  dc_cell_coords(g, p, c0, c1, c2);
  uint64_t key = dc_cell_key(g, c0, c1, c2);
  if (g.sub_bits) {
    // 4x4x4 sub-cell of the point inside its cell, Morton-interleaved: the order of the points INSIDE a cell (any order
    // is correct for the search) becomes spatially coherent
    const double f0 = (p[g.ax[0]] - g.org[0]) * g.inv_cell - (double)c0, f1 = (p[g.ax[1]] - g.org[1]) * g.inv_cell - (double)c1,
                 f2 = (p[g.ax[2]] - g.org[2]) * g.inv_cell - (double)c2;
    const unsigned s0 = (unsigned)dc_clampi((int)(f0 * 4.0), 0, 3), s1 = (unsigned)dc_clampi((int)(f1 * 4.0), 0, 3),
                   s2 = (unsigned)dc_clampi((int)(f2 * 4.0), 0, 3);
    const unsigned m = (s0 & 1u) | ((s1 & 1u) << 1) | ((s2 & 1u) << 2) | ((s0 & 2u) << 2) | ((s1 & 2u) << 3) | ((s2 & 2u) << 4);
    key = (key << g.sub_bits) | (uint64_t)m;
  }
  keys[i] = key;
  ids[i] = (int32_t)i;
}

extern "C" int dc_cell_keys(const void* pts, int dtype, int64_t n, const dc_grid_spec* spec, uint64_t* keys,
                            int32_t* ids, void* stream) {
  if (n <= 0) return DC_OK;
  if (n > 2147483647LL) return dc_set_error(DC_ERR_OVERFLOW, "dc_cell_keys: more than 2^31-1 points");
  dc_grid g;
  int rc = dc_make_grid(spec, &g);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32) cell_keys_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)pts, n, g, keys, ids);
  else cell_keys_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)pts, n, g, keys, ids);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// Stacked form for many small clouds searched at once (per-scan features of all scans in one launch,
// preproc.py:35-64 / train.py:97-104): cloud s gets its own band of `period` cell layers along the slowest key axis,
// [s * period, s * period + band) with band = period - guard, so that no ring of cells of one cloud ever reaches the
// points of another.  Coordinates are untouched: distances are the true ones, bit for bit.
template <typename T>
__global__ void cell_keys_stacked_kernel(const T* __restrict__ pts, int64_t n, dc_grid g, const int64_t* __restrict__ first,
                                         int n_clouds, int period, int band, uint64_t* keys, int32_t* ids) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int lo = 0, hi = n_clouds;          // largest s with first[s] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= i) lo = mid; else hi = mid;
  }
  const double p[3] = {(double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2]};
  int c0, c1, c2;
  dc_cell_coords(g, p, c0, c1, c2);
  const double f2raw = (p[g.ax[2]] - g.org[2]) * g.inv_cell;
  c2 = dc_clampi((int)floor(f2raw), 0, band - 1);
  uint64_t key = dc_cell_key(g, c0, c1, c2 + lo * period);
  if (g.sub_bits) {
    const double f0 = (p[g.ax[0]] - g.org[0]) * g.inv_cell - (double)c0, f1 = (p[g.ax[1]] - g.org[1]) * g.inv_cell - (double)c1,
                 f2 = f2raw - (double)c2;
    const unsigned s0 = (unsigned)dc_clampi((int)(f0 * 4.0), 0, 3), s1 = (unsigned)dc_clampi((int)(f1 * 4.0), 0, 3),
                   s2 = (unsigned)dc_clampi((int)(f2 * 4.0), 0, 3);
    const unsigned m = (s0 & 1u) | ((s1 & 1u) << 1) | ((s2 & 1u) << 2) | ((s0 & 2u) << 2) | ((s1 & 2u) << 3) | ((s2 & 2u) << 4);
    key = (key << g.sub_bits) | (uint64_t)m;
  }
  keys[i] = key;
  ids[i] = (int32_t)i;
}

extern "C" int dc_cell_keys_stacked(const void* pts, int dtype, int64_t n, const dc_grid_spec* spec, const int64_t* first,
                                    int n_clouds, int period, int guard, uint64_t* keys, int32_t* ids, void* stream) {
  if (n <= 0) return DC_OK;
  if (n > 2147483647LL) return dc_set_error(DC_ERR_OVERFLOW, "dc_cell_keys_stacked: more than 2^31-1 points");
  dc_grid g;
  int rc = dc_make_grid(spec, &g);
  if (rc) return rc;
  if (n_clouds < 1 || guard < 1 || period <= guard || (int64_t)n_clouds * period != (int64_t)g.d[2])
    return dc_set_error(DC_ERR_ARG, "dc_cell_keys_stacked: the slowest grid dimension must equal n_clouds * period, period > guard >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    cell_keys_stacked_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)pts, n, g, first, n_clouds, period, period - guard, keys, ids);
  else
    cell_keys_stacked_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)pts, n, g, first, n_clouds, period, period - guard, keys, ids);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_make_grid(const dc_grid_spec* spec, dc_grid* g) {
  if (!spec || !(spec->cell > 0.0)) return dc_set_error(DC_ERR_ARG, "grid spec: cell size must be positive");
  int seen[3] = {0, 0, 0};
  for (int a = 0; a < 3; ++a) {
    if (spec->axis[a] < 0 || spec->axis[a] > 2 || seen[spec->axis[a]]) return dc_set_error(DC_ERR_ARG, "grid spec: axis must be a permutation of 0,1,2");
    seen[spec->axis[a]] = 1;
    if (spec->dims[a] < 1) return dc_set_error(DC_ERR_ARG, "grid spec: dims must be >= 1");
  }
  for (int a = 0; a < 3; ++a) {
    g->ax[a] = spec->axis[a];
    g->d[a] = spec->dims[spec->axis[a]];
    g->org[a] = spec->origin[spec->axis[a]];
  }
  if (spec->sub_bits != 0 && spec->sub_bits != 6) return dc_set_error(DC_ERR_ARG, "grid spec: sub_bits must be 0 or 6");
  g->sub_bits = spec->sub_bits;
  g->cell = spec->cell;
  g->inv_cell = 1.0 / spec->cell;
  const double cells = (double)g->d[0] * (double)g->d[1] * (double)g->d[2];
  if (cells > 9.0e18) return dc_set_error(DC_ERR_OVERFLOW, "grid spec: cell count overflows 63 bits");
  g->n_cells = (int64_t)g->d[0] * g->d[1] * g->d[2];
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// int32 array read as int64 * scale, 0 past the end (so a scan over n + 1 items also yields the total)
struct dc_pad_i32 {
  const int32_t* p;
  int64_t n;
  int64_t scale;
  __host__ __device__ int64_t operator()(int64_t i) const { return i < n ? (int64_t)p[i] * scale : 0; }
};

extern "C" int dc_sort_pairs(const uint64_t* keys_in, uint64_t* keys_out, const int32_t* ids_in, int32_t* ids_out,
                             int64_t n, int end_bit, void* temp, size_t* temp_bytes, void* stream) {
  if (end_bit < 1) end_bit = 1;
  if (end_bit > 64) end_bit = 64;
  DC_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, keys_in, keys_out, ids_in, ids_out, n, 0, end_bit,
                                                (cudaStream_t)stream));
  return DC_OK;
}

extern "C" int dc_sort_keys(const uint64_t* keys_in, uint64_t* keys_out, int64_t n, int begin_bit, int end_bit,
                            void* temp, size_t* temp_bytes, void* stream) {
  if (begin_bit < 0) begin_bit = 0;
  if (end_bit > 64) end_bit = 64;
  if (end_bit <= begin_bit) end_bit = begin_bit + 1;
  DC_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, keys_in, keys_out, n, begin_bit, end_bit,
                                               (cudaStream_t)stream));
  return DC_OK;
}

extern "C" int dc_exclusive_sum_i32_i64(const int32_t* in, int64_t* out, int64_t n, void* temp, size_t* temp_bytes,
                                        void* stream) {
  // out has n + 1 entries: out[n] = total (the input iterator yields 0 past the end)
  cub::CountingInputIterator<int64_t> cnt(0);
  cub::TransformInputIterator<int64_t, dc_pad_i32, cub::CountingInputIterator<int64_t>> it(cnt, dc_pad_i32{in, n, 1});
  DC_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, *temp_bytes, it, out, n + 1, (cudaStream_t)stream));
  return DC_OK;
}

extern "C" int dc_ell_offsets(const int32_t* slice_width, int64_t n_slices, int64_t* slice_ptr, void* temp,
                              size_t* temp_bytes, void* stream) {
  cub::CountingInputIterator<int64_t> cnt(0);
  cub::TransformInputIterator<int64_t, dc_pad_i32, cub::CountingInputIterator<int64_t>> it(cnt, dc_pad_i32{slice_width, n_slices, DC_SLICE});
  DC_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(temp, *temp_bytes, it, slice_ptr, n_slices + 1, (cudaStream_t)stream));
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_points_kernel(const T* __restrict__ pts, const int32_t* __restrict__ order, int64_t n,
                                     dc_point* __restrict__ out, int32_t* __restrict__ inv_order) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t o = order[s];
  dc_point p;
  p.x = (double)pts[3 * o]; p.y = (double)pts[3 * o + 1]; p.z = (double)pts[3 * o + 2];
  p.tag = o;
  out[s] = p;
  if (inv_order) inv_order[o] = (int32_t)s;
}

extern "C" int dc_gather_points(const void* pts, int dtype, const int32_t* order, int64_t n, void* sorted_points,
                                int32_t* inv_order, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    gather_points_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)pts, order, n, (dc_point*)sorted_points, inv_order);
  else
    gather_points_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)pts, order, n, (dc_point*)sorted_points, inv_order);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// cell_start[c] = first sorted position with key >= c, for c = 0 .. n_cells.  Position s "owns" the cells
// (key[s-1], key[s]] (and position n the cells above the last key); a warp fills the gaps of its 32 positions
// cooperatively, so a long run of empty cells costs one coalesced sweep instead of one thread's serial loop
// (the previous version did a 23-step binary search for each of the ~14 M cells of the bench grid).  On SPARSE grids (street
// maps: 20-40 cells per point) the gaps are 10^5 .. 10^6 cells long and each is filled by one warp: 1.4 TB/s, 0.83 ms for
// the 2.8e8-cell grid of one of eight slabs of the 57 M point map.  A kernel with one warp per 2048 CELLS (one coalesced store
// per 32 cells, keys searched only where a step holds any) was written and is correct, but its searches cost ~20 k cycles per
// occupied step (0.50 vs 0.11 ms on the dense bench grid) and it was never measured on a sparse grid: not adopted
// (profiles/r2_knn_experiments.md).
__global__ void cell_table_kernel(const uint64_t* __restrict__ keys, int64_t n, int64_t n_cells, int sub_bits, int32_t* __restrict__ cell_start) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;     // 0 .. n (inclusive), rounded up to a warp
  const int lane = threadIdx.x & 31;
  int64_t first = 0, last = -1;      // cells first .. last get the value s
  if (s <= n) {
    first = s == 0 ? 0 : (int64_t)(keys[s - 1] >> sub_bits) + 1;
    last = s == n ? n_cells : (int64_t)(keys[s] >> sub_bits);
  }
  const int64_t len = last - first + 1;
  if (len > 0 && len <= 4) {         // the common case: a handful of empty cells between occupied ones
    for (int64_t c = first; c <= last; ++c) cell_start[c] = (int32_t)s;
  }
  unsigned int big = __ballot_sync(0xffffffffu, len > 4);
  while (big) {
    const int src = __ffs(big) - 1;
    big &= big - 1u;
    const int64_t f = __shfl_sync(0xffffffffu, first, src), l = __shfl_sync(0xffffffffu, last, src);
    const int32_t v = (int32_t)__shfl_sync(0xffffffffu, s, src);
    for (int64_t c = f + lane; c <= l; c += 32) cell_start[c] = v;
  }
}

extern "C" int dc_cell_table(const uint64_t* keys_sorted, int64_t n, int64_t n_cells, int sub_bits, int32_t* cell_start, void* stream) {
  if (n_cells < 0 || n < 0) return dc_set_error(DC_ERR_ARG, "dc_cell_table: negative size");
  const int64_t threads = ((n + 1 + 31) / 32) * 32;
  cell_table_kernel<<<dc_blocks(threads, 256), 256, 0, (cudaStream_t)stream>>>(keys_sorted, n, n_cells, sub_bits, cell_start);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
