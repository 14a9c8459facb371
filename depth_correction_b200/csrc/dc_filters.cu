// Per-scan preprocessing filters: voxel-grid filter (filters.py:24-82) and
// shadow-point filter (filters.py:257-309) of the reference, as device kernels.
#include "dc_common.cuh"

// ---------------------------------------------------------------------------------------------
// filter_grid.  The reference keeps, for every occupied voxel, the LAST point of a (reversed / shuffled /
// unchanged) sequence -- a Python dict over tuple keys -- and returns the survivors in dict order, i.e. in the
// order in which the voxels FIRST appear in that sequence (or sorted by index with preserve_order).
//
// Here: position t of the sequence gets the key of its voxel (three 21-bit cell coordinates, computed in the
// cloud's own dtype exactly as numpy does: floor(x / grid_res)); a stable radix sort groups the voxels with the
// positions ascending inside every group; the last entry of a group then knows the survivor (its own position)
// and finds the first position of the group by a binary search on the sorted keys.
// ---------------------------------------------------------------------------------------------
#define VOX_BIAS (1 << 20)

template <typename T>
__global__ void voxel_keys_kernel(const T* __restrict__ pts, int64_t n, T grid_res, const int32_t* __restrict__ seq, int reversed,
                                  uint64_t* __restrict__ keys, int32_t* __restrict__ ids, int32_t* __restrict__ bad) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t i = seq ? (int64_t)seq[t] : (reversed ? n - 1 - t : t);
  uint64_t key = 0;
  bool ok = true;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const T q = floor(pts[3 * i + a] / grid_res);      // IEEE division in the cloud's dtype, like numpy
    ok = ok && (q >= (T)(-VOX_BIAS)) && (q < (T)VOX_BIAS);       // also false for NaN
    const long long c = ok ? (long long)q + VOX_BIAS : 0;
    key = (key << 21) | (uint64_t)c;
  }
  if (!ok) atomicAdd(bad, 1);
  keys[t] = key;
  ids[t] = (int32_t)t;
}

extern "C" int dc_voxel_keys(const void* points, int dtype, int64_t n, double grid_res, const int32_t* seq, int reversed,
                             uint64_t* keys, int32_t* ids, int32_t* bad, void* stream) {
  if (n <= 0) return DC_OK;
  if (!(grid_res > 0.0)) return dc_set_error(DC_ERR_ARG, "dc_voxel_keys: grid_res must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    voxel_keys_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)points, n, (float)grid_res, seq, reversed, keys, ids, bad);
  else
    voxel_keys_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)points, n, grid_res, seq, reversed, keys, ids, bad);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// out_key[s] = ordering key of the survivor of the voxel that ENDS at sorted position s (first position of the
// voxel in the sequence, or the survivor's point index with preserve_order), 2^32 for all other positions;
// out_val[s] = point index of the survivor; *count += number of voxels.
__global__ void voxel_pick_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ ids, int64_t n,
                                  const int32_t* __restrict__ seq, int reversed, int preserve_order,
                                  uint64_t* __restrict__ out_key, int32_t* __restrict__ out_val, int32_t* __restrict__ count) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint64_t key = keys[s];
  const bool tail = (s == n - 1) || (keys[s + 1] != key);
  if (!tail) {
    out_key[s] = 1ull << 32;
    out_val[s] = -1;
    return;
  }
  int64_t lo = 0, hi = s;              // first sorted position of this voxel
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  const int64_t t_last = ids[s], t_first = ids[lo];
  const int64_t point = seq ? (int64_t)seq[t_last] : (reversed ? n - 1 - t_last : t_last);
  out_key[s] = (uint64_t)(preserve_order ? point : t_first);
  out_val[s] = (int32_t)point;
  atomicAdd(count, 1);
}

extern "C" int dc_voxel_pick(const uint64_t* keys_sorted, const int32_t* ids_sorted, int64_t n, const int32_t* seq, int reversed,
                             int preserve_order, uint64_t* out_key, int32_t* out_val, int32_t* count, void* stream) {
  if (n <= 0) return DC_OK;
  voxel_pick_kernel<<<dc_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(keys_sorted, ids_sorted, n, seq, reversed, preserve_order,
                                                                        out_key, out_val, count);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// filter_shadow_points (filters.py:257-309): angle at x between (viewpoint - x) and (neighbour - x) for every
// neighbour in DIRECTION space (dir_neighbors); a point is kept when min and max of these angles lie within the
// bounds.  Invalid neighbours count as the mean of the bounds (always inside).  Angles in the cloud's dtype.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void shadow_mask_kernel(const T* __restrict__ pts, const T* __restrict__ vps, const int64_t* __restrict__ nbr,
                                   const float* __restrict__ nbr_w, int64_t n, int K, T a_lo, T a_hi, uint8_t* __restrict__ keep,
                                   T* __restrict__ a_min_out, T* __restrict__ a_max_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  const T ox = vps[3 * i] - x, oy = vps[3 * i + 1] - y, oz = vps[3 * i + 2] - z;
  const T on = sqrt(ox * ox + oy * oy + oz * oz);
  const T fill = (a_lo + a_hi) / (T)2;
  const T eps = (T)1e-8;               // torch.nn.functional.cosine_similarity clamps each norm at eps
  T amin = INFINITY, amax = -INFINITY;
  for (int c = 0; c < K; ++c) {
    const int64_t j = nbr[i * (int64_t)K + c];
    T a = fill;
    const bool valid = nbr_w ? (nbr_w[i * (int64_t)K + c] == 1.0f) : (j >= 0);
    if (valid) {
      const int64_t jj = j >= 0 ? j : j + n;      // (a weight of 1 with a negative index would wrap in torch)
      const T nx = pts[3 * jj] - x, ny = pts[3 * jj + 1] - y, nz = pts[3 * jj + 2] - z;
      const T nn = sqrt(nx * nx + ny * ny + nz * nz);
      const T cs = (ox * nx + oy * ny + oz * nz) / (fmax(on, eps) * fmax(nn, eps));
      a = acos(cs);
    }
    amin = fmin(amin, a);              // NaN angles (the point itself: 0/0 -> clamped to 0 -> acos(0)) handled like torch.amin
    amax = fmax(amax, a);
    if (a != a) { amin = a; amax = a; }
  }
  if (a_min_out) { a_min_out[i] = amin; a_max_out[i] = amax; }
  keep[i] = (amin >= a_lo) && (amax <= a_hi);
}

extern "C" int dc_shadow_mask(const void* points, const void* vps, int dtype, const int64_t* dir_neighbors,
                              const float* dir_neighbor_weights, int64_t n, int K, double angle_lo, double angle_hi,
                              uint8_t* keep, void* angle_min, void* angle_max, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    shadow_mask_kernel<float><<<dc_blocks(n, 128), 128, 0, st>>>((const float*)points, (const float*)vps, dir_neighbors, dir_neighbor_weights,
                                                                 n, K, (float)angle_lo, (float)angle_hi, keep, (float*)angle_min, (float*)angle_max);
  else
    shadow_mask_kernel<double><<<dc_blocks(n, 128), 128, 0, st>>>((const double*)points, (const double*)vps, dir_neighbors, dir_neighbor_weights,
                                                                  n, K, angle_lo, angle_hi, keep, (double*)angle_min, (double*)angle_max);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Neighbourhood statistics of global_cloud_mask (depth_cloud.py:330-354): weighted mean depth and weighted mean
// distance of the neighbours' viewpoints from their weighted mean, one pass over the padded lists (the reference
// materialises vps[neighbors] = [N,K,3] and three more [N,K] temporaries).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void neighbor_stats_kernel(const T* __restrict__ depth, const T* __restrict__ vps, const int64_t* __restrict__ nbr,
                                      const float* __restrict__ w, int64_t n, int K, T* __restrict__ mean_depth,
                                      T* __restrict__ mean_vp_dist) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t* row = nbr + i * (int64_t)K;
  const float* wr = w ? w + i * (int64_t)K : nullptr;
  T ws = 0, ds = 0, mx = 0, my = 0, mz = 0;
  for (int c = 0; c < K; ++c) {
    const int64_t j0 = row[c];
    const T wt = wr ? (T)wr[c] : (j0 >= 0 ? (T)1 : (T)0);
    const int64_t j = j0 >= 0 ? j0 : j0 + n;         // torch indexing wraps negative indices (their weight is 0)
    ws += wt;
    ds += wt * depth[j];
    if (vps) { mx += wt * vps[3 * j]; my += wt * vps[3 * j + 1]; mz += wt * vps[3 * j + 2]; }
  }
  if (mean_depth) mean_depth[i] = ds / ws;
  if (!mean_vp_dist) return;
  mx /= ws; my /= ws; mz /= ws;
  T acc = 0;
  for (int c = 0; c < K; ++c) {
    const int64_t j0 = row[c];
    const T wt = wr ? (T)wr[c] : (j0 >= 0 ? (T)1 : (T)0);
    const int64_t j = j0 >= 0 ? j0 : j0 + n;
    const T dx = vps[3 * j] - mx, dy = vps[3 * j + 1] - my, dz = vps[3 * j + 2] - mz;
    acc += wt * sqrt(dx * dx + dy * dy + dz * dz);
  }
  mean_vp_dist[i] = acc / ws;
}

extern "C" int dc_neighbor_stats(const void* depth, const void* vps, int dtype, const int64_t* neighbors, const float* weights,
                                 int64_t n, int K, void* mean_depth, void* mean_vp_dist, void* stream) {
  if (n <= 0) return DC_OK;
  if (mean_vp_dist && !vps) return dc_set_error(DC_ERR_ARG, "dc_neighbor_stats: mean_vp_dist needs vps");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    neighbor_stats_kernel<float><<<dc_blocks(n, 128), 128, 0, st>>>((const float*)depth, (const float*)vps, neighbors, weights, n, K,
                                                                    (float*)mean_depth, (float*)mean_vp_dist);
  else
    neighbor_stats_kernel<double><<<dc_blocks(n, 128), 128, 0, st>>>((const double*)depth, (const double*)vps, neighbors, weights, n, K,
                                                                     (double*)mean_depth, (double*)mean_vp_dist);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
