// Multi-GPU slab exchange: route every point record of the locally ingested scans to the
// rank that owns its slab and, as a read-only halo copy, to every rank whose slab lies within `halo` of it; and
// rebuild per-scan arrays from what arrives.  The reference is single-process and has no counterpart.
//
//   dc_route_count : destinations g_min .. g_max of every point from its coordinate along the split axis
//                    (b[g] - halo <= x < b[g+1] + halo), per-destination counts
//   dc_route_pack  : records {vp.xyz, dir.xyz, depth, inc_angle} (float32 x 8 or float64 x 8) and
//                    {scan id, row, model mask, owned} (int32 x 4) written contiguously per destination, straight
//                    from the per-scan arrays (scan pointer table of dc_pack_records_batched)
//   dc_route_keys  : (scan id << 32 | row) sort keys of the received index records
//   dc_route_unpack: received records gathered in (scan, row) order into field arrays
#include "dc_common.cuh"

struct dc_scan_ptrs {
  unsigned long long vps, dirs, depth, inc, mask;
};

#define DC_ROUTE_MAX_RANKS 64

struct dc_route_bounds {
  const double* inner;                  // DEVICE array: inner[g] = lower boundary of slab g + 1 (n_ranks - 1 values)
  int n_ranks;
  double halo;
};

__device__ __forceinline__ int dc_route_find_scan(const int64_t* __restrict__ first, int n_scans, int64_t i) {
  int lo = 0, hi = n_scans;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// number of inner boundaries <= v  (== torch.bucketize(v, inner, right=True))
__device__ __forceinline__ int dc_route_bucket(const dc_route_bounds& b, double v) {
  int c = 0;
  for (int g = 0; g < b.n_ranks - 1; ++g) c += (__ldg(b.inner + g) <= v);
  return c;
}

// scan_counts (optional): int32 [n_ranks][scan_stride], records per (destination, GLOBAL scan id) -- lets the receiver
// know the size of every scan it will hold without looking at the data (no read-back after the exchange)
__global__ void route_count_kernel(const double* __restrict__ wp, int axis, int64_t n, dc_route_bounds b,
                                   const int64_t* __restrict__ first, const int32_t* __restrict__ scan_ids, int n_scans,
                                   uint8_t* __restrict__ gmin, uint8_t* __restrict__ gmax, int32_t* __restrict__ counts,
                                   int32_t* __restrict__ scan_counts, int scan_stride) {
  __shared__ int s_cnt[DC_ROUTE_MAX_RANKS];
  if (threadIdx.x < DC_ROUTE_MAX_RANKS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const double x = wp[3 * i + axis];
    const int lo = dc_route_bucket(b, x - b.halo), hi = dc_route_bucket(b, x + b.halo);
    gmin[i] = (uint8_t)lo;
    gmax[i] = (uint8_t)hi;
    for (int g = lo; g <= hi; ++g) atomicAdd(&s_cnt[g], 1);
    if (scan_counts) {
      // consecutive points belong to the same scan and mostly to the same slab: aggregate equal (scan, destination)
      // pairs inside the warp before touching the global counter
      const int sid = scan_ids[dc_route_find_scan(first, n_scans, i)];
      for (int g = lo; g <= hi; ++g) {
        const unsigned key = (unsigned)sid * DC_ROUTE_MAX_RANKS + (unsigned)g;
        const unsigned peers = __match_any_sync(__activemask(), key);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&scan_counts[(int64_t)g * scan_stride + sid], __popc(peers));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < b.n_ranks && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]);
}

template <typename T>
__global__ void route_pack_kernel(const dc_scan_ptrs* __restrict__ tbl, const int64_t* __restrict__ first, const int32_t* __restrict__ scan_ids,
                                  int n_scans, int64_t n, const double* __restrict__ wp, int axis, dc_route_bounds b,
                                  const uint8_t* __restrict__ gmin, const uint8_t* __restrict__ gmax,
                                  const int64_t* __restrict__ dest_offset, int32_t* __restrict__ cursor,
                                  T* __restrict__ send_f, int32_t* __restrict__ send_i) {
  // Slots inside a destination's region are reserved per BLOCK: positions within the block from shared-memory counters,
  // one global atomic per (block, destination).  (One global atomicAdd per record on the n_ranks cursors serialised the
  // whole kernel: 5.7 ms for 8.4 M records, 74 % of the exchange.)
  __shared__ int s_cnt[DC_ROUTE_MAX_RANKS], s_base[DC_ROUTE_MAX_RANKS];
  if (threadIdx.x < DC_ROUTE_MAX_RANKS) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int pos[4] = {0, 0, 0, 0};
  if (i < n) {
    const int lo = gmin[i], hi = gmax[i];
    for (int g = lo; g <= hi && g - lo < 4; ++g) pos[g - lo] = atomicAdd(&s_cnt[g], 1);
  }
  __syncthreads();
  if (threadIdx.x < b.n_ranks) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]) : 0;
  __syncthreads();
  if (i >= n) return;
  const int s = dc_route_find_scan(first, n_scans, i);
  const int64_t li = i - first[s];
  const dc_scan_ptrs p = tbl[s];
  const T* vps = reinterpret_cast<const T*>(p.vps);
  const T* dirs = reinterpret_cast<const T*>(p.dirs);
  const T* depth = reinterpret_cast<const T*>(p.depth);
  const T* inc = reinterpret_cast<const T*>(p.inc);
  const uint8_t* mask = reinterpret_cast<const uint8_t*>(p.mask);
  T rec[8];
  rec[0] = vps ? vps[3 * li] : (T)0; rec[1] = vps ? vps[3 * li + 1] : (T)0; rec[2] = vps ? vps[3 * li + 2] : (T)0;
  rec[3] = dirs[3 * li]; rec[4] = dirs[3 * li + 1]; rec[5] = dirs[3 * li + 2];
  rec[6] = depth[li];
  rec[7] = inc ? inc[li] : (T)0;
  const int owner = dc_route_bucket(b, wp[3 * i + axis]);
  const int mm = (!mask || mask[li]) ? 1 : 0;
  for (int g = gmin[i]; g <= gmax[i]; ++g) {
    // (a point within the halo of more than four slabs -- slabs thinner than the halo -- reserves the rest one by one)
    const int64_t slot = dest_offset[g] + (g - gmin[i] < 4 ? s_base[g] + pos[g - gmin[i]] : atomicAdd(&cursor[g], 1));
    T* f = send_f + 8 * slot;
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = rec[k];
    int32_t* r = send_i + 4 * slot;
    r[0] = scan_ids[s]; r[1] = (int32_t)li; r[2] = mm; r[3] = (g == owner) ? 1 : 0;
  }
}

extern "C" int dc_route_count(const double* world_points, int axis, int64_t n, const double* inner_boundaries, int n_ranks,
                              double halo, const int64_t* first, const int32_t* scan_ids, int n_scans, uint8_t* gmin,
                              uint8_t* gmax, int32_t* counts, int32_t* scan_counts, int scan_stride, void* stream) {
  if (n_ranks < 1 || n_ranks > DC_ROUTE_MAX_RANKS) return dc_set_error(DC_ERR_ARG, "dc_route_count: 1..64 ranks");
  if (axis < 0 || axis > 2) return dc_set_error(DC_ERR_ARG, "dc_route_count: axis must be 0, 1 or 2");
  if (scan_counts && (!first || !scan_ids || n_scans < 1 || scan_stride < 1))
    return dc_set_error(DC_ERR_ARG, "dc_route_count: scan_counts needs the scan table");
  cudaStream_t st = (cudaStream_t)stream;
  DC_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_ranks, st));
  if (scan_counts) DC_CUDA_CHECK(cudaMemsetAsync(scan_counts, 0, sizeof(int32_t) * (size_t)n_ranks * scan_stride, st));
  if (n <= 0) return DC_OK;
  dc_route_bounds b;
  b.n_ranks = n_ranks;
  b.halo = halo;
  b.inner = inner_boundaries;      // DEVICE array (the slab plan stays on the device)
  route_count_kernel<<<dc_blocks(n, 256), 256, 0, st>>>(world_points, axis, n, b, first, scan_ids, n_scans, gmin, gmax, counts,
                                                        scan_counts, scan_stride);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_route_pack(const void* scan_ptr_table, const int64_t* first, const int32_t* scan_ids, int n_scans, int64_t n,
                             int dtype, const double* world_points, int axis, const double* inner_boundaries, int n_ranks,
                             double halo, const uint8_t* gmin, const uint8_t* gmax, const int64_t* dest_offset, int32_t* cursor,
                             void* send_f, int32_t* send_i, void* stream) {
  if (n_ranks < 1 || n_ranks > DC_ROUTE_MAX_RANKS) return dc_set_error(DC_ERR_ARG, "dc_route_pack: 1..64 ranks");
  cudaStream_t st = (cudaStream_t)stream;
  DC_CUDA_CHECK(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n_ranks, st));
  if (n <= 0) return DC_OK;
  dc_route_bounds b;
  b.n_ranks = n_ranks;
  b.halo = halo;
  b.inner = inner_boundaries;      // DEVICE array
  const dc_scan_ptrs* tbl = (const dc_scan_ptrs*)scan_ptr_table;
  if (dtype == DC_F32)
    route_pack_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>(tbl, first, scan_ids, n_scans, n, world_points, axis, b, gmin, gmax, dest_offset,
                                                                cursor, (float*)send_f, send_i);
  else
    route_pack_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>(tbl, first, scan_ids, n_scans, n, world_points, axis, b, gmin, gmax, dest_offset,
                                                                 cursor, (double*)send_f, send_i);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

__global__ void route_keys_kernel(const int32_t* __restrict__ recv_i, int64_t m, uint64_t* __restrict__ keys, int32_t* __restrict__ ids) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  keys[t] = ((uint64_t)(uint32_t)recv_i[4 * t] << 32) | (uint32_t)recv_i[4 * t + 1];
  ids[t] = (int32_t)t;
}

extern "C" int dc_route_keys(const int32_t* recv_i, int64_t m, uint64_t* keys, int32_t* ids, void* stream) {
  if (m <= 0) return DC_OK;
  route_keys_kernel<<<dc_blocks(m, 256), 256, 0, (cudaStream_t)stream>>>(recv_i, m, keys, ids);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

template <typename T>
__global__ void route_unpack_kernel(const T* __restrict__ recv_f, const int32_t* __restrict__ recv_i, const int32_t* __restrict__ order,
                                    int64_t m, T* __restrict__ vps, T* __restrict__ dirs, T* __restrict__ depth, T* __restrict__ inc,
                                    uint8_t* __restrict__ mask, uint8_t* __restrict__ owned, int64_t* __restrict__ gid) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int64_t src = order[t];
  const T* f = recv_f + 8 * src;
  const int32_t* r = recv_i + 4 * src;
  vps[3 * t] = f[0]; vps[3 * t + 1] = f[1]; vps[3 * t + 2] = f[2];
  dirs[3 * t] = f[3]; dirs[3 * t + 1] = f[4]; dirs[3 * t + 2] = f[5];
  depth[t] = f[6];
  inc[t] = f[7];
  mask[t] = (uint8_t)r[2];
  owned[t] = (uint8_t)r[3];
  gid[2 * t] = r[0];
  gid[2 * t + 1] = r[1];
}

extern "C" int dc_route_unpack(const void* recv_f, const int32_t* recv_i, const int32_t* order, int64_t m, int dtype, void* vps,
                               void* dirs, void* depth, void* inc, uint8_t* mask, uint8_t* owned, int64_t* gid, void* stream) {
  if (m <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    route_unpack_kernel<float><<<dc_blocks(m, 256), 256, 0, st>>>((const float*)recv_f, recv_i, order, m, (float*)vps, (float*)dirs, (float*)depth,
                                                                  (float*)inc, mask, owned, gid);
  else
    route_unpack_kernel<double><<<dc_blocks(m, 256), 256, 0, st>>>((const double*)recv_f, recv_i, order, m, (double*)vps, (double*)dirs,
                                                                   (double*)depth, (double*)inc, mask, owned, gid);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// histogram of the coordinate along `axis` (slab boundaries with equal point counts are read off its prefix sum)
__global__ void axis_hist_kernel(const double* __restrict__ wp, int axis, int64_t n, double a0, double scale, int n_bins,
                                 int32_t* __restrict__ hist) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long b = (long long)((wp[3 * i + axis] - a0) * scale);      // truncation like torch's .long()
  b = b < 0 ? 0 : (b > n_bins - 1 ? n_bins - 1 : b);
  atomicAdd(hist + b, 1);
}

extern "C" int dc_axis_histogram(const double* world_points, int axis, int64_t n, double a0, double scale, int n_bins,
                                 int32_t* hist, void* stream) {
  if (axis < 0 || axis > 2 || n_bins < 1) return dc_set_error(DC_ERR_ARG, "dc_axis_histogram: bad axis / bin count");
  cudaStream_t st = (cudaStream_t)stream;
  DC_CUDA_CHECK(cudaMemsetAsync(hist, 0, sizeof(int32_t) * n_bins, st));
  if (n <= 0) return DC_OK;
  axis_hist_kernel<<<dc_blocks(n, 256), 256, 0, st>>>(world_points, axis, n, a0, scale, n_bins, hist);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
