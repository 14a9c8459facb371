// ICP-style losses between consecutive scans (loss.py:373-565 of the reference):
// point-to-plane / point-to-point residuals over the correspondences (i, nn(i)) whose distance is below the
// inlier quantile, reduced to one scalar per pair of scans, and the hand-written backward to both point sets
// and both normal sets.  The correspondences come from dc_knn (k = 1, query != points).
//
// The reference casts the points to float32 first (loss.py:424-425, 509-510); the kernels round the coordinates
// to float32 the same way and then work in fp64.
#include <cub/cub.cuh>
#include "dc_common.cuh"

#define ICP_THREADS 256

template <typename T> __device__ __forceinline__ double icp_f32(T v) { return (double)(float)v; }

struct icp_pair {
  double p1[3], p2[3];
  long long i, j;
  bool on;
};

// correspondence t: explicit lists (sel1[t], sel2[t]) or (t, nn[t]) kept when dist[t] <= threshold (NaN: dropped)
template <typename T>
__device__ __forceinline__ icp_pair icp_load(const T* __restrict__ pts1, const T* __restrict__ pts2,
                                             const int64_t* __restrict__ nn, const double* __restrict__ dist, double th,
                                             const int64_t* __restrict__ sel1, const int64_t* __restrict__ sel2, int64_t t) {
  icp_pair r;
  r.on = true;
  if (sel1) { r.i = sel1[t]; r.j = sel2[t]; }
  else { r.i = t; r.j = nn[t]; r.on = (dist[t] <= th) && r.j >= 0; }
  if (r.on) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { r.p1[a] = icp_f32(pts1[3 * r.i + a]); r.p2[a] = icp_f32(pts2[3 * r.j + a]); }
  }
  return r;
}

// out[0] = sum of the 1->2 terms, out[1] = sum of the 2->1 terms (point-to-plane) / unused, out[2] = count,
// out[3] = sum of the correspondence distances (the reference's "inlier error").  Deterministic: per-block
// partials, last block adds them in index order.
template <typename T, typename TN>
__global__ void __launch_bounds__(ICP_THREADS)
icp_forward_kernel(const T* __restrict__ pts1, const T* __restrict__ pts2, const TN* __restrict__ nrm1, const TN* __restrict__ nrm2,
                   const int64_t* __restrict__ nn, const double* __restrict__ dist, double th, const int64_t* __restrict__ sel1,
                   const int64_t* __restrict__ sel2, int64_t m, int point_to_plane, double* __restrict__ out,
                   double* __restrict__ partials, unsigned int* __restrict__ done) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  if (t < m) {
    const icp_pair c = icp_load(pts1, pts2, nn, dist, th, sel1, sel2, t);
    if (c.on) {
      const double dx = c.p2[0] - c.p1[0], dy = c.p2[1] - c.p1[1], dz = c.p2[2] - c.p1[2];
      v[2] = 1.0;
      v[3] = dist ? dist[t] : 0.0;
      if (point_to_plane) {
        const double ax = (double)nrm1[3 * c.i], ay = (double)nrm1[3 * c.i + 1], az = (double)nrm1[3 * c.i + 2];
        const double bx = (double)nrm2[3 * c.j], by = (double)nrm2[3 * c.j + 1], bz = (double)nrm2[3 * c.j + 2];
        // |n (n . d)| = |n . d| |n|   (loss.py:452-463)
        v[0] = fabs(ax * dx + ay * dy + az * dz) * sqrt(ax * ax + ay * ay + az * az);
        v[1] = fabs(bx * dx + by * dy + bz * dz) * sqrt(bx * bx + by * by + bz * bz);
      } else {
        v[0] = sqrt(dx * dx + dy * dy + dz * dz);      // loss.py:546-547
      }
    }
  }
  typedef cub::BlockReduce<double, ICP_THREADS> BR;
  __shared__ typename BR::TempStorage tmp;
  __shared__ bool is_last;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double s = BR(tmp).Sum(v[q]);
    if (threadIdx.x == 0) partials[4 * (size_t)blockIdx.x + q] = s;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = (atomicAdd(done, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double a = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += ICP_THREADS) a += ((volatile double*)partials)[4 * (size_t)b + q];
    const double s = BR(tmp).Sum(a);
    if (threadIdx.x == 0) out[q] = s;
    __syncthreads();
  }
  if (threadIdx.x == 0) *done = 0;
}

// d loss / d {points1, points2, normals1, normals2} for loss = c12 * sum12 + c21 * sum21 (fp64 accumulators,
// reductions into the second cloud: several queries may share a nearest neighbour).
template <typename T, typename TN>
__global__ void __launch_bounds__(ICP_THREADS)
icp_backward_kernel(const T* __restrict__ pts1, const T* __restrict__ pts2, const TN* __restrict__ nrm1, const TN* __restrict__ nrm2,
                    const int64_t* __restrict__ nn, const double* __restrict__ dist, double th, const int64_t* __restrict__ sel1,
                    const int64_t* __restrict__ sel2, int64_t m, int point_to_plane, const double* __restrict__ coef,
                    double* __restrict__ g1, double* __restrict__ g2, double* __restrict__ gn1, double* __restrict__ gn2) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  const icp_pair c = icp_load(pts1, pts2, nn, dist, th, sel1, sel2, t);
  if (!c.on) return;
  const double c12 = coef[0], c21 = coef[1];
  const double d[3] = {c.p2[0] - c.p1[0], c.p2[1] - c.p1[1], c.p2[2] - c.p1[2]};
  double gd[3] = {0.0, 0.0, 0.0};                    // d loss / d (p2 - p1)
  if (point_to_plane) {
    const TN* ns[2] = {nrm1 + 3 * c.i, nrm2 + 3 * c.j};
    double* gns[2] = {gn1 ? gn1 + 3 * c.i : nullptr, gn2 ? gn2 + 3 * c.j : nullptr};
    const double cs[2] = {c12, c21};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const double n0 = (double)ns[s][0], n1 = (double)ns[s][1], n2 = (double)ns[s][2];
      const double k = n0 * d[0] + n1 * d[1] + n2 * d[2];
      const double nn_ = sqrt(n0 * n0 + n1 * n1 + n2 * n2);
      const double sg = k > 0.0 ? 1.0 : (k < 0.0 ? -1.0 : 0.0);       // torch: d|x|/dx = 0 at 0
      const double f = cs[s] * sg * nn_;
      gd[0] += f * n0; gd[1] += f * n1; gd[2] += f * n2;
      if (gns[s]) {
        const double h = nn_ > 0.0 ? cs[s] * fabs(k) / nn_ : 0.0;
        atomicAdd(gns[s] + 0, f * d[0] + h * n0);
        atomicAdd(gns[s] + 1, f * d[1] + h * n1);
        atomicAdd(gns[s] + 2, f * d[2] + h * n2);
      }
    }
  } else {
    const double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double f = len > 0.0 ? c12 / len : 0.0;
    gd[0] = f * d[0]; gd[1] = f * d[1]; gd[2] = f * d[2];
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (g1) atomicAdd(g1 + 3 * c.i + a, -gd[a]);
    if (g2) atomicAdd(g2 + 3 * c.j + a, gd[a]);
  }
}

template <typename T, typename TN>
static int icp_launch(const void* pts1, const void* pts2, const void* nrm1, const void* nrm2, const int64_t* nn, const double* dist,
                      double th, const int64_t* sel1, const int64_t* sel2, int64_t m, int p2pl, double* out, void* partials,
                      size_t partials_bytes, const double* coef, double* g1, double* g2, double* gn1, double* gn2, cudaStream_t st) {
  const int blocks = dc_blocks(m, ICP_THREADS);
  if (out) {
    if (partials_bytes < (size_t)blocks * 32 + 16) return dc_set_error(DC_ERR_ARG, "dc_icp_forward: partials buffer too small (need 32*blocks+16 bytes)");
    unsigned int* done = (unsigned int*)((char*)partials + (size_t)blocks * 32);
    icp_forward_kernel<T, TN><<<blocks, ICP_THREADS, 0, st>>>((const T*)pts1, (const T*)pts2, (const TN*)nrm1, (const TN*)nrm2, nn, dist, th,
                                                              sel1, sel2, m, p2pl, out, (double*)partials, done);
  } else {
    icp_backward_kernel<T, TN><<<blocks, ICP_THREADS, 0, st>>>((const T*)pts1, (const T*)pts2, (const TN*)nrm1, (const TN*)nrm2, nn, dist, th,
                                                               sel1, sel2, m, p2pl, coef, g1, g2, gn1, gn2);
  }
  DC_LAUNCH_CHECK();
  return DC_OK;
}

static int icp_dispatch(const void* pts1, const void* pts2, int dtype, const void* nrm1, const void* nrm2, int ndtype, const int64_t* nn,
                        const double* dist, double th, const int64_t* sel1, const int64_t* sel2, int64_t m, int p2pl, double* out,
                        void* partials, size_t partials_bytes, const double* coef, double* g1, double* g2, double* gn1, double* gn2,
                        void* stream) {
  if (m <= 0) return DC_OK;
  if (!sel1 && (!nn || !dist)) return dc_set_error(DC_ERR_ARG, "dc_icp: need either (sel1, sel2) or (nn, dist)");
  if (p2pl && (!nrm1 || !nrm2)) return dc_set_error(DC_ERR_ARG, "dc_icp: point-to-plane needs the normals of both clouds");
  cudaStream_t st = (cudaStream_t)stream;
#define ICP_GO(T, TN) return icp_launch<T, TN>(pts1, pts2, nrm1, nrm2, nn, dist, th, sel1, sel2, m, p2pl, out, partials, partials_bytes, coef, g1, g2, gn1, gn2, st)
  if (dtype == DC_F32 && ndtype == DC_F32) ICP_GO(float, float);
  if (dtype == DC_F32) ICP_GO(float, double);
  if (ndtype == DC_F32) ICP_GO(double, float);
  ICP_GO(double, double);
#undef ICP_GO
}

extern "C" int dc_icp_forward(const void* points1, const void* points2, int dtype, const void* normals1, const void* normals2,
                              int normals_dtype, const int64_t* nn, const double* dist, double threshold, const int64_t* sel1,
                              const int64_t* sel2, int64_t m, int point_to_plane, double* out4, void* partials, size_t partials_bytes,
                              void* stream) {
  if (m <= 0) {
    if (out4) DC_CUDA_CHECK(cudaMemsetAsync(out4, 0, 4 * sizeof(double), (cudaStream_t)stream));
    return DC_OK;
  }
  return icp_dispatch(points1, points2, dtype, normals1, normals2, normals_dtype, nn, dist, threshold, sel1, sel2, m, point_to_plane, out4,
                      partials, partials_bytes, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int dc_icp_backward(const void* points1, const void* points2, int dtype, const void* normals1, const void* normals2,
                               int normals_dtype, const int64_t* nn, const double* dist, double threshold, const int64_t* sel1,
                               const int64_t* sel2, int64_t m, int point_to_plane, const double* coef2, double* g_points1,
                               double* g_points2, double* g_normals1, double* g_normals2, void* stream) {
  return icp_dispatch(points1, points2, dtype, normals1, normals2, normals_dtype, nn, dist, threshold, sel1, sel2, m, point_to_plane, nullptr,
                      nullptr, 0, coef2, g_points1, g_points2, g_normals1, g_normals2, stream);
}

// ---------------------------------------------------------------------------------------------
// Order-preserving 64-bit keys of non-negative fp64 values (NaN -> all ones, sorts last; *n_nan counts them):
// the inlier threshold torch.nanquantile(dists, ratio) (loss.py:437, 524) is read from the radix-sorted keys.
// ---------------------------------------------------------------------------------------------
__global__ void f64_keys_kernel(const double* __restrict__ x, int64_t n, uint64_t* __restrict__ keys, int32_t* __restrict__ n_nan) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  if (v != v) { keys[i] = ~0ull; atomicAdd(n_nan, 1); return; }
  const uint64_t b = (uint64_t)__double_as_longlong(v);
  keys[i] = (b >> 63) ? ~b : (b | (1ull << 63));      // total order of IEEE doubles
}

__global__ void f64_unkeys_kernel(const uint64_t* __restrict__ keys, int64_t n, double* __restrict__ x) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  const uint64_t b = (k >> 63) ? (k & ~(1ull << 63)) : ~k;
  x[i] = __longlong_as_double((long long)b);
}

extern "C" int dc_f64_sort_keys(const double* x, int64_t n, uint64_t* keys, int32_t* n_nan, void* stream) {
  if (n <= 0) return DC_OK;
  f64_keys_kernel<<<dc_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, keys, n_nan);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_f64_from_sort_keys(const uint64_t* keys, int64_t n, double* x, void* stream) {
  if (n <= 0) return DC_OK;
  f64_unkeys_kernel<<<dc_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, n, x);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
