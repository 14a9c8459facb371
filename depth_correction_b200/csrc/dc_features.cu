// Unfused per-point feature kernels on caller-ordered tensors: the native work behind
// DepthCloud.update_mean / update_cov / update_eig / update_normals / update_incidence_angles
// (depth_cloud.py:291-424) when they are called one by one, plus their backward passes so the
// staged API stays differentiable like the reference's autograd graph.  The training loop does not
// use these; it runs the fused kernels in dc_step.cu.
#include <string.h>
#include "dc_common.cuh"
#include "dc_math.cuh"

#define FEAT_THREADS 128

template <typename T>
__device__ __forceinline__ void feat_accumulate(const T* __restrict__ pts, int64_t n, const int64_t* __restrict__ nb,
                                                const float* __restrict__ wt, int K, const double pi[3], double& W,
                                                double S1[3], double S2[6]) {
  W = 0.0;
  S1[0] = S1[1] = S1[2] = 0.0;
  for (int k = 0; k < 6; ++k) S2[k] = 0.0;
  for (int c = 0; c < K; ++c) {
    int64_t j = nb[c];
    const double w = wt ? (double)wt[c] : (j >= 0 ? 1.0 : 0.0);
    if (j < 0) j += n;                 // python-style wrap of the -1 padding (depth_cloud.py:304)
    if (j < 0 || j >= n) continue;
    if (w == 0.0) continue;            // 0 * finite == 0 in the reference as well
    const double dx = (double)pts[3 * j] - pi[0], dy = (double)pts[3 * j + 1] - pi[1], dz = (double)pts[3 * j + 2] - pi[2];
    W += w;
    S1[0] += w * dx; S1[1] += w * dy; S1[2] += w * dz;
    S2[0] += w * dx * dx; S2[1] += w * dx * dy; S2[2] += w * dx * dz;
    S2[3] += w * dy * dy; S2[4] += w * dy * dz; S2[5] += w * dz * dz;
  }
}

template <typename T>
__global__ void __launch_bounds__(FEAT_THREADS)
features_kernel(const T* __restrict__ pts, int64_t n, const int64_t* __restrict__ neighbors,
                const float* __restrict__ weights, int K, T* __restrict__ mean, T* __restrict__ cov) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double pi[3] = {(double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2]};
  double W, S1[3], S2[6];
  feat_accumulate<T>(pts, n, neighbors + i * K, weights ? weights + i * K : nullptr, K, pi, W, S1, S2);
  const double iw = 1.0 / W;
  if (mean) {
    mean[3 * i] = (T)(pi[0] + S1[0] * iw); mean[3 * i + 1] = (T)(pi[1] + S1[1] * iw); mean[3 * i + 2] = (T)(pi[2] + S1[2] * iw);
  }
  if (cov) {
    const double icw = 1.0 / fmax(W - 1.0, 1e-6);
    const double xx = (S2[0] - S1[0] * S1[0] * iw) * icw, xy = (S2[1] - S1[0] * S1[1] * iw) * icw,
                 xz = (S2[2] - S1[0] * S1[2] * iw) * icw, yy = (S2[3] - S1[1] * S1[1] * iw) * icw,
                 yz = (S2[4] - S1[1] * S1[2] * iw) * icw, zz = (S2[5] - S1[2] * S1[2] * iw) * icw;
    T* c = cov + 9 * i;
    c[0] = (T)xx; c[1] = (T)xy; c[2] = (T)xz; c[3] = (T)xy; c[4] = (T)yy; c[5] = (T)yz; c[6] = (T)xz; c[7] = (T)yz; c[8] = (T)zz;
  }
}

extern "C" int dc_features(const void* points, int dtype, int64_t n, const int64_t* neighbors, const float* weights,
                           int K, void* mean, void* cov, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(n, FEAT_THREADS);
  if (dtype == DC_F32)
    features_kernel<float><<<blocks, FEAT_THREADS, 0, st>>>((const float*)points, n, neighbors, weights, K, (float*)mean, (float*)cov);
  else
    features_kernel<double><<<blocks, FEAT_THREADS, 0, st>>>((const double*)points, n, neighbors, weights, K, (double*)mean, (double*)cov);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// d mean_i / d p_j = w_ij / W_i ;  d cov_i / d p_j : (w_ij / cw_i) (G + G^T)(p_j - m_i)   (weights constant)
template <typename T>
__global__ void __launch_bounds__(FEAT_THREADS)
features_bwd_kernel(const T* __restrict__ pts, int64_t n, const int64_t* __restrict__ neighbors,
                    const float* __restrict__ weights, int K, const T* __restrict__ gmean, const T* __restrict__ gcov,
                    T* __restrict__ gpts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double pi[3] = {(double)pts[3 * i], (double)pts[3 * i + 1], (double)pts[3 * i + 2]};
  const int64_t* nb = neighbors + i * K;
  const float* wt = weights ? weights + i * K : nullptr;
  double W, S1[3], S2[6];
  feat_accumulate<T>(pts, n, nb, wt, K, pi, W, S1, S2);
  const double iw = 1.0 / W, icw = 1.0 / fmax(W - 1.0, 1e-6);
  const double mo[3] = {S1[0] * iw, S1[1] * iw, S1[2] * iw};   // mean - p_i
  double gm[3] = {0, 0, 0}, H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (gmean) { gm[0] = (double)gmean[3 * i] * iw; gm[1] = (double)gmean[3 * i + 1] * iw; gm[2] = (double)gmean[3 * i + 2] * iw; }
  if (gcov) {
    const T* g = gcov + 9 * i;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) H[3 * a + b] = ((double)g[3 * a + b] + (double)g[3 * b + a]) * icw;
  }
  for (int c = 0; c < K; ++c) {
    int64_t j = nb[c];
    const double w = wt ? (double)wt[c] : (j >= 0 ? 1.0 : 0.0);
    if (j < 0) j += n;
    if (j < 0 || j >= n || w == 0.0) continue;
    const double ex = (double)pts[3 * j] - pi[0] - mo[0], ey = (double)pts[3 * j + 1] - pi[1] - mo[1],
                 ez = (double)pts[3 * j + 2] - pi[2] - mo[2];
    const double fx = w * (gm[0] + H[0] * ex + H[1] * ey + H[2] * ez);
    const double fy = w * (gm[1] + H[3] * ex + H[4] * ey + H[5] * ez);
    const double fz = w * (gm[2] + H[6] * ex + H[7] * ey + H[8] * ez);
    atomicAdd(gpts + 3 * j, (T)fx);
    atomicAdd(gpts + 3 * j + 1, (T)fy);
    atomicAdd(gpts + 3 * j + 2, (T)fz);
  }
}

extern "C" int dc_features_backward(const void* points, int dtype, int64_t n, const int64_t* neighbors,
                                    const float* weights, int K, const void* gmean, const void* gcov, void* gpoints,
                                    void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(n, FEAT_THREADS);
  if (dtype == DC_F32)
    features_bwd_kernel<float><<<blocks, FEAT_THREADS, 0, st>>>((const float*)points, n, neighbors, weights, K,
                                                                (const float*)gmean, (const float*)gcov, (float*)gpoints);
  else
    features_bwd_kernel<double><<<blocks, FEAT_THREADS, 0, st>>>((const double*)points, n, neighbors, weights, K,
                                                                 (const double*)gmean, (const double*)gcov, (double*)gpoints);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Batched symmetric 3x3 eigen-decomposition (replaces torch.linalg.eigh on the host,
// depth_cloud.py:383-386).  eigvecs[i][a][j] = component a of the eigenvector of eigvals[i][j].
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void eigh3_kernel(const T* __restrict__ cov, int64_t n, T* __restrict__ eigvals, T* __restrict__ eigvecs) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T* c = cov + 9 * i;
  // LAPACK's default (UPLO='L') reads the lower triangle only; so do we
  dc_sym3 C = {(double)c[0], (double)c[3], (double)c[6], (double)c[4], (double)c[7], (double)c[8]};
  double lam[3], V[9];
  dc_sym3_eig(C, lam, V, eigvecs ? 3 : 0);
  if (eigvals) { eigvals[3 * i] = (T)lam[0]; eigvals[3 * i + 1] = (T)lam[1]; eigvals[3 * i + 2] = (T)lam[2]; }
  if (eigvecs) {
    T* o = eigvecs + 9 * i;
    for (int j = 0; j < 3; ++j)
      for (int a = 0; a < 3; ++a) o[3 * a + j] = (T)V[3 * j + a];
  }
}

extern "C" int dc_eigh3(const void* cov, int dtype, int64_t n, void* eigvals, void* eigvecs, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32) eigh3_kernel<float><<<dc_blocks(n, 128), 128, 0, st>>>((const float*)cov, n, (float*)eigvals, (float*)eigvecs);
  else eigh3_kernel<double><<<dc_blocks(n, 128), 128, 0, st>>>((const double*)cov, n, (double*)eigvals, (double*)eigvecs);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// gA = V (diag(gL) + skew(V^T gV) / E) V^T, E_ab = L_b - L_a   (the symmetric-eigh adjoint)
template <typename T>
__global__ void eigh3_bwd_kernel(const T* __restrict__ eigvals, const T* __restrict__ eigvecs, int64_t n,
                                 const T* __restrict__ gl, const T* __restrict__ gv, T* __restrict__ gcov) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double L[3], V[9], M[9];
  for (int k = 0; k < 3; ++k) L[k] = (double)eigvals[3 * i + k];
  for (int k = 0; k < 9; ++k) V[k] = (double)eigvecs[9 * i + k];   // V[3*a + j]
  for (int k = 0; k < 9; ++k) M[k] = 0.0;
  if (gv) {
    double GV[9], X[9];
    for (int k = 0; k < 9; ++k) GV[k] = (double)gv[9 * i + k];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) X[3 * a + b] = V[a] * GV[b] + V[3 + a] * GV[3 + b] + V[6 + a] * GV[6 + b];   // V^T gV
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        if (a != b) {
          const double num = 0.5 * (X[3 * a + b] - X[3 * b + a]);
          M[3 * a + b] = num != 0.0 ? num / (L[b] - L[a]) : 0.0;   // no 0/0 for repeated eigenvalues without eigenvector gradient
        }
  }
  if (gl) for (int a = 0; a < 3; ++a) M[4 * a] = (double)gl[3 * i + a];
  // gA = V M V^T
  double VM[9];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) VM[3 * a + b] = V[3 * a] * M[b] + V[3 * a + 1] * M[3 + b] + V[3 * a + 2] * M[6 + b];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b)
      gcov[9 * i + 3 * a + b] = (T)(VM[3 * a] * V[3 * b] + VM[3 * a + 1] * V[3 * b + 1] + VM[3 * a + 2] * V[3 * b + 2]);
}

extern "C" int dc_eigh3_backward(const void* eigvals, const void* eigvecs, int dtype, int64_t n, const void* geigvals,
                                 const void* geigvecs, void* gcov, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    eigh3_bwd_kernel<float><<<dc_blocks(n, 128), 128, 0, st>>>((const float*)eigvals, (const float*)eigvecs, n,
                                                               (const float*)geigvals, (const float*)geigvecs, (float*)gcov);
  else
    eigh3_bwd_kernel<double><<<dc_blocks(n, 128), 128, 0, st>>>((const double*)eigvals, (const double*)eigvecs, n,
                                                                (const double*)geigvals, (const double*)geigvecs, (double*)gcov);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// normals = -sign(dirs . v0) v0 (depth_cloud.py:401-415); inc = arccos(|dirs . n|) or arccos(-dirs . n) (:417-424)
template <typename T>
__global__ void normals_angles_kernel(const T* __restrict__ dirs, const T* __restrict__ eigvecs, int64_t n,
                                      int use_normal_sign, T* __restrict__ normals, T* __restrict__ inc) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d[3] = {(double)dirs[3 * i], (double)dirs[3 * i + 1], (double)dirs[3 * i + 2]};
  double v[3] = {(double)eigvecs[9 * i], (double)eigvecs[9 * i + 3], (double)eigvecs[9 * i + 6]};
  const double c = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
  const double sgn = c > 0.0 ? 1.0 : (c < 0.0 ? -1.0 : (c == 0.0 ? 0.0 : c));   // torch.sign, NaN propagates
  v[0] = -sgn * v[0]; v[1] = -sgn * v[1]; v[2] = -sgn * v[2];
  if (normals) { normals[3 * i] = (T)v[0]; normals[3 * i + 1] = (T)v[1]; normals[3 * i + 2] = (T)v[2]; }
  if (inc) {
    double cn = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
    // The reference does not clamp (depth_cloud.py:417-424): a ray exactly along the normal gives |cos| = 1 + rounding
    // and arccos -> NaN, which then poisons the corrected depth of that point and the loss terms of all its neighbours
    // (5 of 57 M points on the street map, 151 NaN loss terms).  Rounding overshoot of unit vectors stored in the
    // cloud's dtype (<= 1e-6) is clamped; anything farther out of range stays NaN like the reference.
    if (fabs(cn) > 1.0 && fabs(cn) < 1.0 + 1e-6) cn = cn > 0.0 ? 1.0 : -1.0;
    inc[i] = (T)acos(use_normal_sign ? -cn : fabs(cn));
  }
}

extern "C" int dc_normals_angles(const void* dirs, const void* eigvecs, int dtype, int64_t n, int use_normal_sign,
                                 void* normals, void* inc_angles, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    normals_angles_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)dirs, (const float*)eigvecs, n, use_normal_sign,
                                                                   (float*)normals, (float*)inc_angles);
  else
    normals_angles_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)dirs, (const double*)eigvecs, n, use_normal_sign,
                                                                    (double*)normals, (double*)inc_angles);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Points of one scan in the map frame, p = R (vp + depth * dir) + t, evaluated in fp64 from the stored
// records: cloud.transform(pose).to_points() of preproc.py:108-119 / depth_cloud.py:122-152 for the initial
// (uncorrected) global cloud that the neighbour search runs on.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void world_points_kernel(const T* __restrict__ vps, const T* __restrict__ dirs, const T* __restrict__ depth,
                                    int64_t n, const double* __restrict__ pose, double* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double d = (double)depth[i];
  double x = d * (double)dirs[3 * i], y = d * (double)dirs[3 * i + 1], z = d * (double)dirs[3 * i + 2];
  if (vps) { x += (double)vps[3 * i]; y += (double)vps[3 * i + 1]; z += (double)vps[3 * i + 2]; }
  out[3 * i] = pose[0] * x + pose[1] * y + pose[2] * z + pose[3];
  out[3 * i + 1] = pose[4] * x + pose[5] * y + pose[6] * z + pose[7];
  out[3 * i + 2] = pose[8] * x + pose[9] * y + pose[10] * z + pose[11];
}

extern "C" int dc_world_points(const void* vps, const void* dirs, const void* depth, int dtype, int64_t n,
                               const double* pose, double* out, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    world_points_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)vps, (const float*)dirs, (const float*)depth, n, pose, out);
  else
    world_points_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)vps, (const double*)dirs, (const double*)depth, n, pose, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// DepthCloud.from_points (depth_cloud.py:592-638): dirs = (x - vp) / |x - vp| where the depth is positive,
// depth = |x - vp|, in the cloud's own dtype.  One launch, no host synchronisation (the reference's boolean-mask
// assignment `dirs[valid] = ...` costs a nonzero() + device sync per scan).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void from_points_kernel(const T* __restrict__ pts, const T* __restrict__ vps, int64_t n, T* __restrict__ dirs,
                                   T* __restrict__ depth, T* __restrict__ vps_out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  T vx = 0, vy = 0, vz = 0;
  if (vps) { vx = vps[3 * i]; vy = vps[3 * i + 1]; vz = vps[3 * i + 2]; }
  T dx = pts[3 * i] - vx, dy = pts[3 * i + 1] - vy, dz = pts[3 * i + 2] - vz;
  const T d = sqrt(dx * dx + dy * dy + dz * dz);
  if (d > (T)0) { dx /= d; dy /= d; dz /= d; }      // NaN depth: left as is, like the reference's mask
  dirs[3 * i] = dx; dirs[3 * i + 1] = dy; dirs[3 * i + 2] = dz;
  depth[i] = d;
  if (vps_out) { vps_out[3 * i] = vx; vps_out[3 * i + 1] = vy; vps_out[3 * i + 2] = vz; }
}

extern "C" int dc_from_points(const void* points, const void* vps, int dtype, int64_t n, void* dirs, void* depth,
                              void* vps_out, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    from_points_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)points, (const float*)vps, n, (float*)dirs,
                                                                 (float*)depth, (float*)vps_out);
  else
    from_points_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)points, (const double*)vps, n, (double*)dirs,
                                                                  (double*)depth, (double*)vps_out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Feature masks (filters.py:85-113, 184-254; composed at preproc.py:53-62 and :130-142): every bound of a
// configuration in ONE launch instead of one compare + one AND per bound.  Comparisons are inclusive and done in the
// tensor's own dtype like torch (`x >= min` with a python scalar compares in x.dtype; the ratio is a division in
// x.dtype), NaN fails every bound.
// ---------------------------------------------------------------------------------------------
#define DC_MASK_MAX_BOUNDS 16
struct dc_mask_bounds {
  int n;
  int kind[DC_MASK_MAX_BOUNDS];      // 0: lo <= v[a] <= hi ; 1: lo <= v[a] / v[b] <= hi
  int a[DC_MASK_MAX_BOUNDS], b[DC_MASK_MAX_BOUNDS];
  int use_lo[DC_MASK_MAX_BOUNDS], use_hi[DC_MASK_MAX_BOUNDS];
  double lo[DC_MASK_MAX_BOUNDS], hi[DC_MASK_MAX_BOUNDS];
};

template <typename T>
__global__ void feature_mask_kernel(const T* __restrict__ vals, int64_t n, int stride, dc_mask_bounds bd,
                                    const int64_t* __restrict__ valid_counts, int64_t min_valid, int init,
                                    uint8_t* __restrict__ mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool keep = init ? true : (mask[i] != 0);
  if (valid_counts) keep = keep && valid_counts[i] >= min_valid;
  T v0 = (T)0, v1 = (T)0, v2 = (T)0;
  if (vals && bd.n > 0) {
    v0 = vals[i * stride];
    if (stride > 1) v1 = vals[i * stride + 1];
    if (stride > 2) v2 = vals[i * stride + 2];
  }
#pragma unroll 1
  for (int t = 0; t < bd.n; ++t) {
    const T va = bd.a[t] == 0 ? v0 : (bd.a[t] == 1 ? v1 : v2), vb = bd.b[t] == 0 ? v0 : (bd.b[t] == 1 ? v1 : v2);
    const T x = bd.kind[t] == 0 ? va : va / vb;
    if (bd.use_lo[t]) keep = keep && (x >= (T)bd.lo[t]);
    if (bd.use_hi[t]) keep = keep && (x <= (T)bd.hi[t]);
  }
  mask[i] = keep ? 1 : 0;
}

extern "C" int dc_feature_mask(const void* vals, int dtype, int64_t n, int stride, const double* bounds_host, int n_bounds,
                               const int64_t* valid_counts, int64_t min_valid, int init, uint8_t* mask, void* stream) {
  if (n <= 0) return DC_OK;
  if (n_bounds < 0 || n_bounds > DC_MASK_MAX_BOUNDS) return dc_set_error(DC_ERR_ARG, "dc_feature_mask: at most 16 bounds per call");
  if (stride < 1 || stride > 3) return dc_set_error(DC_ERR_ARG, "dc_feature_mask: stride must be 1..3");
  if (n_bounds > 0 && !vals) return dc_set_error(DC_ERR_ARG, "dc_feature_mask: bounds without values");
  dc_mask_bounds bd;
  memset(&bd, 0, sizeof(bd));
  bd.n = n_bounds;
  for (int t = 0; t < n_bounds; ++t) {
    const double* r = bounds_host + 5 * t;      // {kind, a, b, lo, hi}; lo = -inf / hi = +inf / NaN disable a side
    bd.kind[t] = (int)r[0];
    bd.a[t] = (int)r[1];
    bd.b[t] = (int)r[2];
    if (bd.kind[t] < 0 || bd.kind[t] > 1 || bd.a[t] < 0 || bd.a[t] >= stride || bd.b[t] < 0 || bd.b[t] >= stride)
      return dc_set_error(DC_ERR_ARG, "dc_feature_mask: bad bound record");
    bd.use_lo[t] = r[3] > -INFINITY;             // false for -inf and NaN (within_bounds: `min > -inf`)
    bd.use_hi[t] = r[4] < INFINITY;
    bd.lo[t] = r[3];
    bd.hi[t] = r[4];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    feature_mask_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>((const float*)vals, n, stride, bd, valid_counts, min_valid, init, mask);
  else
    feature_mask_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>((const double*)vals, n, stride, bd, valid_counts, min_valid, init, mask);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Per-point features of MANY clouds at once (local_feature_cloud of every scan, preproc.py:35-64): the neighbourhood
// pass is dc_step_forward on the stacked graph (sorted space, fp64); this kernel turns its outputs into the
// reference's per-point fields in the caller's order and dtype: eigvals, mean, oriented normals
// (depth_cloud.py:401-415) and incidence angles (:417-424).  stash: [n,8] = {mean xyz, v0 xyz, ., .} per sorted row.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void local_finish_kernel(const double* __restrict__ stash, const double* __restrict__ eig_sorted,
                                    const int32_t* __restrict__ order, const T* __restrict__ dirs, int64_t n, int use_normal_sign,
                                    T* __restrict__ eigvals, T* __restrict__ mean, T* __restrict__ normals, T* __restrict__ inc) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t o = order[s];
  const double* st = stash + 8 * s;
  if (eigvals) { eigvals[3 * o] = (T)eig_sorted[3 * s]; eigvals[3 * o + 1] = (T)eig_sorted[3 * s + 1]; eigvals[3 * o + 2] = (T)eig_sorted[3 * s + 2]; }
  if (mean) { mean[3 * o] = (T)st[0]; mean[3 * o + 1] = (T)st[1]; mean[3 * o + 2] = (T)st[2]; }
  // the normal goes through the cloud's dtype like the staged path (eigvecs are stored in T there)
  const double d[3] = {(double)dirs[3 * o], (double)dirs[3 * o + 1], (double)dirs[3 * o + 2]};
  double v[3] = {(double)(T)st[3], (double)(T)st[4], (double)(T)st[5]};
  const double c = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
  const double sgn = c > 0.0 ? 1.0 : (c < 0.0 ? -1.0 : (c == 0.0 ? 0.0 : c));
  v[0] = -sgn * v[0]; v[1] = -sgn * v[1]; v[2] = -sgn * v[2];
  if (normals) { normals[3 * o] = (T)v[0]; normals[3 * o + 1] = (T)v[1]; normals[3 * o + 2] = (T)v[2]; }
  if (inc) {
    double cn = d[0] * v[0] + d[1] * v[1] + d[2] * v[2];
    if (fabs(cn) > 1.0 && fabs(cn) < 1.0 + 1e-6) cn = cn > 0.0 ? 1.0 : -1.0;      // rounding overshoot (see dc_normals_angles)
    inc[o] = (T)acos(use_normal_sign ? -cn : fabs(cn));
  }
}

extern "C" int dc_local_features_finish(const double* stash, const double* eigvals_sorted, const int32_t* order, const void* dirs,
                                        int dtype, int64_t n, int use_normal_sign, void* eigvals, void* mean, void* normals,
                                        void* inc_angles, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DC_F32)
    local_finish_kernel<float><<<dc_blocks(n, 256), 256, 0, st>>>(stash, eigvals_sorted, order, (const float*)dirs, n, use_normal_sign,
                                                                 (float*)eigvals, (float*)mean, (float*)normals, (float*)inc_angles);
  else
    local_finish_kernel<double><<<dc_blocks(n, 256), 256, 0, st>>>(stash, eigvals_sorted, order, (const double*)dirs, n, use_normal_sign,
                                                                  (double*)eigvals, (double*)mean, (double*)normals, (double*)inc_angles);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
