// Kernels 2 and 3: the fused fixed-graph map-consistency step (forward A: corrected world points,
// forward B: neighbourhood covariance / eigen / loss, backward C: gather-form gradient down to the
// model weights, exponents and per-scan poses).  See include/dc_b200.h for the reference mapping.
//
// HBM layout (sorted space): packed scan records 2 x vec4 + u32 per point, one 32-byte fp64 point
// record per point (written by A, gathered by B and C), one 64-byte fp64 stash per point (written
// by B, gathered by C), sliced-ELL int32 neighbour indices (read once per pass, coalesced).
#include <cub/cub.cuh>
#include "dc_common.cuh"
#include "dc_math.cuh"

#define STEP_THREADS 128

template <typename T> struct vec4_of;
template <> struct vec4_of<float> { typedef float4 type; };
template <> struct vec4_of<double> { typedef double4 type; };

struct dc_model {
  int kind, n_terms;
  const double* w;
  const double* e;
};

// depth correction d' and (optionally) the powers g^e_k
__device__ __forceinline__ double dc_correct_depth(const dc_model& m, bool masked, double d, double g, double* pw) {
  if (m.kind == DC_MODEL_NONE || !masked) return d;
  double bias = 0.0;
#pragma unroll 1
  for (int k = 0; k < m.n_terms; ++k) {
    const double p = dc_pow_exp(g, m.e[k]);
    if (pw) pw[k] = p;
    bias += m.w[k] * p;
  }
  return m.kind == DC_MODEL_SCALED_POLYNOMIAL ? d * (1.0 - bias) : d - bias;
}

// ---------------------------------------------------------------------------------------------
// pack scan rows into sorted space (one-time, per scan)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_records_kernel(const T* __restrict__ vps, const T* __restrict__ dirs, const T* __restrict__ depth,
                                    const T* __restrict__ inc, const uint8_t* __restrict__ model_mask,
                                    const uint8_t* __restrict__ loss_mask, int64_t first, int64_t count, int scan_id,
                                    const int32_t* __restrict__ inv_order, typename vec4_of<T>::type* __restrict__ rec_dir,
                                    typename vec4_of<T>::type* __restrict__ rec_vp, uint32_t* __restrict__ rec_meta) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int64_t s = inv_order ? (int64_t)inv_order[first + i] : first + i;   // NULL: keep the original order
  typename vec4_of<T>::type a, b;
  a.x = dirs[3 * i]; a.y = dirs[3 * i + 1]; a.z = dirs[3 * i + 2]; a.w = depth[i];
  if (vps) { b.x = vps[3 * i]; b.y = vps[3 * i + 1]; b.z = vps[3 * i + 2]; }
  else { b.x = 0; b.y = 0; b.z = 0; }
  b.w = inc ? inc[i] : (T)0;
  uint32_t f = 0;
  if (!model_mask || model_mask[i]) f |= DC_PT_MODEL_MASK;
  if (!loss_mask || loss_mask[first + i]) f |= DC_PT_LOSS_MASK;
  rec_dir[s] = a;
  rec_vp[s] = b;
  rec_meta[s] = ((uint32_t)scan_id << 2) | f;
}

extern "C" int dc_pack_records(const void* vps, const void* dirs, const void* depth, const void* inc_angles,
                               const uint8_t* model_mask, const uint8_t* loss_mask, int dtype, int64_t first,
                               int64_t count, int scan_id, const int32_t* inv_order, void* rec_dir, void* rec_vp,
                               uint32_t* rec_meta, void* stream) {
  if (count <= 0) return DC_OK;
  if (scan_id < 0 || scan_id >= (1 << 30)) return dc_set_error(DC_ERR_ARG, "dc_pack_records: scan id out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(count, 256);
  if (dtype == DC_F32)
    pack_records_kernel<float><<<blocks, 256, 0, st>>>((const float*)vps, (const float*)dirs, (const float*)depth,
                                                       (const float*)inc_angles, model_mask, loss_mask, first, count,
                                                       scan_id, inv_order, (float4*)rec_dir, (float4*)rec_vp, rec_meta);
  else
    pack_records_kernel<double><<<blocks, 256, 0, st>>>((const double*)vps, (const double*)dirs, (const double*)depth,
                                                        (const double*)inc_angles, model_mask, loss_mask, first, count,
                                                        scan_id, inv_order, (double4*)rec_dir, (double4*)rec_vp, rec_meta);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// All scans in one launch: tbl[s] = {vps, dirs, depth, inc_angles, model_mask} device addresses of scan s (0 = absent),
// first[s] = global row of its first point.  Writes the sorted-space copy (through inv_order) and, optionally, the
// original-order copy used by the chain stage of the backward pass.
struct dc_scan_ptrs {
  unsigned long long vps, dirs, depth, inc, mask;
};

__device__ __forceinline__ int dc_find_scan(const int64_t* __restrict__ first, int n_scans, int64_t i) {
  int lo = 0, hi = n_scans;      // largest s with first[s] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(first + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void pack_records_batched_kernel(const dc_scan_ptrs* __restrict__ tbl, const int64_t* __restrict__ first,
                                            int n_scans, int64_t n, const int32_t* __restrict__ inv_order,
                                            typename vec4_of<T>::type* __restrict__ rec_dir,
                                            typename vec4_of<T>::type* __restrict__ rec_vp, uint32_t* __restrict__ rec_meta,
                                            typename vec4_of<T>::type* __restrict__ rec_dir_o,
                                            typename vec4_of<T>::type* __restrict__ rec_vp_o, uint32_t* __restrict__ rec_meta_o) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int s = dc_find_scan(first, n_scans, g);
  const int64_t i = g - first[s];
  const dc_scan_ptrs p = tbl[s];
  const T* vps = reinterpret_cast<const T*>(p.vps);
  const T* dirs = reinterpret_cast<const T*>(p.dirs);
  const T* depth = reinterpret_cast<const T*>(p.depth);
  const T* inc = reinterpret_cast<const T*>(p.inc);
  const uint8_t* mask = reinterpret_cast<const uint8_t*>(p.mask);
  typename vec4_of<T>::type a, b;
  a.x = dirs[3 * i]; a.y = dirs[3 * i + 1]; a.z = dirs[3 * i + 2]; a.w = depth[i];
  if (vps) { b.x = vps[3 * i]; b.y = vps[3 * i + 1]; b.z = vps[3 * i + 2]; }
  else { b.x = 0; b.y = 0; b.z = 0; }
  b.w = inc ? inc[i] : (T)0;
  const uint32_t meta = ((uint32_t)s << 2) | ((!mask || mask[i]) ? DC_PT_MODEL_MASK : 0u) | DC_PT_LOSS_MASK;
  const int64_t d = inv_order[g];
  rec_dir[d] = a; rec_vp[d] = b; rec_meta[d] = meta;
  if (rec_dir_o) { rec_dir_o[g] = a; rec_vp_o[g] = b; rec_meta_o[g] = meta; }
}

extern "C" int dc_pack_records_batched(const void* scan_ptr_table, const int64_t* first, int n_scans, int64_t n, int dtype,
                                       const int32_t* inv_order, void* rec_dir, void* rec_vp, uint32_t* rec_meta,
                                       void* rec_dir_o, void* rec_vp_o, uint32_t* rec_meta_o, void* stream) {
  if (n <= 0) return DC_OK;
  if (n_scans < 1 || n_scans >= (1 << 30)) return dc_set_error(DC_ERR_ARG, "dc_pack_records_batched: bad scan count");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(n, 256);
  const dc_scan_ptrs* tbl = (const dc_scan_ptrs*)scan_ptr_table;
  if (dtype == DC_F32)
    pack_records_batched_kernel<float><<<blocks, 256, 0, st>>>(tbl, first, n_scans, n, inv_order, (float4*)rec_dir, (float4*)rec_vp,
                                                               rec_meta, (float4*)rec_dir_o, (float4*)rec_vp_o, rec_meta_o);
  else
    pack_records_batched_kernel<double><<<blocks, 256, 0, st>>>(tbl, first, n_scans, n, inv_order, (double4*)rec_dir, (double4*)rec_vp,
                                                                rec_meta, (double4*)rec_dir_o, (double4*)rec_vp_o, rec_meta_o);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// Map-frame points of ALL scans in one launch (batched form of dc_world_points): out[g] = R_s (vp + depth dir) + t_s
// in fp64, poses = fp64 [S,16] row-major 4x4.
template <typename T>
__global__ void world_points_batched_kernel(const dc_scan_ptrs* __restrict__ tbl, const int64_t* __restrict__ first,
                                            int n_scans, int64_t n, const double* __restrict__ poses, double* __restrict__ out) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int s = dc_find_scan(first, n_scans, g);
  const int64_t i = g - first[s];
  const dc_scan_ptrs p = tbl[s];
  const T* vps = reinterpret_cast<const T*>(p.vps);
  const T* dirs = reinterpret_cast<const T*>(p.dirs);
  const T* depth = reinterpret_cast<const T*>(p.depth);
  const double* Tm = poses + 16 * (size_t)s;
  const double d = (double)depth[i];
  double x = d * (double)dirs[3 * i], y = d * (double)dirs[3 * i + 1], z = d * (double)dirs[3 * i + 2];
  if (vps) { x += (double)vps[3 * i]; y += (double)vps[3 * i + 1]; z += (double)vps[3 * i + 2]; }
  out[3 * g] = Tm[0] * x + Tm[1] * y + Tm[2] * z + Tm[3];
  out[3 * g + 1] = Tm[4] * x + Tm[5] * y + Tm[6] * z + Tm[7];
  out[3 * g + 2] = Tm[8] * x + Tm[9] * y + Tm[10] * z + Tm[11];
}

extern "C" int dc_world_points_batched(const void* scan_ptr_table, const int64_t* first, int n_scans, int64_t n, int dtype,
                                       const double* poses, double* out, void* stream) {
  if (n <= 0) return DC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(n, 256);
  const dc_scan_ptrs* tbl = (const dc_scan_ptrs*)scan_ptr_table;
  if (dtype == DC_F32) world_points_batched_kernel<float><<<blocks, 256, 0, st>>>(tbl, first, n_scans, n, poses, out);
  else world_points_batched_kernel<double><<<blocks, 256, 0, st>>>(tbl, first, n_scans, n, poses, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

__global__ void set_loss_mask_kernel(const uint8_t* __restrict__ loss_mask, int64_t n, const int32_t* __restrict__ order,
                                     uint32_t* __restrict__ rec_meta) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t m = rec_meta[s] & ~DC_PT_LOSS_MASK;
  if (!loss_mask || loss_mask[order[s]]) m |= DC_PT_LOSS_MASK;
  rec_meta[s] = m;
}

extern "C" int dc_set_loss_mask(const uint8_t* loss_mask, int64_t n, const int32_t* order, uint32_t* rec_meta, void* stream) {
  if (n <= 0) return DC_OK;
  set_loss_mask_kernel<<<dc_blocks(n, 256), 256, 0, (cudaStream_t)stream>>>(loss_mask, n, order, rec_meta);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Pass A: p = R_s (vp + d' dir) + t_s     (model.py:250-261, depth_cloud.py:122-152)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
step_points_kernel(const typename vec4_of<T>::type* __restrict__ rec_dir, const typename vec4_of<T>::type* __restrict__ rec_vp,
                   const uint32_t* __restrict__ rec_meta, int64_t n, const double* __restrict__ poses, dc_model model,
                   dc_point* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const typename vec4_of<T>::type a = rec_dir[i], b = rec_vp[i];
  const uint32_t meta = rec_meta[i];
  const double* Tm = poses + 12 * (size_t)(meta >> 2);
  const double d = dc_correct_depth(model, meta & DC_PT_MODEL_MASK, (double)a.w, (double)b.w, nullptr);
  const double x = (double)b.x + d * (double)a.x, y = (double)b.y + d * (double)a.y, z = (double)b.z + d * (double)a.z;
  dc_point p;
  p.x = Tm[0] * x + Tm[1] * y + Tm[2] * z + Tm[3];
  p.y = Tm[4] * x + Tm[5] * y + Tm[6] * z + Tm[7];
  p.z = Tm[8] * x + Tm[9] * y + Tm[10] * z + Tm[11];
  p.tag = 0;
  out[i] = p;
}

extern "C" int dc_step_points(const void* rec_dir, const void* rec_vp, const uint32_t* rec_meta, int dtype, int64_t n,
                              const double* poses, int n_scans, int model_kind, const double* w, const double* exponent,
                              int n_terms, void* points_out, void* stream) {
  if (n <= 0) return DC_OK;
  if (n_terms < 0 || n_terms > DC_MAX_TERMS) return dc_set_error(DC_ERR_ARG, "dc_step_points: too many polynomial terms");
  (void)n_scans;
  dc_model m = {model_kind, n_terms, w, exponent};
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(n, 256);
  if (dtype == DC_F32)
    step_points_kernel<float><<<blocks, 256, 0, st>>>((const float4*)rec_dir, (const float4*)rec_vp, rec_meta, n, poses, m, (dc_point*)points_out);
  else
    step_points_kernel<double><<<blocks, 256, 0, st>>>((const double4*)rec_dir, (const double4*)rec_vp, rec_meta, n, poses, m, (dc_point*)points_out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Pass B: per point, gather the neighbourhood (coalesced index columns, 256-bit point records),
// accumulate first and second moments relative to the query point in fp64, then
// mean / covariance (utils.py:109-149) / eigen (depth_cloud.py:376-399) / loss (loss.py:216-370).
// ---------------------------------------------------------------------------------------------
struct dc_stash {   // 64 bytes
  double mx, my, mz;   // neighbourhood mean (world)
  double vx, vy, vz;   // eigenvector of the smallest eigenvalue
  double alpha, beta;  // dl/dp_j = (alpha v v^T + beta I)(p_j - m)
};

__device__ __forceinline__ void dc_store_stash(dc_stash* dst, const dc_stash& s) {
  double4* d4 = reinterpret_cast<double4*>(dst);
  d4[0] = make_double4(s.mx, s.my, s.mz, s.vx);
  d4[1] = make_double4(s.vy, s.vz, s.alpha, s.beta);
}

// SCAT: the backward scatter (pass C1, float32 form) runs in the same kernel.  For the mean / sum reductions the
// upstream gradient of every loss term is ONE scalar that the chain stage applies afterwards, so right after the
// eigen epilogue the thread still holds everything the scatter needs (mean, v0, alpha, beta) in registers: it walks its
// index column a second time (L1/L2-hot) and issues the vector reductions into g32.  No stash is written or re-read
// and the index array and the neighbour records are not fetched from HBM a second time.
template <int LOSS, bool SCAT>
__global__ void __launch_bounds__(STEP_THREADS)
step_forward_kernel(const dc_point* __restrict__ P, const uint32_t* __restrict__ rec_meta, int64_t n,
                    const int64_t* __restrict__ slice_ptr, const int32_t* __restrict__ ell_idx, int flags,
                    double* __restrict__ loss_pp, dc_stash* __restrict__ stash, double* __restrict__ eigvals,
                    double* __restrict__ loss_sum, double* __restrict__ partials, unsigned int* __restrict__ done,
                    float4* __restrict__ g32) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = row < n;
  double l_acc = 0.0, c_acc = 0.0;
  if (live) {
    const int64_t base = slice_ptr[row >> 5];
    const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
    const int32_t* col = ell_idx + base + lane;
    const dc_point pi = dc_ld_point(P + row);
    const int self = (int)row;
    int W = 0;
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    int c = 0;
    // 4 independent gathers in flight per thread; a missing neighbour (-1) is replaced by the query
    // itself, whose offset is exactly zero, so the accumulators need no predication.
    for (; c + 4 <= width; c += 4) {
      const int j0 = __ldg(col + (c + 0) * DC_SLICE), j1 = __ldg(col + (c + 1) * DC_SLICE);
      const int j2 = __ldg(col + (c + 2) * DC_SLICE), j3 = __ldg(col + (c + 3) * DC_SLICE);
      const dc_point p0 = dc_ld_point(P + (j0 >= 0 ? j0 : self));
      const dc_point p1 = dc_ld_point(P + (j1 >= 0 ? j1 : self));
      const dc_point p2 = dc_ld_point(P + (j2 >= 0 ? j2 : self));
      const dc_point p3 = dc_ld_point(P + (j3 >= 0 ? j3 : self));
      W += (j0 >= 0) + (j1 >= 0) + (j2 >= 0) + (j3 >= 0);
#define DC_ACC(p)                                                   \
  {                                                                 \
    const double dx = p.x - pi.x, dy = p.y - pi.y, dz = p.z - pi.z; \
    sx += dx; sy += dy; sz += dz;                                   \
    sxx = fma(dx, dx, sxx); sxy = fma(dx, dy, sxy); sxz = fma(dx, dz, sxz); \
    syy = fma(dy, dy, syy); syz = fma(dy, dz, syz); szz = fma(dz, dz, szz); \
  }
      DC_ACC(p0) DC_ACC(p1) DC_ACC(p2) DC_ACC(p3)
    }
    for (; c < width; ++c) {
      const int j0 = __ldg(col + c * DC_SLICE);
      const dc_point p0 = dc_ld_point(P + (j0 >= 0 ? j0 : self));
      W += (j0 >= 0);
      DC_ACC(p0)
    }
#undef DC_ACC
    const double Wd = (double)W;
    const double iw = 1.0 / Wd;                       // W == 0 -> inf -> NaN features, like the reference's 0/0
    const double cw = fmax(Wd - 1.0, 1e-6);           // utils.py:143-146
    const double icw = 1.0 / cw;
    dc_sym3 C;
    C.xx = (sxx - sx * sx * iw) * icw; C.xy = (sxy - sx * sy * iw) * icw; C.xz = (sxz - sx * sz * iw) * icw;
    C.yy = (syy - sy * sy * iw) * icw; C.yz = (syz - sy * sz * iw) * icw; C.zz = (szz - sz * sz * iw) * icw;
    const double tr = C.xx + C.yy + C.zz;
    double raw, alpha = 0.0, beta = 0.0;
    double v0[3] = {0.0, 0.0, 0.0};
    if (LOSS == DC_LOSS_TRACE) {
      raw = tr;                                        // loss.py:329-330
      beta = 1.0;
      if (eigvals) {
        double lam[3];
        dc_sym3_eig(C, lam, nullptr, 0);
        eigvals[3 * row] = lam[0]; eigvals[3 * row + 1] = lam[1]; eigvals[3 * row + 2] = lam[2];
      }
    } else {
      double lam[3];
      dc_sym3_eig(C, lam, v0, 1);
      // rank-deficient neighbourhoods (<= 3 points, exact planes): lambda0 is rounding noise of either sign;
      // treat it as zero so that sqrt'(noise) ~ 1e9 cannot leak noise into the gradients
      if (fabs(lam[0]) <= 1e-14 * fabs(lam[2])) lam[0] = 0.0;
      if (eigvals) { eigvals[3 * row] = lam[0]; eigvals[3 * row + 1] = lam[1]; eigvals[3 * row + 2] = lam[2]; }
      if (flags & DC_FLAG_NORMALIZATION) {             // loss.py:253-254
        const double tc = fmax(tr, 1e-6);
        raw = lam[0] / tc;
        alpha = 1.0 / tc;
        beta = tr > 1e-6 ? -lam[0] / (tc * tc) : 0.0;
      } else {
        raw = lam[0];
        alpha = 1.0;
      }
    }
    const bool masked = rec_meta[row] & DC_PT_LOSS_MASK;
    double val = raw, dfac = 1.0;
    if (!(flags & DC_FLAG_RAW)) {
      // relu (loss.py:284) then optional sqrt (:286-287); relu'(x <= 0) = 0 also kills sqrt'(0) = inf
      if (raw > 0.0) {
        if (flags & DC_FLAG_SQRT) { val = sqrt(raw); dfac = 0.5 / val; }
      } else if (raw <= 0.0) {
        val = 0.0; dfac = 0.0;
      }                                                // NaN falls through unchanged
      if (!masked) { val = 0.0; dfac = 0.0; }
    }
    if (loss_pp) loss_pp[row] = val;
    if (masked) { l_acc = val; c_acc = 1.0; }
    if (stash) {
      const double f = 2.0 * icw * dfac;
      dc_stash s;
      s.mx = pi.x + sx * iw; s.my = pi.y + sy * iw; s.mz = pi.z + sz * iw;
      s.vx = v0[0]; s.vy = v0[1]; s.vz = v0[2];
      s.alpha = f * alpha; s.beta = f * beta;
      dc_store_stash(stash + row, s);
    }
    if (SCAT) {
      const double f = 2.0 * icw * dfac;
      const double ua = f * alpha, ub = f * beta;
      if (ua != 0.0 || ub != 0.0) {        // masked-out / inactive loss term: nothing to scatter
        const double mx = pi.x + sx * iw, my = pi.y + sy * iw, mz = pi.z + sz * iw;
#define DC_SCAT(j_)                                                                  \
  if ((j_) >= 0) {                                                                   \
    const dc_point pj = dc_ld_point(P + (j_));                                       \
    const double ex = pj.x - mx, ey = pj.y - my, ez = pj.z - mz;                     \
    const double a = ua * (v0[0] * ex + v0[1] * ey + v0[2] * ez);                    \
    const double vx = a * v0[0] + ub * ex, vy = a * v0[1] + ub * ey, vz = a * v0[2] + ub * ez; \
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g32 + (j_)), "f"((float)vx), \
                 "f"((float)vy), "f"((float)vz), "f"(0.f) : "memory");               \
  }
        int c2 = 0;
        for (; c2 + 4 <= width; c2 += 4) {
          const int j0 = __ldg(col + (c2 + 0) * DC_SLICE), j1 = __ldg(col + (c2 + 1) * DC_SLICE);
          const int j2 = __ldg(col + (c2 + 2) * DC_SLICE), j3 = __ldg(col + (c2 + 3) * DC_SLICE);
          DC_SCAT(j0) DC_SCAT(j1) DC_SCAT(j2) DC_SCAT(j3)
        }
        for (; c2 < width; ++c2) {
          const int j0 = __ldg(col + c2 * DC_SLICE);
          DC_SCAT(j0)
        }
#undef DC_SCAT
      }
    }
  }
  if (!loss_sum) return;
  // deterministic two-level reduction: per-block partials, last block adds them in index order
  typedef cub::BlockReduce<double, STEP_THREADS> BR;
  __shared__ typename BR::TempStorage tmp;
  __shared__ bool is_last;
  const double bl = BR(tmp).Sum(l_acc);
  __syncthreads();
  const double bc = BR(tmp).Sum(c_acc);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = bl;
    partials[2 * blockIdx.x + 1] = bc;
    __threadfence();
    const unsigned int t = atomicAdd(done, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += STEP_THREADS) {
      a += ((volatile double*)partials)[2 * i];
      b += ((volatile double*)partials)[2 * i + 1];
    }
    __syncthreads();
    const double ta = BR(tmp).Sum(a);
    __syncthreads();
    const double tb = BR(tmp).Sum(b);
    if (threadIdx.x == 0) {
      loss_sum[0] = ta;
      loss_sum[1] = tb;
      *done = 0;   // re-arm for the next launch
    }
  }
}

extern "C" int dc_step_forward(const void* points, const uint32_t* rec_meta, int64_t n, const int64_t* slice_ptr,
                               const int32_t* ell_idx, int loss_kind, int flags, double* loss_pp, double* stash,
                               double* eigvals, double* loss_sum, void* partials, size_t partials_bytes, void* stream) {
  if (n <= 0) return DC_OK;
  const int blocks = dc_blocks(((n + 31) / 32) * 32, STEP_THREADS);
  if (loss_sum && partials_bytes < (size_t)blocks * 16 + 16)
    return dc_set_error(DC_ERR_ARG, "dc_step_forward: partials buffer too small (need 16*blocks+16 bytes)");
  double* part = (double*)partials;
  unsigned int* done = loss_sum ? (unsigned int*)((char*)partials + (size_t)blocks * 16) : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (loss_kind == DC_LOSS_TRACE)
    step_forward_kernel<DC_LOSS_TRACE, false><<<blocks, STEP_THREADS, 0, st>>>((const dc_point*)points, rec_meta, n, slice_ptr, ell_idx, flags,
                                                                                loss_pp, (dc_stash*)stash, eigvals, loss_sum, part, done, nullptr);
  else if (loss_kind == DC_LOSS_MIN_EIGVAL)
    step_forward_kernel<DC_LOSS_MIN_EIGVAL, false><<<blocks, STEP_THREADS, 0, st>>>((const dc_point*)points, rec_meta, n, slice_ptr, ell_idx, flags,
                                                                                     loss_pp, (dc_stash*)stash, eigvals, loss_sum, part, done, nullptr);
  else
    return dc_set_error(DC_ERR_ARG, "dc_step_forward: unknown loss kind");
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_step_forward_scatter(const void* points, const uint32_t* rec_meta, int64_t n, const int64_t* slice_ptr,
                                       const int32_t* ell_idx, int loss_kind, int flags, double* loss_pp, void* g_sorted32,
                                       double* loss_sum, void* partials, size_t partials_bytes, void* stream) {
  if (n <= 0) return DC_OK;
  if (flags & DC_FLAG_RAW) return dc_set_error(DC_ERR_ARG, "dc_step_forward_scatter: per-point upstream gradients need the two-kernel form");
  if (!g_sorted32 || !loss_sum) return dc_set_error(DC_ERR_ARG, "dc_step_forward_scatter: g_sorted32 and loss_sum are required");
  const int blocks = dc_blocks(((n + 31) / 32) * 32, STEP_THREADS);
  if (partials_bytes < (size_t)blocks * 16 + 16)
    return dc_set_error(DC_ERR_ARG, "dc_step_forward_scatter: partials buffer too small (need 16*blocks+16 bytes)");
  double* part = (double*)partials;
  unsigned int* done = (unsigned int*)((char*)partials + (size_t)blocks * 16);
  cudaStream_t st = (cudaStream_t)stream;
  DC_CUDA_CHECK(cudaMemsetAsync(g_sorted32, 0, (size_t)n * 16, st));
  if (loss_kind == DC_LOSS_TRACE)
    step_forward_kernel<DC_LOSS_TRACE, true><<<blocks, STEP_THREADS, 0, st>>>((const dc_point*)points, rec_meta, n, slice_ptr, ell_idx, flags,
                                                                               loss_pp, nullptr, nullptr, loss_sum, part, done, (float4*)g_sorted32);
  else if (loss_kind == DC_LOSS_MIN_EIGVAL)
    step_forward_kernel<DC_LOSS_MIN_EIGVAL, true><<<blocks, STEP_THREADS, 0, st>>>((const dc_point*)points, rec_meta, n, slice_ptr, ell_idx, flags,
                                                                                    loss_pp, nullptr, nullptr, loss_sum, part, done, (float4*)g_sorted32);
  else
    return dc_set_error(DC_ERR_ARG, "dc_step_forward_scatter: unknown loss kind");
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Pass C: backward in two atomic-free stages.
//
// C1 (sorted space, gather form).  Row j of the TRANSPOSED graph lists every i with j in N(i), so
//   g_j = sum_i u_i (alpha_i v_i v_i^T + beta_i I)(p_j - m_i)
// is a gather over 64-byte stash records; g_j is written to the point's ORIGINAL position (24 bytes).
//
// C2 (original order = scan-major, fully coalesced).  g_j is chained through p_j = R_s (vp + d' dir) + t_s
// to dL/dw_k, dL/de_k and the 3x4 pose gradient.  Blocks are aligned to scans (block table built once on
// the host), so all threads of a block belong to one scan and a plain block reduction yields one partial
// record per block -- no atomics, no per-scan scatter.
//
// C3.  Partial records are summed in a fixed order per scan (pose gradients) and over all blocks (model
// gradients): the gradients are bitwise reproducible.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double dc_warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(STEP_THREADS)
step_backward_gather_kernel(const dc_point* __restrict__ P, int64_t n, const int64_t* __restrict__ slice_ptr,
                            const int32_t* __restrict__ ell_idx, const dc_stash* __restrict__ stash,
                            const double* __restrict__ upstream, const int32_t* __restrict__ order,
                            double* __restrict__ g_out) {
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  const int32_t* col = ell_idx + base + lane;
  const dc_point pj = dc_ld_point(P + row);
  const int self = (int)row;
  double gx = 0, gy = 0, gz = 0;
#define DC_BWD(i_)                                                                       \
  {                                                                                      \
    const int ii = (i_) >= 0 ? (i_) : self;                                              \
    double4 s0, s1;                                                                      \
    dc_ld256(stash + ii, s0.x, s0.y, s0.z, s0.w);                                        \
    dc_ld256(reinterpret_cast<const char*>(stash + ii) + 32, s1.x, s1.y, s1.z, s1.w);    \
    double u = (i_) >= 0 ? 1.0 : 0.0;                                                    \
    if (upstream) u *= __ldg(upstream + ii);                                             \
    const double ex = pj.x - s0.x, ey = pj.y - s0.y, ez = pj.z - s0.z;                   \
    const double a = u * s1.z * (s0.w * ex + s1.x * ey + s1.y * ez), b = u * s1.w;       \
    gx += a * s0.w + b * ex; gy += a * s1.x + b * ey; gz += a * s1.y + b * ez;           \
  }
  int c = 0;
  for (; c + 4 <= width; c += 4) {
    const int i0 = __ldg(col + (c + 0) * DC_SLICE), i1 = __ldg(col + (c + 1) * DC_SLICE);
    const int i2 = __ldg(col + (c + 2) * DC_SLICE), i3 = __ldg(col + (c + 3) * DC_SLICE);
    DC_BWD(i0) DC_BWD(i1) DC_BWD(i2) DC_BWD(i3)
  }
  for (; c < width; ++c) {
    const int i0 = __ldg(col + c * DC_SLICE);
    DC_BWD(i0)
  }
#undef DC_BWD
  double* o = g_out + 3 * (size_t)order[row];
  o[0] = gx; o[1] = gy; o[2] = gz;
}

// C1, scatter form (no transposed graph needed): row i walks its own neighbour list and adds
// u_i A_i (p_j - m_i) to g_j with fire-and-forget fp64 reductions (RED.E.ADD.F64, addresses spread over
// all points -> no hot spots).  Used for the first backward passes on an asymmetric (kNN) graph, before
// building the transpose has paid off; g is accumulated in SORTED space and must be zeroed by the caller.
// F32 = true: one 16-byte vector reduction (red.global.add.v4.f32) per edge into a float4 accumulator instead of
// three fp64 reductions.  The L2 retires ~220 G reductions / s whatever their width (tools/micro/red_bench.cu), so
// this form is 3x faster; the accumulators then carry fp32 rounding (relative 6e-8 per addition, random sign), which
// averages out over the >= 1e6 points the chain stage sums -- used for large maps only, see fused.py.
template <bool F32>
__global__ void __launch_bounds__(STEP_THREADS)
step_backward_scatter_kernel(const dc_point* __restrict__ P, int64_t n, const int64_t* __restrict__ slice_ptr,
                             const int32_t* __restrict__ ell_idx, const dc_stash* __restrict__ stash,
                             const double* __restrict__ upstream, void* __restrict__ g_out) {
  double* g_sorted = reinterpret_cast<double*>(g_out);
  float4* g_sorted32 = reinterpret_cast<float4*>(g_out);
  const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t base = slice_ptr[row >> 5];
  const int width = (int)((slice_ptr[(row >> 5) + 1] - base) >> 5);
  const int32_t* col = ell_idx + base + lane;
  double4 s0, s1;
  dc_ld256(stash + row, s0.x, s0.y, s0.z, s0.w);
  dc_ld256(reinterpret_cast<const char*>(stash + row) + 32, s1.x, s1.y, s1.z, s1.w);
  double u = upstream ? upstream[row] : 1.0;
  const double ua = u * s1.z, ub = u * s1.w;
  if (ua == 0.0 && ub == 0.0) return;        // masked-out / inactive loss term: nothing to scatter
#define DC_SCAT(j_)                                                                  \
  if ((j_) >= 0) {                                                                   \
    const dc_point pj = dc_ld_point(P + (j_));                                       \
    const double ex = pj.x - s0.x, ey = pj.y - s0.y, ez = pj.z - s0.z;               \
    const double a = ua * (s0.w * ex + s1.x * ey + s1.y * ez);                       \
    const double vx = a * s0.w + ub * ex, vy = a * s1.x + ub * ey, vz = a * s1.y + ub * ez; \
    if (F32) {                                                                       \
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g_sorted32 + (j_)), "f"((float)vx), \
                   "f"((float)vy), "f"((float)vz), "f"(0.f) : "memory");             \
    } else {                                                                         \
      double* o = g_sorted + 3 * (size_t)(j_);                                       \
      atomicAdd(o, vx);                                                              \
      atomicAdd(o + 1, vy);                                                          \
      atomicAdd(o + 2, vz);                                                          \
    }                                                                                \
  }
  int c = 0;
  for (; c + 4 <= width; c += 4) {
    const int j0 = __ldg(col + (c + 0) * DC_SLICE), j1 = __ldg(col + (c + 1) * DC_SLICE);
    const int j2 = __ldg(col + (c + 2) * DC_SLICE), j3 = __ldg(col + (c + 3) * DC_SLICE);
    DC_SCAT(j0) DC_SCAT(j1) DC_SCAT(j2) DC_SCAT(j3)
  }
  for (; c < width; ++c) {
    const int j0 = __ldg(col + c * DC_SLICE);
    DC_SCAT(j0)
  }
#undef DC_SCAT
}

extern "C" int dc_step_backward_scatter(const void* points, int64_t n, const int64_t* slice_ptr, const int32_t* ell_idx,
                                        const double* stash, const double* upstream_pp, void* g_sorted, int g_dtype,
                                        void* stream) {
  if (n <= 0) return DC_OK;
  const int blocks = dc_blocks(n, STEP_THREADS);
  if (g_dtype == DC_F32)
    step_backward_scatter_kernel<true><<<blocks, STEP_THREADS, 0, (cudaStream_t)stream>>>((const dc_point*)points, n, slice_ptr, ell_idx,
                                                                                         (const dc_stash*)stash, upstream_pp, g_sorted);
  else
    step_backward_scatter_kernel<false><<<blocks, STEP_THREADS, 0, (cudaStream_t)stream>>>((const dc_point*)points, n, slice_ptr, ell_idx,
                                                                                          (const dc_stash*)stash, upstream_pp, g_sorted);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_step_backward(const void* points, int64_t n, const int64_t* slice_ptr_t, const int32_t* ell_idx_t,
                                const double* stash, const double* upstream_pp, const int32_t* order, double* g_out,
                                void* stream) {
  if (n <= 0) return DC_OK;
  const int blocks = dc_blocks(n, STEP_THREADS);
  step_backward_gather_kernel<<<blocks, STEP_THREADS, 0, (cudaStream_t)stream>>>((const dc_point*)points, n, slice_ptr_t, ell_idx_t,
                                                                                (const dc_stash*)stash, upstream_pp, order, g_out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

#define CHAIN_THREADS 256
#define CHAIN_REC (12 + 2 * DC_MAX_TERMS)   // doubles per partial record: pose 3x4, dw, dexp

template <typename T>
__global__ void __launch_bounds__(CHAIN_THREADS)
step_chain_kernel(const void* __restrict__ g_any, int g_f32, const int32_t* __restrict__ g_index,
                  const typename vec4_of<T>::type* __restrict__ rec_dir,
                  const typename vec4_of<T>::type* __restrict__ rec_vp, const uint32_t* __restrict__ rec_meta,
                  const int32_t* __restrict__ block_scan, const int64_t* __restrict__ block_start,
                  const int32_t* __restrict__ block_count, const double* __restrict__ poses, dc_model model,
                  int want_exp, double* __restrict__ partials) {
  const int scan = block_scan[blockIdx.x];
  const int64_t first = block_start[blockIdx.x];
  const int count = block_count[blockIdx.x];
  const double* Tm = poses + 12 * (size_t)scan;
  const double r00 = Tm[0], r01 = Tm[1], r02 = Tm[2], r10 = Tm[4], r11 = Tm[5], r12 = Tm[6], r20 = Tm[8], r21 = Tm[9], r22 = Tm[10];
  double acc[CHAIN_REC];
#pragma unroll
  for (int k = 0; k < CHAIN_REC; ++k) acc[k] = 0.0;
  for (int t = threadIdx.x; t < count; t += CHAIN_THREADS) {
    const int64_t i = first + t;
    const typename vec4_of<T>::type a = rec_dir[i], b = rec_vp[i];
    const uint32_t meta = rec_meta[i];
    const size_t gi = g_index ? (size_t)g_index[i] : (size_t)i;   // g in sorted space (scatter form) or in place
    double gx, gy, gz;
    if (g_f32) {
      const float4 gv = reinterpret_cast<const float4*>(g_any)[gi];
      gx = (double)gv.x; gy = (double)gv.y; gz = (double)gv.z;
    } else {
      const double* g = reinterpret_cast<const double*>(g_any);
      gx = g[3 * gi]; gy = g[3 * gi + 1]; gz = g[3 * gi + 2];
    }
    const bool mm = (model.kind != DC_MODEL_NONE) && (meta & DC_PT_MODEL_MASK);
    double pw[DC_MAX_TERMS];
    const double d0 = (double)a.w, gam = (double)b.w;
    const double d = dc_correct_depth(model, mm, d0, gam, pw);
    const double dx = (double)a.x, dy = (double)a.y, dz = (double)a.z;
    // d L / d d' = (R dir) . g
    const double gd = (r00 * dx + r01 * dy + r02 * dz) * gx + (r10 * dx + r11 * dy + r12 * dz) * gy +
                      (r20 * dx + r21 * dy + r22 * dz) * gz;
    if (mm) {
      const double f = (model.kind == DC_MODEL_SCALED_POLYNOMIAL ? -d0 : -1.0) * gd;
#pragma unroll
      for (int k = 0; k < DC_MAX_TERMS; ++k) {
        if (k < model.n_terms) {
          acc[12 + k] += f * pw[k];
          if (want_exp) acc[12 + DC_MAX_TERMS + k] += f * model.w[k] * dc_pow_exp_dlog(gam, model.e[k], pw[k]);
        }
      }
    }
    const double x = (double)b.x + d * dx, y = (double)b.y + d * dy, z = (double)b.z + d * dz;
    acc[0] += gx * x; acc[1] += gx * y; acc[2] += gx * z; acc[3] += gx;
    acc[4] += gy * x; acc[5] += gy * y; acc[6] += gy * z; acc[7] += gy;
    acc[8] += gz * x; acc[9] += gz * y; acc[10] += gz * z; acc[11] += gz;
  }
  __shared__ double red[CHAIN_THREADS / 32][CHAIN_REC];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < CHAIN_REC; ++k) {
    const double v = dc_warp_sum(acc[k]);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < CHAIN_REC) {
    double v = 0.0;
#pragma unroll
    for (int wv = 0; wv < CHAIN_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
    partials[(size_t)blockIdx.x * CHAIN_REC + threadIdx.x] = v;
  }
}

// one block: pose gradients per scan (blocks of a scan are contiguous), model gradients over all blocks
__global__ void __launch_bounds__(256)
step_chain_reduce_kernel(const double* __restrict__ partials, int n_blocks, const int32_t* __restrict__ scan_block_first,
                         int n_scans, int n_terms, double* __restrict__ dw, double* __restrict__ dexp,
                         double* __restrict__ dposes) {
  for (int idx = threadIdx.x; idx < n_scans * 12; idx += 256) {
    const int s = idx / 12, k = idx - 12 * s;
    double v = 0.0;
    for (int b = scan_block_first[s]; b < scan_block_first[s + 1]; ++b) v += partials[(size_t)b * CHAIN_REC + k];
    dposes[idx] += v;
  }
  __shared__ double red[256];
  for (int k = 0; k < 2 * n_terms; ++k) {
    const int col = 12 + (k < n_terms ? k : DC_MAX_TERMS + (k - n_terms));
    double* out = k < n_terms ? dw : dexp;
    if (!out) continue;
    double v = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 256) v += partials[(size_t)b * CHAIN_REC + col];
    red[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) out[k < n_terms ? k : k - n_terms] += red[0];
    __syncthreads();
  }
}

extern "C" int dc_step_chain(const void* g, int g_dtype, const int32_t* g_index, const void* rec_dir, const void* rec_vp, const uint32_t* rec_meta, int dtype,
                             const int32_t* block_scan, const int64_t* block_start, const int32_t* block_count,
                             int n_blocks, const int32_t* scan_block_first, const double* poses, int n_scans,
                             int model_kind, const double* w, const double* exponent, int n_terms, double* partials,
                             double* dw, double* dexponent, double* dposes, void* stream) {
  if (n_blocks <= 0) return DC_OK;
  if (n_terms < 0 || n_terms > DC_MAX_TERMS) return dc_set_error(DC_ERR_ARG, "dc_step_chain: too many polynomial terms");
  dc_model m = {model_kind, n_terms, w, exponent};
  cudaStream_t st = (cudaStream_t)stream;
  const int want_exp = dexponent != nullptr;
  if (dtype == DC_F32)
    step_chain_kernel<float><<<n_blocks, CHAIN_THREADS, 0, st>>>(g, g_dtype == DC_F32, g_index, (const float4*)rec_dir, (const float4*)rec_vp, rec_meta, block_scan,
                                                                 block_start, block_count, poses, m, want_exp, partials);
  else
    step_chain_kernel<double><<<n_blocks, CHAIN_THREADS, 0, st>>>(g, g_dtype == DC_F32, g_index, (const double4*)rec_dir, (const double4*)rec_vp, rec_meta, block_scan,
                                                                  block_start, block_count, poses, m, want_exp, partials);
  DC_LAUNCH_CHECK();
  step_chain_reduce_kernel<<<1, 256, 0, st>>>(partials, n_blocks, scan_block_first, n_scans, model_kind != DC_MODEL_NONE ? n_terms : 0,
                                              dw, dexponent, dposes);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// SE(3) pose corrections (eval.py:68-82, transform.py:68-78): one thread per scan.
// ---------------------------------------------------------------------------------------------
__global__ void pose_compose_kernel(const double* __restrict__ poses, const double* __restrict__ deltas, int n_scans,
                                    int n_deltas, double* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  double P[16], d[6], T[12];
  for (int k = 0; k < 16; ++k) P[k] = poses[16 * s + k];
  const double* dp = deltas + 6 * (n_deltas == 1 ? 0 : s);
  for (int k = 0; k < 6; ++k) d[k] = dp[k];
  dc_pose_compose(P, d, T);
  for (int k = 0; k < 12; ++k) out[12 * s + k] = T[k];
}

__global__ void pose_compose_bwd_kernel(const double* __restrict__ poses, const double* __restrict__ deltas, int n_scans,
                                        int n_deltas, const double* __restrict__ dout, double* __restrict__ ddeltas) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  double P[16], d[6], g[12], gd[6];
  for (int k = 0; k < 16; ++k) P[k] = poses[16 * s + k];
  const double* dp = deltas + 6 * (n_deltas == 1 ? 0 : s);
  for (int k = 0; k < 6; ++k) d[k] = dp[k];
  for (int k = 0; k < 12; ++k) g[k] = dout[12 * s + k];
  dc_pose_compose_bwd(P, d, g, gd);
  if (n_deltas == 1) {
    for (int k = 0; k < 6; ++k) atomicAdd(ddeltas + k, gd[k]);
  } else {
    for (int k = 0; k < 6; ++k) ddeltas[6 * s + k] = gd[k];
  }
}

extern "C" int dc_pose_compose(const double* poses, const double* deltas, int n_scans, int n_deltas, double* out, void* stream) {
  if (n_scans <= 0) return DC_OK;
  if (n_deltas != 1 && n_deltas != n_scans) return dc_set_error(DC_ERR_ARG, "dc_pose_compose: n_deltas must be 1 or n_scans");
  pose_compose_kernel<<<dc_blocks(n_scans, 64), 64, 0, (cudaStream_t)stream>>>(poses, deltas, n_scans, n_deltas, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

extern "C" int dc_pose_compose_backward(const double* poses, const double* deltas, int n_scans, int n_deltas,
                                        const double* dout, double* ddeltas, void* stream) {
  if (n_scans <= 0) return DC_OK;
  if (n_deltas != 1 && n_deltas != n_scans) return dc_set_error(DC_ERR_ARG, "dc_pose_compose_backward: n_deltas must be 1 or n_scans");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_deltas == 1) DC_CUDA_CHECK(cudaMemsetAsync(ddeltas, 0, 6 * sizeof(double), st));
  pose_compose_bwd_kernel<<<dc_blocks(n_scans, 64), 64, 0, st>>>(poses, deltas, n_scans, n_deltas, dout, ddeltas);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
