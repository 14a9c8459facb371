// Kernel 1, kNN / kNN-within-r mode: k-nearest SELECTION on the cell-sorted map.
// Replaces cKDTree.query(k, distance_upper_bound=r) (nearest_neighbors.py:48-49).
//
// The k nearest of a query are selected, not sorted: a histogram of the squared distances finds the bin that
// holds the k-th distance, everything below that bin is a neighbour, and only the few candidates inside the
// boundary bin are ranked (by (d2, sorted index)).  Every d2 is computed with the identical non-fused
// instruction sequence wherever it is needed, so the classification of a candidate never changes between
// passes and the selection is exact.  Rows are emitted UNSORTED (the step kernels only need the set);
// dc_knn_sort_rows orders them by distance when the reference layout is exported.
//
// One query per thread (see knn_thread_query).  Two alternatives were built, were bit-exact, and lost on lidar maps
// (profiles/r1c_knn_experiments.md): a warp-cooperative kernel (lanes = queries, the union of the candidate cells of a
// slice staged once in shared memory as fp32 offsets with an fp64 tie-break; consecutive queries share too few
// candidates -- union 1.6x the own set -- and every warp pays for its sparsest lane), and skipping rings from the
// cell-table population with smaller cells (the scan is bound by per-row latency and divergence, not by candidates).
#include "dc_common.cuh"
#include "dc_grid.cuh"

#define KNN_THREADS 128
#define KNN_WARPS (KNN_THREADS / 32)
#define KNN_BINS 64

__device__ __forceinline__ bool knn_less(double a, int ja, double b, int jb) { return a < b || (a == b && ja < jb); }

__device__ __forceinline__ int knn_bin(double d2, double scale) {
  const int b = __double2int_rz(d2 * scale);
  return b > KNN_BINS - 1 ? KNN_BINS - 1 : b;
}

template <typename F>
__device__ __forceinline__ void knn_scan(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                         const int32_t* __restrict__ cell_start, const dc_point* __restrict__ P,
                                         const dc_point& pq, int c0, int c1, int c2, int rho, F&& f) {
  for (int e2 = -rho; e2 <= rho; ++e2) {
    for (int e1 = -rho; e1 <= rho; ++e1) {
      int lo, hi;
      dc_row_range(g, pkeys, n, cell_start, c0 - rho, c0 + rho, c1 + e1, c2 + e2, lo, hi);
      // four independent candidate loads in flight per thread (the loop is latency bound otherwise; a software
      // pipeline with eight in flight cost registers / occupancy and was 30 % slower).  The tail of a row goes
      // through the same four-wide body with clamped addresses: a one-at-a-time tail loop exposed a full load
      // latency per candidate on up to three candidates of every row.
      for (int j = lo; j < hi; j += 4) {
        const int last = hi - 1;
        const int j1 = j + 1 < last ? j + 1 : last, j2 = j + 2 < last ? j + 2 : last, j3 = j + 3 < last ? j + 3 : last;
        const dc_point p0 = dc_ld_point(P + j), p1 = dc_ld_point(P + j1);
        const dc_point p2 = dc_ld_point(P + j2), p3 = dc_ld_point(P + j3);
        const double d0 = dc_dist2(p0, pq), d1 = dc_dist2(p1, pq), d2 = dc_dist2(p2, pq), d3 = dc_dist2(p3, pq);
        f(j, d0);
        if (j + 1 < hi) f(j + 1, d1);
        if (j + 2 < hi) f(j + 2, d2);
        if (j + 3 < hi) f(j + 3, d3);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Thread path: one query per thread.  h = this thread's private histogram column (stride KNN_THREADS).
// The ring of cells grows until k candidates lie inside the radius the block is guaranteed to cover; a second
// histogram level splits the boundary bin when it holds more than 8 candidates.
// ---------------------------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ void knn_thread_query(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                                                 const dc_grid& g, const int32_t* __restrict__ cell_start,
                                                 const dc_point& pq, int c0, int c1, int c2, int k, double r2cap,
                                                 int max_ring, int first_ring, unsigned short* h, Emit&& emit) {
  const double slack_cell = g.cell * (1.0 - 1e-9);
  // ---- 1. ring growth + level-1 histogram
  int rho = first_ring;
  double bound2, scale1;
  unsigned int n_in;
  for (;;) {
    if (rho > max_ring) rho = max_ring;
    const bool last = rho >= max_ring;
    const double reach = rho * slack_cell;
    bound2 = last ? r2cap : fmin(reach * reach, r2cap);
    scale1 = (double)KNN_BINS / bound2;
#pragma unroll
    for (int b = 0; b < KNN_BINS; ++b) h[b * KNN_THREADS] = (unsigned short)0;
    n_in = 0u;
    knn_scan(g, pkeys, n, cell_start, P, pq, c0, c1, c2, rho, [&](int j, double d2) {
      if (d2 < bound2) {
        const int b = knn_bin(d2, scale1);
        const unsigned short v = h[b * KNN_THREADS];
        h[b * KNN_THREADS] = v == 65535 ? v : (unsigned short)(v + 1);
        ++n_in;
      }
    });
    if (n_in >= (unsigned int)k || last) break;
    rho = rho < 4 ? rho + 1 : rho * 2;
  }
  if (n_in <= (unsigned int)k) {
    // everything inside the bound is a neighbour (fewer than k exist within r / in the map)
    knn_scan(g, pkeys, n, cell_start, P, pq, c0, c1, c2, rho, [&](int j, double d2) {
      if (d2 < bound2) emit(j, d2);
    });
    return;
  }
  // ---- level 1: bin of the k-th distance
  unsigned int c_lo = 0u, cnt1 = 0u;
  int b1 = 0;
  for (; b1 < KNN_BINS; ++b1) {
    cnt1 = h[b1 * KNN_THREADS];
    if (c_lo + cnt1 >= (unsigned int)k) break;
    c_lo += cnt1;
  }
  int b2 = KNN_BINS;          // boundary sub-bin; KNN_BINS = "no second level: bin b1 is the boundary set"
  unsigned int cnt2 = cnt1;
  const bool lvl2 = cnt1 > 8u && c_lo + cnt1 > (unsigned int)k;
  if (lvl2) {
    // ---- 2. level-2 histogram inside bin b1
#pragma unroll
    for (int b = 0; b < KNN_BINS; ++b) h[b * KNN_THREADS] = (unsigned short)0;
    knn_scan(g, pkeys, n, cell_start, P, pq, c0, c1, c2, rho, [&](int j, double d2) {
      if (d2 < bound2) {
        const double s = d2 * scale1;
        int b = __double2int_rz(s);
        b = b > KNN_BINS - 1 ? KNN_BINS - 1 : b;
        if (b == b1) {
          int bb = __double2int_rz((s - (double)b1) * (double)KNN_BINS);
          bb = bb < 0 ? 0 : (bb > KNN_BINS - 1 ? KNN_BINS - 1 : bb);
          const unsigned short v = h[bb * KNN_THREADS];
          h[bb * KNN_THREADS] = v == 65535 ? v : (unsigned short)(v + 1);
        }
      }
    });
    for (b2 = 0; b2 < KNN_BINS; ++b2) {
      cnt2 = h[b2 * KNN_THREADS];
      if (c_lo + cnt2 >= (unsigned int)k) break;
      c_lo += cnt2;
    }
  }
  // ---- 3. emit: everything below the boundary (sub-)bin, and the t smallest of the boundary sub-bin
  const unsigned int t = (unsigned int)k - c_lo;     // how many of the cnt2 boundary candidates are neighbours
  const bool take_all = (t == cnt2);
  const bool use_list = !take_all && cnt2 <= 8u;
  // The (<= 8) candidates of the boundary bin are only collected during the scan (in this thread's histogram
  // column, which is no longer needed) and ranked afterwards with the warp converged: ranking inside the scan ran one lane at a time and
  // cost 17 % of all instructions of the kernel.
  int nb = 0;
  knn_scan(g, pkeys, n, cell_start, P, pq, c0, c1, c2, rho, [&](int j, double d2) {
    if (d2 < bound2) {
      const double s = d2 * scale1;
      int b = __double2int_rz(s);
      b = b > KNN_BINS - 1 ? KNN_BINS - 1 : b;
      if (b < b1) {
        emit(j, d2);
      } else if (b == b1) {
        bool boundary = true;
        if (lvl2) {
          int bb = __double2int_rz((s - (double)b1) * (double)KNN_BINS);
          bb = bb < 0 ? 0 : (bb > KNN_BINS - 1 ? KNN_BINS - 1 : bb);
          if (bb < b2) emit(j, d2);
          boundary = (bb == b2);
        }
        if (boundary) {
          if (take_all) {
            emit(j, d2);
          } else if (use_list && nb < 8) {
            // the histogram is dead by now: entry nb lives in this thread's counters 6 nb .. 6 nb + 5
            // (four 16-bit pieces of d2, two of j), which keeps the block at 16 KB of shared memory
            const unsigned long long u = (unsigned long long)__double_as_longlong(d2);
            unsigned short* e = h + 6 * nb * KNN_THREADS;
            e[0] = (unsigned short)u;
            e[KNN_THREADS] = (unsigned short)(u >> 16);
            e[2 * KNN_THREADS] = (unsigned short)(u >> 32);
            e[3 * KNN_THREADS] = (unsigned short)(u >> 48);
            e[4 * KNN_THREADS] = (unsigned short)j;
            e[5 * KNN_THREADS] = (unsigned short)((unsigned int)j >> 16);
            ++nb;
          }
        }
      }
    }
  });
  if (take_all) return;
  if (use_list) {
    double bd[8];
    int bj[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned short* e = h + 6 * i * KNN_THREADS;
      const unsigned long long u = (unsigned long long)e[0] | ((unsigned long long)e[KNN_THREADS] << 16) |
                                   ((unsigned long long)e[2 * KNN_THREADS] << 32) | ((unsigned long long)e[3 * KNN_THREADS] << 48);
      const int j = (int)((unsigned int)e[4 * KNN_THREADS] | ((unsigned int)e[5 * KNN_THREADS] << 16));
      bd[i] = i < nb ? __longlong_as_double((long long)u) : INFINITY;
      bj[i] = i < nb ? j : 0x7fffffff;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned int rank = 0u;          // candidates of the bin that precede candidate i in (d2, index) order
#pragma unroll
      for (int m = 0; m < 8; ++m) rank += knn_less(bd[m], bj[m], bd[i], bj[i]) ? 1u : 0u;
      if (i < nb && rank < t) emit(bj[i], bd[i]);
    }
    return;
  }
  // more than 8 candidates share the boundary sub-bin (exact ties / duplicates): repeated minimum
  // selection in (d2, index) order -- O(t * candidates), rare
  double last_d = -1.0;
  int last_j = -1;
  for (unsigned int s_ = 0; s_ < t; ++s_) {
    double best_d = INFINITY;
    int best_j = 0x7fffffff;
    knn_scan(g, pkeys, n, cell_start, P, pq, c0, c1, c2, rho, [&](int j, double d2) {
      if (d2 < bound2) {
        const double s = d2 * scale1;
        int b = __double2int_rz(s);
        b = b > KNN_BINS - 1 ? KNN_BINS - 1 : b;
        if (b == b1) {
          int bb = b2;
          if (lvl2) {
            bb = __double2int_rz((s - (double)b1) * (double)KNN_BINS);
            bb = bb < 0 ? 0 : (bb > KNN_BINS - 1 ? KNN_BINS - 1 : bb);
          }
          if (bb == b2 && knn_less(last_d, last_j, d2, j) && knn_less(d2, j, best_d, best_j)) {
            best_d = d2;
            best_j = j;
          }
        }
      }
    });
    emit(best_j, best_d);
    last_d = best_d;
    last_j = best_j;
  }
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_thread_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                  const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
                  const int32_t* __restrict__ cell_start, int k, double r2cap, int max_ring,
                  int32_t* __restrict__ ell_idx, double* __restrict__ ell_d2) {
  __shared__ unsigned short hist[KNN_BINS][KNN_THREADS];
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int32_t* out_j = ell_idx + (q >> 5) * (int64_t)k * DC_SLICE + lane;
  double* out_d = ell_d2 ? ell_d2 + (q >> 5) * (int64_t)k * DC_SLICE + lane : nullptr;
  int cnt = 0;
  if (q < nq) {
    const dc_point pq = dc_ld_point(Q + q);
    int c0, c1, c2;
    dc_key_coords(g, qkeys[q], c0, c1, c2);
    knn_thread_query(P, pkeys, n, g, cell_start, pq, c0, c1, c2, k, r2cap, max_ring, 1, &hist[0][threadIdx.x],
                     [&](int j, double d2) {
                       out_j[(int64_t)cnt * DC_SLICE] = j;
                       if (out_d) out_d[(int64_t)cnt * DC_SLICE] = d2;
                       ++cnt;
                     });
  }
  if (nq > 0 && (q >> 5) <= ((nq - 1) >> 5)) {
    for (int c = cnt; c < k; ++c) {
      out_j[(int64_t)c * DC_SLICE] = -1;
      if (out_d) out_d[(int64_t)c * DC_SLICE] = INFINITY;
    }
  }
}

extern "C" int dc_knn(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                      const dc_grid_spec* spec, const int32_t* cell_start, int k, double r, int32_t* ell_idx,
                      double* ell_d2, void* stream) {
  if (nq <= 0) return DC_OK;
  if (k < 1) return dc_set_error(DC_ERR_ARG, "dc_knn: k must be positive");
  dc_grid g;
  int rc = dc_make_grid(spec, &g);
  if (rc) return rc;
  int max_ring = g.d[0] > g.d[1] ? g.d[0] : g.d[1];
  max_ring = max_ring > g.d[2] ? max_ring : g.d[2];
  // finite cap on every squared distance inside (or clamped into) the grid box, used when there is no r
  const double ex = g.d[0] * g.cell, ey = g.d[1] * g.cell, ez = g.d[2] * g.cell;
  double r2cap = 16.0 * (ex * ex + ey * ey + ez * ez) + 1.0;
  if (r > 0.0) {
    r2cap = r * r;      // cKDTree: d2 < distance_upper_bound ** 2, strict
    const int rr = (int)ceil(r / g.cell);
    if (rr < max_ring) max_ring = rr;
  }
  if (max_ring < 1) max_ring = 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_slices = (nq + DC_SLICE - 1) / DC_SLICE;
  const int blocks = dc_blocks(n_slices * DC_SLICE, KNN_THREADS);
  knn_thread_kernel<<<blocks, KNN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g, cell_start, k,
                                                    r2cap, max_ring, ell_idx, ell_d2);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Order every row by (d2, index): the reference returns distance-sorted rows.  Export path only.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void knn_sort_rows_kernel(int k, int32_t* __restrict__ ell_idx, double* __restrict__ ell_d2, int64_t nq) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const int lane = (int)(q & 31);
  int32_t* pj = ell_idx + (q >> 5) * (int64_t)k * DC_SLICE + lane;
  double* pd = ell_d2 + (q >> 5) * (int64_t)k * DC_SLICE + lane;
  double d[KMAX];
  int j[KMAX];
  int m = 0;
  for (int c = 0; c < k; ++c) {
    const int jj = pj[(int64_t)c * DC_SLICE];
    if (jj < 0) continue;
    const double dd = pd[(int64_t)c * DC_SLICE];
    int pos = m++;
    while (pos > 0 && knn_less(dd, jj, d[pos - 1], j[pos - 1])) { d[pos] = d[pos - 1]; j[pos] = j[pos - 1]; --pos; }
    d[pos] = dd;
    j[pos] = jj;
  }
  for (int c = 0; c < k; ++c) {
    pj[(int64_t)c * DC_SLICE] = c < m ? j[c] : -1;
    pd[(int64_t)c * DC_SLICE] = c < m ? d[c] : INFINITY;
  }
}

extern "C" int dc_knn_sort_rows(int k, int32_t* ell_idx, double* ell_d2, int64_t nq, void* stream) {
  if (nq <= 0) return DC_OK;
  if (k < 1 || k > 1024) return dc_set_error(DC_ERR_ARG, "dc_knn_sort_rows: k must be in [1, 1024]");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(nq, 128);
  if (k <= 32) knn_sort_rows_kernel<32><<<blocks, 128, 0, st>>>(k, ell_idx, ell_d2, nq);
  else if (k <= 128) knn_sort_rows_kernel<128><<<blocks, 128, 0, st>>>(k, ell_idx, ell_d2, nq);
  else knn_sort_rows_kernel<1024><<<blocks, 128, 0, st>>>(k, ell_idx, ell_d2, nq);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// Squared distances of a kNN graph, recomputed from the records with the same instruction sequence the
// selection used (bit-identical values).  The hot search does not store them (8 bytes per edge would be the
// largest write of the whole search); they are only needed to export distance-sorted rows and `distances`.
__global__ void knn_distances_kernel(const dc_point* __restrict__ P, const dc_point* __restrict__ Q, int k,
                                     const int32_t* __restrict__ ell_idx, int64_t nq, double* __restrict__ ell_d2) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if ((q >> 5) > ((nq - 1) >> 5)) return;
  const int lane = (int)(q & 31);
  const int64_t base = (q >> 5) * (int64_t)k * DC_SLICE + lane;
  if (q >= nq) {
    for (int c = 0; c < k; ++c) ell_d2[base + (int64_t)c * DC_SLICE] = INFINITY;
    return;
  }
  const dc_point pq = dc_ld_point(Q + q);
  for (int c = 0; c < k; ++c) {
    const int j = ell_idx[base + (int64_t)c * DC_SLICE];
    ell_d2[base + (int64_t)c * DC_SLICE] = j >= 0 ? dc_dist2(dc_ld_point(P + j), pq) : INFINITY;
  }
}

extern "C" int dc_knn_distances(const void* P, const void* Q, int k, const int32_t* ell_idx, int64_t nq, double* ell_d2,
                                void* stream) {
  if (nq <= 0) return DC_OK;
  const int blocks = dc_blocks(((nq + 31) / 32) * 32, 128);
  knn_distances_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>((const dc_point*)P, (const dc_point*)Q, k, ell_idx, nq, ell_d2);
  DC_LAUNCH_CHECK();
  return DC_OK;
}
