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
#include <cub/cub.cuh>
#include <stdlib.h>
#include <string.h>
#include "dc_common.cuh"
#include "dc_grid.cuh"

#define KNN_THREADS 128
#define KNN_WARPS (KNN_THREADS / 32)
#define KNN_BINS 64

// (d2, original index) lexicographic order.  The original index (dc_point.tag) -- not the position in the cell-sorted
// map -- breaks exact ties, so the selected set does not depend on the cell size.  The tag is only loaded when two
// squared distances are bit-equal; j >= n stands for "no candidate" and sorts last.
__device__ __forceinline__ long long knn_tag(const dc_point* __restrict__ P, int64_t n, int j) {
  return (j >= 0 && (int64_t)j < n) ? P[j].tag : (j < 0 ? -1LL : 0x7fffffffffffffffLL);
}
__device__ __forceinline__ bool knn_less(const dc_point* __restrict__ P, int64_t n, double a, int ja, double b, int jb) {
  return a < b || (a == b && knn_tag(P, n, ja) < knn_tag(P, n, jb));
}

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
      for (int m = 0; m < 8; ++m) rank += knn_less(P, n, bd[m], bj[m], bd[i], bj[i]) ? 1u : 0u;
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
          if (bb == b2 && knn_less(P, n, last_d, last_j, d2, j) && knn_less(P, n, d2, j, best_d, best_j)) {
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
// Cell path (dc_knn_cells): one WARP per occupied query cell.
//
// Every query of a cell has the identical candidate block ((2m+1)^2 rows of cells x [c0-m, c0+m] along the fastest
// axis), so the warp stages that block ONCE -- one candidate per lane and register slot, as fp32 offsets from the cell
// centre plus |c|^2 -- and then streams the cell's queries through it: lanes = candidates, the query is warp-uniform.
// Per query: v = 63 * |c - q|^2 / bound for 32 candidates per instruction (3 FFMA from |c|^2 - 2 q.c + |q|^2), a
// 63-bin shared-memory histogram of v (one atomic per in-range candidate), a warp scan for the bin of the k-th
// distance, and ONE emit sweep over the v's still sitting in registers (ballot + popc compaction into a shared tile
// that is written out column-wise, coalesced per slice).  One distance evaluation per (query, candidate) pair.
//
// Exactness.  fp32 only CLASSIFIES: |v_fp32 - v_exact| <= dv/2 (ring table below), so the selection is provably the
// fp64 one whenever the gap between the last selected and the first rejected candidate exceeds dv (and, for queries
// with fewer than k candidates inside r, no candidate lies in the band dr around r^2).  Every other query -- 0.1-1 %
// on lidar maps -- plus cells whose block exceeds the register budget or needs more than KT_MMAX rings goes to a
// list that the one-thread-per-query fp64 kernel above (knn_thread_query) finishes.  Both paths select by
// (fp64 d2, original index), so the union is bit-identical to the thread path alone.
// ---------------------------------------------------------------------------------------------
#define KT_WARPS 4
#define KT_THREADS (KT_WARPS * 32)
#define KT_NB 8                    // register slots per lane: blocks of up to 256 candidates
#define KT_CMAX (KT_NB * 32)
#define KT_MMAX 4
#define KT_ROWCAP 96               // >= (2 KT_MMAX + 1)^2 rows
#define KT_LIST 32
#define KT_GRAB 8                  // cells fetched per atomic
#define KT_WARP_WORDS (KT_CMAX + KT_ROWCAP + 100 + 64 + KT_LIST + KT_LIST)

struct kt_ring {
  float sc;    // bins per unit of d2: 63 / (usable bound)
  float dv;    // fp32 classification uncertainty in bins (two candidates closer than this are re-ranked in fp64)
  float dr;    // band above bin 62 in which `d2 < r^2` cannot be decided in fp32
  int rlim;    // ring covers r (or the whole grid): fewer than k candidates in range is final
  int usable;
};
struct kt_params {
  kt_ring ring[KT_MMAX + 1];
  int mmax, k, pop_min, tile_stride;
};

#define KT_FULL 0xffffffffu

// Non-empty rows of the block of ring m around cell (c0, c1, c2): s_rlo[i] = first sorted position, s_rpre[i] =
// candidates before row i; returns the candidate count, n_rows by reference.
__device__ __forceinline__ int kt_rows(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                       const int32_t* __restrict__ cell_start, int c0, int c1, int c2, int m, int lane,
                                       int* s_rlo, int* s_rpre, int& n_rows) {
  const int w = 2 * m + 1, R = w * w;
  const unsigned lt = (1u << lane) - 1u;
  int base_rows = 0, base_cnt = 0;
  __syncwarp();
  for (int r0 = 0; r0 < R; r0 += 32) {
    const int rr = r0 + lane;
    int lo = 0, hi = 0;
    if (rr < R) dc_row_range(g, pkeys, n, cell_start, c0 - m, c0 + m, c1 + (rr % w) - m, c2 + (rr / w) - m, lo, hi);
    const int cnt = hi - lo;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(KT_FULL, inc, d);
      if (lane >= d) inc += t;
    }
    const unsigned ne = __ballot_sync(KT_FULL, cnt > 0);
    if (cnt > 0) {
      const int pos = base_rows + __popc(ne & lt);
      s_rlo[pos] = lo;
      s_rpre[pos] = base_cnt + inc - cnt;
    }
    base_rows += __popc(ne);
    base_cnt += __shfl_sync(KT_FULL, inc, 31);
  }
  if (lane == 0) s_rpre[base_rows] = base_cnt;
  __syncwarp();
  n_rows = base_rows;
  return base_cnt;
}

__device__ __forceinline__ float kt_warp_max(float v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = fmaxf(v, __shfl_xor_sync(KT_FULL, v, d));
  return v;
}

__device__ __forceinline__ void kt_push(unsigned mask, int cs, int ring, int lane, int32_t* counters, int2* fb) {
  if ((mask >> lane) & 1u) {
    const int pos = atomicAdd(&counters[1], 1);
    fb[pos] = make_int2(cs + lane, ring);
  }
}

__global__ void __launch_bounds__(KT_THREADS)
knn_cell_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
                const int32_t* __restrict__ cell_start, kt_params prm, const int32_t* __restrict__ task_start,
                const int32_t* __restrict__ n_tasks_p, int32_t* counters, int2* fb_list, int32_t* __restrict__ ell_idx) {
  extern __shared__ __align__(16) int kt_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int k = prm.k, ts = prm.tile_stride;
  int* s_idx = kt_smem + wib * (KT_WARP_WORDS + 32 * ts);
  int* s_rlo = s_idx + KT_CMAX;
  int* s_rpre = s_rlo + KT_ROWCAP;
  unsigned* s_hist = (unsigned*)(s_rpre + 100);
  float* s_lv = (float*)(s_hist + 64);
  int* s_ls = (int*)(s_lv + KT_LIST);
  int* s_tile = s_ls + KT_LIST;
  // ring table in shared memory: indexing the kernel parameter with a run-time ring would copy it to local memory
  __shared__ kt_ring s_ring[KT_MMAX + 1];
#pragma unroll
  for (int i = 0; i <= KT_MMAX; ++i)
    if (threadIdx.x == i) s_ring[i] = prm.ring[i];
  __syncthreads();
  const int n_tasks = *n_tasks_p;
  const unsigned lt = (1u << lane) - 1u;
  int task_next = 0, task_end = 0;
  for (;;) {
    if (task_next >= task_end) {
      int t0 = 0;
      if (lane == 0) t0 = atomicAdd(&counters[0], KT_GRAB);
      t0 = __shfl_sync(KT_FULL, t0, 0);
      if (t0 >= n_tasks) break;
      task_next = t0;
      task_end = t0 + KT_GRAB < n_tasks ? t0 + KT_GRAB : n_tasks;
    }
    const int task = task_next++;
    const int s0 = task_start[task];
    const int s1 = task + 1 < n_tasks ? task_start[task + 1] : (int)nq;
    int c0, c1, c2;
    dc_key_coords(g, qkeys[s0], c0, c1, c2);
    // centre of the query cell = origin of the fp32 offsets (xyz order)
    const double o0 = g.org[0] + ((double)c0 + 0.5) * g.cell, o1 = g.org[1] + ((double)c1 + 0.5) * g.cell,
                 o2 = g.org[2] + ((double)c2 + 0.5) * g.cell;
    const double ox = g.ax[0] == 0 ? o0 : (g.ax[1] == 0 ? o1 : o2);
    const double oy = g.ax[0] == 1 ? o0 : (g.ax[1] == 1 ? o1 : o2);
    const double oz = g.ax[0] == 2 ? o0 : (g.ax[1] == 2 ? o1 : o2);

    float cx[KT_NB], cy[KT_NB], cz[KT_NB], cw[KT_NB];
    int staged_m = 0, m_start = 0, C = 0, nb = 0;
    for (int cs = s0; cs < s1; cs += 32) {
      const int Gc = s1 - cs < 32 ? s1 - cs : 32;
      unsigned pend = Gc == 32 ? KT_FULL : ((1u << Gc) - 1u);
      unsigned done = 0u;
      float qax = 0.f, qay = 0.f, qaz = 0.f, qk = 0.f;
      if (lane < Gc) {
        const dc_point pq = dc_ld_point(Q + cs + lane);
        const float x = (float)(pq.x - ox), y = (float)(pq.y - oy), z = (float)(pq.z - oz);
        qax = -2.f * x; qay = -2.f * y; qaz = -2.f * z;
        qk = fmaf(z, z, fmaf(y, y, x * x));
      }
      int m = m_start ? m_start : 1;
      while (pend) {
        if (staged_m != m) {
          int n_rows;
          C = kt_rows(g, pkeys, n, cell_start, c0, c1, c2, m, lane, s_rlo, s_rpre, n_rows);
          if (m_start == 0) {
            // first visit of this cell: skip rings whose block cannot hold k neighbours with some margin
            while (C < prm.pop_min && m < prm.mmax) {
              ++m;
              C = kt_rows(g, pkeys, n, cell_start, c0, c1, c2, m, lane, s_rlo, s_rpre, n_rows);
            }
            m_start = m;
          }
          staged_m = m;
          nb = (C + 31) >> 5;
          if (C <= KT_CMAX) {
            int rho = 0;
#pragma unroll
            for (int b = 0; b < KT_NB; ++b) {
              cx[b] = 0.f; cy[b] = 0.f; cz[b] = 0.f; cw[b] = 3.0e38f;
              const int i = 32 * b + lane;
              if (i < C) {
                while (s_rpre[rho + 1] <= i) ++rho;
                const int j = s_rlo[rho] + (i - s_rpre[rho]);
                const dc_point p = dc_ld_point(P + j);
                const float x = (float)(p.x - ox), y = (float)(p.y - oy), z = (float)(p.z - oz);
                cx[b] = x; cy[b] = y; cz[b] = z;
                cw[b] = fmaf(z, z, fmaf(y, y, x * x));
                s_idx[i] = j;
              }
            }
            __syncwarp();
          }
        }
        const kt_ring rg = s_ring[m];
        if (C > KT_CMAX || !rg.usable) {       // block too large for the register slots: fp64 thread path
          kt_push(pend, cs, m, lane, counters, fb_list);
          pend = 0u;
          break;
        }
        unsigned still = 0u;
        for (unsigned rem = pend; rem; rem &= rem - 1u) {
          const int gq = __ffs(rem) - 1;
          const float ax = __shfl_sync(KT_FULL, qax, gq), ay = __shfl_sync(KT_FULL, qay, gq),
                      az = __shfl_sync(KT_FULL, qaz, gq);
          // + dv: every v >= 0 (the self pair evaluates to a few ulps around 0); conservative for the bound
          const float off = fmaf(__shfl_sync(KT_FULL, qk, gq), rg.sc, rg.dv);
          __syncwarp();
          s_hist[lane] = 0u;
          s_hist[lane + 32] = 0u;
          __syncwarp();
          float v[KT_NB];
#pragma unroll
          for (int b = 0; b < KT_NB; ++b) {
            v[b] = 3.0e38f;
            if (b < nb) {
              float s = fmaf(ax, cx[b], cw[b]);
              s = fmaf(ay, cy[b], s);
              s = fmaf(az, cz[b], s);
              v[b] = fmaf(s, rg.sc, off);
              if (v[b] < 63.f) atomicAdd(&s_hist[(int)v[b]], 1u);
            }
          }
          __syncwarp();
          const uint2 hh = *(const uint2*)&s_hist[2 * lane];
          const int hsum = (int)(hh.x + hh.y);
          int inc = hsum;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(KT_FULL, inc, d);
            if (lane >= d) inc += t;
          }
          const int n_in = __shfl_sync(KT_FULL, inc, 31);
          float thr = 63.f;
          int t_take = 0, cnt1 = 0;
          const bool rl = n_in < k;
          if (rl) {
            if (!rg.rlim) {               // the block does not reach far enough: next ring
              still |= 1u << gq;
              continue;
            }
          } else {
            const unsigned mk = __ballot_sync(KT_FULL, inc >= k);
            const int L = __ffs(mk) - 1;
            const int e = __shfl_sync(KT_FULL, inc - hsum, L), a = __shfl_sync(KT_FULL, (int)hh.x, L),
                      bb = __shfl_sync(KT_FULL, (int)hh.y, L);
            int b1, c_lo;
            if (e + a >= k) { b1 = 2 * L; c_lo = e; cnt1 = a; } else { b1 = 2 * L + 1; c_lo = e + a; cnt1 = bb; }
            t_take = k - c_lo;
            thr = (float)(cnt1 == t_take ? b1 + 1 : b1);
          }
          const bool take_all = rl || cnt1 == t_take;
          // ---- emit sweep: everything below thr is a neighbour
          int* trow = s_tile + gq * ts;
          int cnt = 0;
          float vmax = -1.f;
#pragma unroll
          for (int b = 0; b < KT_NB; ++b) {
            if (b < nb) {
              const bool in = v[b] < thr;
              const unsigned mm = __ballot_sync(KT_FULL, in);
              if (in) {
                trow[cnt + __popc(mm & lt)] = s_idx[32 * b + lane];
                vmax = fmaxf(vmax, v[b]);
              }
              cnt += __popc(mm);
            }
          }
          bool amb = false;
          if (take_all) {
            if (rl) {
              bool band = false;
#pragma unroll
              for (int b = 0; b < KT_NB; ++b) band |= (v[b] >= 63.f && v[b] < 63.f + rg.dr);
              amb = __any_sync(KT_FULL, band);
            } else {
              amb = thr - kt_warp_max(vmax) <= rg.dv;     // the nearest rejected candidate has v >= thr
            }
          } else if (cnt1 > KT_LIST) {
            amb = true;
          } else {
            // ---- the cnt1 candidates of the boundary bin: rank them, take the t_take smallest
            int pos = 0;
#pragma unroll
            for (int b = 0; b < KT_NB; ++b) {
              if (b < nb) {
                const bool mb = v[b] >= thr && v[b] < thr + 1.f;
                const unsigned mm = __ballot_sync(KT_FULL, mb);
                if (mb) {
                  const int p = pos + __popc(mm & lt);
                  if (p < KT_LIST) { s_lv[p] = v[b]; s_ls[p] = 32 * b + lane; }
                }
                pos += __popc(mm);
              }
            }
            __syncwarp();
            const bool have = lane < cnt1;
            const float lv = have ? s_lv[lane] : 3.0e38f;
            const int ls = have ? s_ls[lane] : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < cnt1; ++j) {
              const float vj = __shfl_sync(KT_FULL, lv, j);
              const int sj = __shfl_sync(KT_FULL, ls, j);
              rank += (vj < lv || (vj == lv && sj < ls)) ? 1 : 0;
            }
            const unsigned mt = __ballot_sync(KT_FULL, have && rank == t_take - 1);
            const unsigned mn = __ballot_sync(KT_FULL, have && rank == t_take);
            if (pos != cnt1 || mt == 0u || mn == 0u) {
              amb = true;       // cannot happen (the two sweeps classify identically); fp64 path if it ever does
            } else {
              const float v_t = __shfl_sync(KT_FULL, lv, __ffs(mt) - 1), v_n = __shfl_sync(KT_FULL, lv, __ffs(mn) - 1);
              amb = v_n - v_t <= rg.dv;
              const bool chosen = have && rank < t_take;
              const unsigned mc = __ballot_sync(KT_FULL, chosen);
              if (chosen) trow[cnt + __popc(mc & lt)] = s_idx[ls];
              cnt += __popc(mc);
            }
          }
          if (amb) {
            kt_push(1u << gq, cs, m, lane, counters, fb_list);
          } else {
            for (int c = cnt + lane; c < k; c += 32) trow[c] = -1;
            done |= 1u << gq;
          }
        }
        pend = still;
        if (pend) {
          if (m >= prm.mmax) {
            kt_push(pend, cs, m + 1, lane, counters, fb_list);
            pend = 0u;
          } else {
            ++m;
          }
        }
      }
      // ---- write the finished rows of this chunk: lanes = (query, column) so that short chunks still fill the warp
      __syncwarp();
      if (done) {
        const int qw = Gc <= 8 ? 8 : (Gc <= 16 ? 16 : 32);
        const int lq = lane & (qw - 1), cc = lane / qw, cstep = 32 / qw;
        const int q = cs + lq;
        if ((done >> lq) & 1u) {
          int32_t* dst = ell_idx + (int64_t)(q >> 5) * k * DC_SLICE + (q & 31);
          for (int c = cc; c < k; c += cstep) dst[(int64_t)c * DC_SLICE] = s_tile[lq * ts + c];
        }
      }
      __syncwarp();
    }
  }
}

// the queries the cell kernel could not finish, one per thread, exact fp64 selection
__global__ void __launch_bounds__(KNN_THREADS)
knn_thread_list_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                       const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, dc_grid g,
                       const int32_t* __restrict__ cell_start, int k, double r2cap, int max_ring,
                       const int2* __restrict__ fb_list, const int32_t* __restrict__ counters, int32_t* __restrict__ ell_idx) {
  __shared__ unsigned short hist[KNN_BINS][KNN_THREADS];
  const int count = counters[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const int2 e = fb_list[i];
    const int q = e.x;
    int32_t* out_j = ell_idx + (int64_t)(q >> 5) * k * DC_SLICE + (q & 31);
    const dc_point pq = dc_ld_point(Q + q);
    int c0, c1, c2;
    dc_key_coords(g, qkeys[q], c0, c1, c2);
    int cnt = 0;
    const int first = e.y < 1 ? 1 : (e.y > max_ring ? max_ring : e.y);
    knn_thread_query(P, pkeys, n, g, cell_start, pq, c0, c1, c2, k, r2cap, max_ring, first, &hist[0][threadIdx.x],
                     [&](int j, double d2) {
                       out_j[(int64_t)cnt * DC_SLICE] = j;
                       ++cnt;
                     });
    for (int c = cnt; c < k; ++c) out_j[(int64_t)c * DC_SLICE] = -1;
  }
}

struct kt_is_cell_start {
  const uint64_t* keys;
  __host__ __device__ bool operator()(int i) const { return i == 0 || keys[i] != keys[i - 1]; }
};

__global__ void knn_cells_init_kernel(int32_t* header, int32_t* ell_idx, int64_t nq, int k) {
  // header: [0] number of cells (written by the select), [4] next cell, [5] fallback count
  if (threadIdx.x == 0) { header[4] = 0; header[5] = 0; }
  // unused rows of the last slice
  const int64_t slice = (nq - 1) >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32 && slice * 32 + lane >= nq)
    for (int c = 0; c < k; ++c) ell_idx[(slice * k + c) * DC_SLICE + lane] = -1;
}

extern "C" int dc_knn_cells(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                            const dc_grid_spec* spec, const int32_t* cell_start, int k, double r, int32_t* ell_idx,
                            void* temp, size_t* temp_bytes, void* stream) {
  if (!temp_bytes) return dc_set_error(DC_ERR_ARG, "dc_knn_cells: temp_bytes is NULL");
  if (k < 1 || k > 128) return dc_set_error(DC_ERR_ARG, "dc_knn_cells: k must be in [1, 128] (dc_knn handles any k)");
  if (nq > 2147483647LL || n > 2147483647LL) return dc_set_error(DC_ERR_OVERFLOW, "dc_knn_cells: more than 2^31-1 points");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nq1 = nq > 0 ? nq : 1;
  // workspace: 64-byte header | cell starts int32[nq] | fallback list int2[nq] | cub temp
  const size_t off_tasks = 64, off_fb = off_tasks + (((size_t)nq1 * 4 + 15) / 16) * 16, off_cub = off_fb + (size_t)nq1 * 8;
  size_t cub_bytes = 0;
  cub::CountingInputIterator<int> iota(0);
  kt_is_cell_start pred{qkeys};
  DC_CUDA_CHECK(cub::DeviceSelect::If(nullptr, cub_bytes, iota, (int32_t*)nullptr, (int32_t*)nullptr, (int)nq, pred, st));
  const size_t need = off_cub + cub_bytes;
  if (!temp) { *temp_bytes = need; return DC_OK; }
  if (*temp_bytes < need) return dc_set_error(DC_ERR_ARG, "dc_knn_cells: workspace too small");
  if (nq <= 0) return DC_OK;
  dc_grid g;
  int rc = dc_make_grid(spec, &g);
  if (rc) return rc;
  int max_ring = g.d[0] > g.d[1] ? g.d[0] : g.d[1];
  max_ring = max_ring > g.d[2] ? max_ring : g.d[2];
  const double ex = g.d[0] * g.cell, ey = g.d[1] * g.cell, ez = g.d[2] * g.cell;
  double r2cap = 16.0 * (ex * ex + ey * ey + ez * ez) + 1.0;
  if (r > 0.0) {
    r2cap = r * r;
    const int rr = (int)ceil(r / g.cell);
    if (rr < max_ring) max_ring = rr;
  }
  if (max_ring < 1) max_ring = 1;

  kt_params prm;
  memset(&prm, 0, sizeof(prm));
  prm.k = k;
  prm.tile_stride = k + 1;
  prm.mmax = max_ring < KT_MMAX ? max_ring : KT_MMAX;
  const char* pop = getenv("DC_KNN_POP_X10");
  prm.pop_min = (int)((pop ? atof(pop) : 25.0) * 0.1 * k);
  for (int m = 1; m <= KT_MMAX; ++m) {
    const bool last = m >= max_ring;
    const double reach = m * g.cell * (1.0 - 1e-9);
    const double bound2 = last ? r2cap : fmin(reach * reach, r2cap);
    // offsets from the centre of the query cell: |component| <= (m + 1/2) cell for every candidate of ring m
    const double ext = (m + 0.5) * g.cell * (1.0 + 1e-6);
    const double eps = ldexp(3.0 * ext * ext, -19);      // >= 32 u R^2: rounding of the offsets, |c|^2 and the 3 FFMAs
    const double b2u = bound2 * (1.0 - 1e-6) - 8.0 * eps;
    kt_ring& rg = prm.ring[m];
    rg.usable = b2u > 0.0 && 63.0 / b2u < 1e30 && 2.0 * eps * 63.0 / b2u < 0.05;
    if (!rg.usable) continue;
    rg.sc = (float)(63.0 / b2u);
    rg.dv = (float)(2.0 * eps * 63.0 / b2u + 2e-5);
    rg.dr = (float)(63.0 * (2e-6 + 12.0 * eps / b2u) + 2.0 * rg.dv + 1e-4);
    rg.rlim = last ? 1 : 0;
  }

  char* ws = (char*)temp;
  int32_t* header = (int32_t*)ws;
  int32_t* task_start = (int32_t*)(ws + off_tasks);
  int2* fb = (int2*)(ws + off_fb);
  knn_cells_init_kernel<<<1, 32, 0, st>>>(header, ell_idx, nq, k);
  DC_LAUNCH_CHECK();
  DC_CUDA_CHECK(cub::DeviceSelect::If(ws + off_cub, cub_bytes, iota, task_start, header, (int)nq, pred, st));
  const size_t smem = (size_t)KT_WARPS * (KT_WARP_WORDS + 32 * prm.tile_stride) * sizeof(int);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    DC_CUDA_CHECK(cudaFuncSetAttribute(knn_cell_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  int per_sm = 0;
  DC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_cell_kernel, KT_THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = 148;
  DC_CUDA_CHECK(cudaGetDevice(&dev));
  DC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  knn_cell_kernel<<<sms * per_sm, KT_THREADS, smem, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g,
                                                          cell_start, prm, task_start, header, header + 4, fb, ell_idx);
  DC_LAUNCH_CHECK();
  knn_thread_list_kernel<<<sms * 4, KNN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, g, cell_start,
                                                         k, r2cap, max_ring, fb, header + 4, ell_idx);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

// ---------------------------------------------------------------------------------------------
// Order every row by (d2, original index): the reference returns distance-sorted rows.  Export path only.
// ---------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void knn_sort_rows_kernel(const dc_point* __restrict__ P, int64_t n, int k, int32_t* __restrict__ ell_idx,
                                     double* __restrict__ ell_d2, int64_t nq) {
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
    while (pos > 0 && knn_less(P, n, dd, jj, d[pos - 1], j[pos - 1])) { d[pos] = d[pos - 1]; j[pos] = j[pos - 1]; --pos; }
    d[pos] = dd;
    j[pos] = jj;
  }
  for (int c = 0; c < k; ++c) {
    pj[(int64_t)c * DC_SLICE] = c < m ? j[c] : -1;
    pd[(int64_t)c * DC_SLICE] = c < m ? d[c] : INFINITY;
  }
}

extern "C" int dc_knn_sort_rows(const void* P_, int64_t n, int k, int32_t* ell_idx, double* ell_d2, int64_t nq, void* stream) {
  const dc_point* P = (const dc_point*)P_;
  if (nq <= 0) return DC_OK;
  if (k < 1 || k > 1024) return dc_set_error(DC_ERR_ARG, "dc_knn_sort_rows: k must be in [1, 1024]");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = dc_blocks(nq, 128);
  if (k <= 32) knn_sort_rows_kernel<32><<<blocks, 128, 0, st>>>(P, n, k, ell_idx, ell_d2, nq);
  else if (k <= 128) knn_sort_rows_kernel<128><<<blocks, 128, 0, st>>>(P, n, k, ell_idx, ell_d2, nq);
  else knn_sort_rows_kernel<1024><<<blocks, 128, 0, st>>>(P, n, k, ell_idx, ell_d2, nq);
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
