// Kernel 1, kNN / kNN-within-r mode: k-nearest SELECTION on the cell-sorted map.
// Replaces cKDTree.query(k, distance_upper_bound=r) (nearest_neighbors.py:48-49).
//
// The k nearest of a query are selected, not sorted: a histogram of the squared distances finds the bin that
// holds the k-th distance, everything below that bin is a neighbour, and only the few candidates inside the
// boundary bin are ranked (by (d2, original index)).  Every d2 is computed with the identical non-fused
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

#define KNN_THREADS 64
#define KNN_WARPS (KNN_THREADS / 32)
#ifndef KNN_BLOCKS
#define KNN_BLOCKS (1024 / KNN_THREADS)   // resident blocks per SM the register allocation is held to (64 registers)
#endif
#define KNN_BINS 64

// (d2, original index) lexicographic order.  The original index (dc_point.tag) -- not the position in the cell-sorted
// map -- breaks exact ties, so the selected set does not depend on the cell size.  Tags are read for the candidates
// of the boundary bin only; 0x7fffffff stands for "no candidate".
__device__ __forceinline__ bool knn_less(double a, int ta, double b, int tb) { return a < b || (a == b && ta < tb); }

__device__ __forceinline__ int knn_bin(double d2, double scale) {
  const int b = __double2int_rz(d2 * scale);
  return b > KNN_BINS - 1 ? KNN_BINS - 1 : b;
}

// Rows of the (2 rho + 1)^2 x [c0 - rho, c0 + rho] block in increasing key order, STEP candidates per callback f(j, hi).
// The cell-table lookups of three consecutive rows are issued together: one dependent pair of loads per row left every
// row waiting for the table (L2 latency once the block is larger than a few cells: the sparse queries that grow to 8+
// rings walk hundreds of mostly empty rows).
template <int STEP, typename F>
__device__ __forceinline__ void knn_rows_t(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                           const int32_t* __restrict__ cell_start, int c0, int c1, int c2, int rho, F&& f) {
  if (!cell_start) {
    for (int e2 = -rho; e2 <= rho; ++e2)
      for (int e1 = -rho; e1 <= rho; ++e1) {
        int lo, hi;
        dc_row_range(g, pkeys, n, cell_start, c0 - rho, c0 + rho, c1 + e1, c2 + e2, lo, hi);
        for (int j = lo; j < hi; j += STEP) f(j, hi);
      }
    return;
  }
  for (int e2 = -rho; e2 <= rho; ++e2) {
    for (int e1 = -rho; e1 <= rho; e1 += 3) {
      int lo0, hi0, lo1, hi1, lo2, hi2;
      dc_row_range_nb(g, cell_start, c0 - rho, c0 + rho, c1 + e1, c2 + e2, lo0, hi0);
      dc_row_range_nb(g, cell_start, c0 - rho, c0 + rho, e1 + 1 <= rho ? c1 + e1 + 1 : -1, c2 + e2, lo1, hi1);
      dc_row_range_nb(g, cell_start, c0 - rho, c0 + rho, e1 + 2 <= rho ? c1 + e1 + 2 : -1, c2 + e2, lo2, hi2);
      for (int u = 0; u < 3; ++u) {
        const int lo = u == 0 ? lo0 : (u == 1 ? lo1 : lo2), hi = u == 0 ? hi0 : (u == 1 ? hi1 : hi2);
        for (int j = lo; j < hi; j += STEP) f(j, hi);
      }
    }
  }
}

template <typename F>
__device__ __forceinline__ void knn_scan(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                         const int32_t* __restrict__ cell_start, const dc_point* __restrict__ P,
                                         const dc_point& pq, int c0, int c1, int c2, int rho, F&& f) {
  knn_rows_t<4>(g, pkeys, n, cell_start, c0, c1, c2, rho, [&](int j, int hi) {
    {
      {
      // four independent candidate loads in flight per thread (the loop is latency bound otherwise; a software
      // pipeline with eight in flight cost registers / occupancy and was 30 % slower).  The tail of a row goes
      // through the same four-wide body with clamped addresses: a one-at-a-time tail loop exposed a full load
      // latency per candidate on up to three candidates of every row.
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
  });
}

// Distance from a query to the nearest face of its own cell (key axes), a hair less.  A block of rho rings of cells is
// guaranteed to cover the ball of radius rho * cell + face around the QUERY (rho * cell around any point of its cell):
// fewer queries need a second ring.  A query clamped into a border cell, or sitting in a cell of a stacked band (whose
// index is not floor(coordinate / cell)), comes out negative on that axis and gets 0.
__device__ __forceinline__ double knn_face(const dc_grid& g, const dc_point& pq, int c0, int c1, int c2) {
  const double a0 = g.ax[0] == 0 ? pq.x : (g.ax[0] == 1 ? pq.y : pq.z);
  const double a1 = g.ax[1] == 0 ? pq.x : (g.ax[1] == 1 ? pq.y : pq.z);
  const double a2 = g.ax[2] == 0 ? pq.x : (g.ax[2] == 1 ? pq.y : pq.z);
  const double f0 = (a0 - g.org[0]) * g.inv_cell - (double)c0, f1 = (a1 - g.org[1]) * g.inv_cell - (double)c1,
               f2 = (a2 - g.org[2]) * g.inv_cell - (double)c2;
  const double m = fmin(fmin(fmin(f0, 1.0 - f0), fmin(f1, 1.0 - f1)), fmin(f2, 1.0 - f2));
  return fmax(m - 1e-6, 0.0) * g.cell;
}

// ---------------------------------------------------------------------------------------------
// Thread path: one query per thread.  h = this thread's private histogram column (stride KNN_THREADS).
// The ring of cells grows until k candidates lie inside the radius the block is guaranteed to cover; a second
// histogram level splits the boundary bin when it holds more than 8 candidates.
// ---------------------------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ bool knn_thread_query(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
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
    const double reach = rho * slack_cell + knn_face(g, pq, c0, c1, c2);
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
    return true;
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
            // the histogram is dead by now: entry nb lives in this thread's counters 8 nb .. 8 nb + 7 (four 16-bit
            // pieces of d2, two of j, two of the original index = tie-break key): 8 entries fill the 64 counters
            // exactly, which keeps the block at 16 KB of shared memory
            // (the tag is re-read here, for the <= 8 boundary candidates of a query, instead of being carried through
            // the scan for every candidate: that cost the scan loop four live registers and 4 % of the kernel)
            const int tag = (int)P[j].tag;
            const unsigned long long u = (unsigned long long)__double_as_longlong(d2);
            unsigned short* e = h + 8 * nb * KNN_THREADS;
            e[0] = (unsigned short)u;
            e[KNN_THREADS] = (unsigned short)(u >> 16);
            e[2 * KNN_THREADS] = (unsigned short)(u >> 32);
            e[3 * KNN_THREADS] = (unsigned short)(u >> 48);
            e[4 * KNN_THREADS] = (unsigned short)j;
            e[5 * KNN_THREADS] = (unsigned short)((unsigned int)j >> 16);
            e[6 * KNN_THREADS] = (unsigned short)tag;
            e[7 * KNN_THREADS] = (unsigned short)((unsigned int)tag >> 16);
            ++nb;
          }
        }
      }
    }
  });
  if (take_all) return true;
  if (use_list) {
    double bd[8];
    int bj[8], bt[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned short* e = h + 8 * i * KNN_THREADS;
      const unsigned long long u = (unsigned long long)e[0] | ((unsigned long long)e[KNN_THREADS] << 16) |
                                   ((unsigned long long)e[2 * KNN_THREADS] << 32) | ((unsigned long long)e[3 * KNN_THREADS] << 48);
      const int j = (int)((unsigned int)e[4 * KNN_THREADS] | ((unsigned int)e[5 * KNN_THREADS] << 16));
      bd[i] = i < nb ? __longlong_as_double((long long)u) : INFINITY;
      bj[i] = i < nb ? j : 0x7fffffff;
      bt[i] = i < nb ? (int)((unsigned int)e[6 * KNN_THREADS] | ((unsigned int)e[7 * KNN_THREADS] << 16)) : 0x7fffffff;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      unsigned int rank = 0u;          // candidates of the bin that precede candidate i in (d2, original index) order
#pragma unroll
      for (int m = 0; m < 8; ++m) rank += knn_less(bd[m], bt[m], bd[i], bt[i]) ? 1u : 0u;
      if (i < nb && rank < t) emit(bj[i], bd[i]);
    }
    return true;
  }
  // more than 8 candidates share the boundary sub-bin (exact ties / duplicates): repeated minimum
  // selection in (d2, original index) order -- O(t * candidates), rare
  double last_d = -1.0;
  int last_t = -1;
  for (unsigned int s_ = 0; s_ < t; ++s_) {
    double best_d = INFINITY;
    int best_j = 0x7fffffff, best_t = 0x7fffffff;
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
          if (bb == b2) {
            const int tag = (int)P[j].tag;
            if (knn_less(last_d, last_t, d2, tag) && knn_less(d2, tag, best_d, best_t)) {
              best_d = d2;
              best_j = j;
              best_t = tag;
            }
          }
        }
      }
    });
    emit(best_j, best_d);
    last_d = best_d;
    last_t = best_t;
  }
  return true;
}

__global__ void __launch_bounds__(KNN_THREADS, KNN_BLOCKS)
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
#define KT_WARPS 1                 // one warp per block: everything derived from blockIdx / shared memory at uniform
                                   // addresses is warp-uniform FOR THE COMPILER, so the warp collectives need no
                                   // convergence wrappers (BSSY / WARPSYNC / ENDCOLLECTIVE around every shuffle)
#define KT_THREADS (KT_WARPS * 32)
#define KT_NB 8                    // register slots per lane: blocks of up to 256 candidates
#define KT_CMAX (KT_NB * 32)
#define KT_MMAX 16                 // rings the cell kernel grows to (r / cell in practice)
#define KT_ROWCAP 96               // non-empty rows of a block (more: fp64 path)
#define KT_TAIL 8                  // candidates of the boundary bin kept per query (more: fp64 path)
#define KT_HSTRIDE 67              // words per histogram row: bins 0..62, dummy 63, "below" 64; odd: lanes = queries read conflict-free
#define KT_MIN_BLOCKS 16           // resident warps per SM the register allocation is held to
#define KT_GRAB 8                  // cells fetched per atomic
#define KT_LIST_WORDS (3 * KT_NB * 32)   // lane-private emit lists (front idx, tail idx, tail z); aliases the staging table
#define KT_WARP_FIXED (KT_LIST_WORDS + KT_ROWCAP + 100 + 128 + 32 * KT_TAIL + 8)

struct kt_ring {
  float sc;    // bins per unit of d2: 63 / (usable bound)
  float dv;    // fp32 classification uncertainty in bins (two candidates closer than this are re-ranked in fp64)
  float dr;    // band above bin 62 in which `d2 < r^2` cannot be decided in fp32
  int rlim;    // ring covers r (or the whole grid): fewer than k candidates in range is final
  int usable;
};
struct kt_params {
  kt_ring ring[KT_MMAX + 1];
  int mmax, k, pop_min, tile_stride, warp_words;
};

#define KT_FULL 0xffffffffu

// Non-empty rows of the block of ring m around cell (c0, c1, c2): s_rlo[i] = first sorted position, s_rpre[i] =
// candidates before row i; returns the candidate count.
__device__ __forceinline__ int kt_rows(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                       const int32_t* __restrict__ cell_start, int c0, int c1, int c2, int m, int lane,
                                       int* s_rlo, int* s_rpre, int* s_ctl) {
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
      if (pos < KT_ROWCAP) {
        s_rlo[pos] = lo;
        s_rpre[pos] = base_cnt + inc - cnt;
      }
    }
    base_rows += __popc(ne);
    base_cnt += __shfl_sync(KT_FULL, inc, 31);
  }
  if (lane == 0) { s_rpre[base_rows < KT_ROWCAP ? base_rows : KT_ROWCAP] = base_cnt; s_ctl[1] = base_cnt; s_ctl[2] = base_rows; }
  __syncwarp();
  return s_ctl[1];      // read back from a uniform shared address: a warp-uniform value as far as the compiler can tell
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

// ---- inner loops.  The block's candidates sit in register slots 0 .. nb-1 of every lane; the slot loops are Duff's
// devices: ONE straight-line copy of the eight slot bodies, entered at slot nb-1 (a switch with fall-through on the
// warp-uniform nb), so that short blocks do not pay for empty slots, the independent FFMA chains of the slots overlap,
// and the hot code stays small (with separate unrolled variants per block size the kernel was instruction-cache bound:
// 8.7 no-instruction stall cycles per issue against 0.4).

// phase 1 slot: histogram of z = zsc * (|c|^2 - 2 q.c) + zoff.  Every lane issues its atomic (ptxas turns a predicated
// shared atomic into a branch with convergence bookkeeping): out-of-range candidates count in the dummy word 63; in a
// refine pass (second level inside one bin) candidates below the bin count in word 64.
#define KT_HIST_SLOT(b)                                                                         \
  {                                                                                             \
    float s_ = fmaf(ax, cx[b], cw[b]);                                                          \
    s_ = fmaf(ay, cy[b], s_);                                                                   \
    s_ = fmaf(az, cz[b], s_);                                                                   \
    const float z_ = fmaf(s_, zsc, zoff);                                                       \
    unsigned bin_ = min(__float2uint_rz(z_), 63u);                                              \
    if (REFINE) bin_ = z_ < 0.f ? 64u : bin_;                                                   \
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(hbase + 4u * bin_), "r"(1u));             \
  }

template <bool REFINE>
__device__ __forceinline__ void kt_hist(int nb, const float (&cx)[KT_NB], const float (&cy)[KT_NB], const float (&cz)[KT_NB],
                                        const float (&cw)[KT_NB], float ax, float ay, float az, float zsc, float zoff,
                                        unsigned hbase) {
  // slot pairs: two independent chains per basic block (a case label ends the scheduler's window)
  switch ((nb + 1) >> 1) {
    case 4: KT_HIST_SLOT(7) KT_HIST_SLOT(6)
    case 3: KT_HIST_SLOT(5) KT_HIST_SLOT(4)
    case 2: KT_HIST_SLOT(3) KT_HIST_SLOT(2)
    case 1: KT_HIST_SLOT(1) KT_HIST_SLOT(0)
    default: break;
  }
}

// phase 3 slot: z < thr -> the lane's private front list; thr <= z < thr_hi -> its private tail list (index and z).
// No collectives and no atomics here; the lists are compacted into the query's row once per query.
#define KT_EMIT_SLOT(b)                                                                         \
  {                                                                                             \
    float s_ = fmaf(ax, cx[b], cw[b]);                                                          \
    s_ = fmaf(ay, cy[b], s_);                                                                   \
    s_ = fmaf(az, cz[b], s_);                                                                   \
    const float z_ = fmaf(s_, zsc, zoff);                                                       \
    if (z_ < thr) {                                                                             \
      *pf = cj[b];                                                                              \
      pf += 32;                                                                                 \
    } else if (z_ < thr_hi) {                                                                   \
      pt[0] = cj[b];                                                                            \
      pt[KT_NB * 32] = __float_as_int(z_);                                                      \
      pt += 32;                                                                                 \
    }                                                                                           \
  }

// phase 2 (lanes = queries x histogram segments): first bin b with base + sum(bins <= b) >= k for every query row that
// `want` selects; results {n_in, b, count before b, count of b} -> s_res[4 qi ..].  base = word 63 of the row when
// `refined` (candidates below the refined bin), else 0; bins 0 .. 62 are summed.
__device__ __forceinline__ void kt_scan(const int* s_u, int* s_res, int Gc, int k, int lpq, int sub, int qi, int seg,
                                        bool want, bool refined) {
  int S = 0, base = 0;
  const int* hrow = s_u + qi * KT_HSTRIDE + sub * seg;
  const int nbin = sub * seg + seg > 63 ? 63 - sub * seg : seg;       // the last segment stops at bin 62
  const bool act = want && qi < Gc;
  if (act) {
    for (int b = 0; b < nbin; ++b) S += hrow[b];
    if (refined) base = s_u[qi * KT_HSTRIDE + 64];
  }
  int inc = S;
  for (int d = 1; d < lpq; d <<= 1) {
    const int t = __shfl_up_sync(KT_FULL, inc, d, lpq);
    if (sub >= d) inc += t;
  }
  inc += base;
  const int n_in = __shfl_sync(KT_FULL, inc, lpq - 1, lpq);
  const int E = inc - S;
  if (act) {
    if (n_in < k) {
      if (sub == 0) { s_res[4 * qi] = n_in; s_res[4 * qi + 1] = 63; s_res[4 * qi + 2] = n_in; s_res[4 * qi + 3] = 0; }
    } else if (E < k && E + S >= k) {
      int acc = E, b = 0, h = hrow[0];
      while (acc + h < k) { acc += h; h = hrow[++b]; }
      s_res[4 * qi] = n_in; s_res[4 * qi + 1] = sub * seg + b; s_res[4 * qi + 2] = acc; s_res[4 * qi + 3] = h;
    } else if (E >= k && sub == 0) {
      // (refined rows only) k candidates already lie below the refined bin: cannot happen, flagged for the fp64 path
      s_res[4 * qi] = n_in; s_res[4 * qi + 1] = 0; s_res[4 * qi + 2] = 0; s_res[4 * qi + 3] = 1 << 20;
    }
  }
}

#define KT_FB_AMBIGUOUS 1     // gap between the last selected and the first rejected candidate below the fp32 uncertainty
#define KT_FB_CROWDED 2       // more than KT_TAIL candidates share the (refined) boundary bin: exact ties / duplicates
#define KT_FB_RING 3          // needs more rings than KT_MMAX, or the ring table entry is unusable

__global__ void __launch_bounds__(KT_THREADS, KT_MIN_BLOCKS)
knn_cell_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
                const int32_t* __restrict__ cell_start, kt_params prm, const int32_t* __restrict__ task_start,
                const int32_t* __restrict__ n_tasks_p, int32_t* counters, int2* fb_list, int32_t* __restrict__ ell_idx) {
  extern __shared__ __align__(16) int kt_smem[];
  const int lane = threadIdx.x;
  const int k = prm.k, ts = prm.tile_stride;
  int* s_list = kt_smem;                              // lane-private emit lists; during staging: sorted positions
  int* s_rlo = s_list + KT_LIST_WORDS;                // row table of the block
  int* s_rpre = s_rlo + KT_ROWCAP;
  int* s_res = s_rpre + 100;                          // scan results [32][4]; during emit: row counters [32][2]
  float* s_sv = (float*)(s_res + 128);                // z of the boundary-bin candidates [32][KT_TAIL]
  int* s_ctl = (int*)(s_sv + 32 * KT_TAIL);           // control words published by one lane, read by all
  int* s_u = s_ctl + 8;                               // histograms [32][KT_HSTRIDE], later the output tile [32][ts]
  // ring table in shared memory: indexing the kernel parameter with a run-time ring would copy it to local memory
  __shared__ kt_ring s_ring[KT_MMAX + 1];
  for (int i = lane; i <= KT_MMAX; i += 32) s_ring[i] = prm.ring[i < KT_MMAX ? i : KT_MMAX];
  __syncthreads();
  const int n_tasks = *n_tasks_p;
  const unsigned lt = (1u << lane) - 1u;
  int task_next = 0, task_end = 0;
  for (;;) {
    if (task_next >= task_end) {
      __syncwarp();
      if (lane == 0) s_ctl[0] = atomicAdd(&counters[0], KT_GRAB);
      __syncwarp();
      const int t0 = s_ctl[0];
      if (t0 >= n_tasks) break;
      task_next = t0;
      task_end = t0 + KT_GRAB < n_tasks ? t0 + KT_GRAB : n_tasks;
    }
    const int task = task_next++;
    const int s0 = task_start[task];
    const int s1 = task + 1 < n_tasks ? task_start[task + 1] : (int)nq;
    // cell of the task from its first query (the very computation that produced the sort key, without the 64-bit
    // divisions of decoding the key)
    int c0, c1, c2;
    {
      const dc_point p0 = dc_ld_point(Q + s0);
      const double a0 = g.ax[0] == 0 ? p0.x : (g.ax[0] == 1 ? p0.y : p0.z), a1 = g.ax[1] == 0 ? p0.x : (g.ax[1] == 1 ? p0.y : p0.z),
                   a2 = g.ax[2] == 0 ? p0.x : (g.ax[2] == 1 ? p0.y : p0.z);
      // identical arithmetic to dc_cell_coords (which indexes p[ax[a]] and would put p in local memory here)
      c0 = dc_clampi((int)floor((a0 - g.org[0]) * g.inv_cell), 0, g.d[0] - 1);
      c1 = dc_clampi((int)floor((a1 - g.org[1]) * g.inv_cell), 0, g.d[1] - 1);
      c2 = dc_clampi((int)floor((a2 - g.org[2]) * g.inv_cell), 0, g.d[2] - 1);
    }
    // centre of the query cell = origin of the fp32 offsets (xyz order)
    const double o0 = g.org[0] + ((double)c0 + 0.5) * g.cell, o1 = g.org[1] + ((double)c1 + 0.5) * g.cell,
                 o2 = g.org[2] + ((double)c2 + 0.5) * g.cell;
    const double ox = g.ax[0] == 0 ? o0 : (g.ax[1] == 0 ? o1 : o2);
    const double oy = g.ax[0] == 1 ? o0 : (g.ax[1] == 1 ? o1 : o2);
    const double oz = g.ax[0] == 2 ? o0 : (g.ax[1] == 2 ? o1 : o2);

    float cx[KT_NB], cy[KT_NB], cz[KT_NB], cw[KT_NB];
    int cj[KT_NB];
    int rows_m = 0, staged_sb = -1, m_start = 0, C = 0;

    for (int cs = s0; cs < s1; cs += 32) {
      const int Gc = s1 - cs < 32 ? s1 - cs : 32;
      unsigned pend = Gc == 32 ? KT_FULL : ((1u << Gc) - 1u);
      float qax = 0.f, qay = 0.f, qaz = 0.f, qk = 0.f;
      if (lane < Gc) {
        const dc_point pq = dc_ld_point(Q + cs + lane);
        const float x = (float)(pq.x - ox), y = (float)(pq.y - oy), z = (float)(pq.z - oz);
        qax = -2.f * x; qay = -2.f * y; qaz = -2.f * z;
        qk = fmaf(z, z, fmaf(y, y, x * x));
      }
      // lanes <-> (query, segment of its histogram) for the scans: qw = pow2 >= Gc queries x lpq lanes each
      const int qw = Gc <= 1 ? 1 : (1 << (32 - __clz(Gc - 1)));
      const int lpq = 32 / qw, sub = lane & (lpq - 1), qi = lane / lpq, seg = 64 / lpq;
      int m = m_start ? m_start : 1;
      while (pend) {
        if (rows_m != m) {
          C = kt_rows(g, pkeys, n, cell_start, c0, c1, c2, m, lane, s_rlo, s_rpre, s_ctl);
          if (m_start == 0) {
            // first visit of this cell: skip rings whose block cannot hold k neighbours with some margin
            while (C < prm.pop_min && m < prm.mmax) {
              ++m;
              C = kt_rows(g, pkeys, n, cell_start, c0, c1, c2, m, lane, s_rlo, s_rpre, s_ctl);
            }
            m_start = m;
          }
          rows_m = m;
          staged_sb = -1;
        }
        const kt_ring rg = s_ring[m];
        const int n_rows = s_ctl[2];
        if (!rg.usable || n_rows > KT_ROWCAP) {
          kt_push(pend, cs, (m < 255 ? m : 255) | (KT_FB_RING << 8), lane, counters, fb_list);
          break;
        }
        const int nsb = (C + KT_CMAX - 1) / KT_CMAX;
        // per-lane state of query `lane`.  Transform: z = zsc * s + zoff (level 1: z = v = 63 d2 / bound; + dv keeps
        // every v >= 0: the self pair evaluates to a few ulps around 0; conservative for the bound)
        float q_zsc = rg.sc, q_zoff = fmaf(qk, rg.sc, rg.dv), q_dv = rg.dv;
        float q_thr = 63.f, q_hi = 63.f;
        int q_clo = 0, q_cnt1 = 0, q_t = 0;
        bool q_rl = false;
        int q_status = 4;            // 0 ready, 1 needs a larger ring, 2 fp64 path, 3 refine the boundary bin, 4 not in this round
        unsigned active = pend;      // queries of the current pass
        unsigned ready = 0u, crowded = 0u, still = 0u;
        // pass 0: level-1 histograms + scan; pass 1: second histogram level for crowded boundary bins (usually
        // skipped); pass 2: emit.  ONE staging site and one loop nest for all passes keeps the code small.
        for (int pass = 0; pass < 3; ++pass) {
          if (pass < 2) {
            if (active == 0u) continue;
            __syncwarp();
            for (int i = lane; i < Gc * KT_HSTRIDE; i += 32) s_u[i] = 0;
          } else {
            active = ready;
            __syncwarp();
            s_res[lane] = 0;             // row counters: [2 g] front entries, [2 g + 1] boundary-bin candidates
            s_res[lane + 32] = 0;
          }
          __syncwarp();
          for (int t = 0; t < nsb; ++t) {
            const int sb = pass == 2 ? nsb - 1 - t : t;       // the emit pass starts with the block staged last
            if (staged_sb != sb) {
              // ---- stage candidates [256 sb, 256 sb + 256) of the block into the register slots
              __syncwarp();
              const int w0 = KT_CMAX * sb, w1 = C < w0 + KT_CMAX ? C : w0 + KT_CMAX;
              for (int r = 0; r < n_rows; ++r) {
                const int pre = s_rpre[r], nxt = s_rpre[r + 1];
                if (nxt <= w0 || pre >= w1) continue;
                const int a = pre > w0 ? pre : w0, e = nxt < w1 ? nxt : w1, lo = s_rlo[r];
                for (int i = a + lane; i < e; i += 32) s_list[i - w0] = lo + (i - pre);
              }
              __syncwarp();
#pragma unroll
              for (int b = 0; b < KT_NB; ++b) {
                cx[b] = 0.f; cy[b] = 0.f; cz[b] = 0.f; cw[b] = 3.0e38f; cj[b] = -1;
                if (w0 + 32 * b + lane < w1) {
                  const int j = s_list[32 * b + lane];
                  const dc_point p = dc_ld_point(P + j);
                  const float x = (float)(p.x - ox), y = (float)(p.y - oy), z = (float)(p.z - oz);
                  cx[b] = x; cy[b] = y; cz[b] = z;
                  cw[b] = fmaf(z, z, fmaf(y, y, x * x));
                  cj[b] = j;
                }
              }
              __syncwarp();
              staged_sb = sb;
            }
            const int left = C - KT_CMAX * sb;
            const int nb = left >= KT_CMAX ? KT_NB : (left + 31) >> 5;
            for (unsigned rem = active; rem; rem &= rem - 1u) {
              const int gq = __ffs(rem) - 1;
              const float ax = __shfl_sync(KT_FULL, qax, gq), ay = __shfl_sync(KT_FULL, qay, gq),
                          az = __shfl_sync(KT_FULL, qaz, gq), zsc = __shfl_sync(KT_FULL, q_zsc, gq),
                          zoff = __shfl_sync(KT_FULL, q_zoff, gq);
              if (pass == 0) {
                kt_hist<false>(nb, cx, cy, cz, cw, ax, ay, az, zsc, zoff, (unsigned)__cvta_generic_to_shared(s_u + gq * KT_HSTRIDE));
              } else if (pass == 1) {
                kt_hist<true>(nb, cx, cy, cz, cw, ax, ay, az, zsc, zoff, (unsigned)__cvta_generic_to_shared(s_u + gq * KT_HSTRIDE));
              } else {
                const float thr = __shfl_sync(KT_FULL, q_thr, gq), thr_hi = __shfl_sync(KT_FULL, q_hi, gq);
                const int clo = __shfl_sync(KT_FULL, q_clo, gq);
                int* const pf0 = s_list + lane;
                int* const pt0 = s_list + KT_NB * 32 + lane;
                int* pf = pf0;
                int* pt = pt0;
                switch ((nb + 1) >> 1) {
                  case 4: KT_EMIT_SLOT(7) KT_EMIT_SLOT(6)
                  case 3: KT_EMIT_SLOT(5) KT_EMIT_SLOT(4)
                  case 2: KT_EMIT_SLOT(3) KT_EMIT_SLOT(2)
                  case 1: KT_EMIT_SLOT(1) KT_EMIT_SLOT(0)
                  default: break;
                }
                // compact the private lists into the query's row: front entries by a warp scan of the counts,
                // the (rare) boundary-bin candidates by a counter
                const int cfl = (int)(pf - pf0) >> 5, ctl = (int)(pt - pt0) >> 5;
                int inc = cfl;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                  const int u = __shfl_up_sync(KT_FULL, inc, d);
                  if (lane >= d) inc += u;
                }
                const int total = __shfl_sync(KT_FULL, inc, 31);
                int* trow = s_u + gq * ts;
                const int base = s_res[2 * gq];
                int* dst = trow + base + inc - cfl;
                for (int i = 0; i < cfl; ++i) dst[i] = pf0[32 * i];
                if (ctl > 0) {
                  for (int i = 0; i < ctl; ++i) {
                    const int p = atomicAdd(&s_res[2 * gq + 1], 1);
                    if (p < KT_TAIL) { trow[clo + p] = pt0[32 * i]; s_sv[gq * KT_TAIL + p] = __int_as_float(pt0[32 * i + KT_NB * 32]); }
                  }
                }
                __syncwarp();
                if (lane == 0) s_res[2 * gq] = base + total;
              }
            }
          }
          __syncwarp();
          if (pass == 0) {
            // ---- phase 2: bin of the k-th distance
            kt_scan(s_u, s_res, Gc, k, lpq, sub, qi, seg, (pend >> qi) & 1u, false);
            __syncwarp();
            if (lane < Gc && ((pend >> lane) & 1u)) {
              const int n_in = s_res[4 * lane], b1 = s_res[4 * lane + 1];
              q_clo = s_res[4 * lane + 2];
              q_cnt1 = s_res[4 * lane + 3];
              if (n_in < k) {
                q_rl = true;
                q_status = rg.rlim ? 0 : 1;
                q_thr = 63.f; q_hi = 63.f + rg.dr;           // tail = the band in which d2 < r^2 is undecidable in fp32
              } else {
                q_t = k - q_clo;
                q_thr = (float)b1; q_hi = (float)(b1 + 1);   // tail = the boundary bin
                q_status = 0;
                if (q_cnt1 > KT_TAIL) {
                  // crowded boundary bin (dense cells: d_k << reach): 63 sub-bins inside it, z = 63 (v - b1)
                  q_status = 3;
                  q_zsc = rg.sc * 63.f;
                  q_zoff = fmaf(q_zoff, 63.f, -63.f * (float)b1);
                  q_dv = fmaf(rg.dv, 63.f, 1e-3f);
                }
              }
            }
            active = __ballot_sync(KT_FULL, q_status == 3);
          } else if (pass == 1) {
            kt_scan(s_u, s_res, Gc, k, lpq, sub, qi, seg, (active >> qi) & 1u, true);
            __syncwarp();
            if (q_status == 3) {
              const int n_in = s_res[4 * lane], b2 = s_res[4 * lane + 1];
              q_clo = s_res[4 * lane + 2];
              q_cnt1 = s_res[4 * lane + 3];
              q_t = k - q_clo;
              q_thr = (float)b2; q_hi = (float)(b2 + 1);
              q_status = (n_in < k || q_cnt1 > KT_TAIL) ? 2 : 0;
            }
          }
          if (pass < 2) {
            ready = __ballot_sync(KT_FULL, q_status == 0);
            crowded = __ballot_sync(KT_FULL, q_status == 2);
            still = __ballot_sync(KT_FULL, q_status == 1);
          }
        }
        __syncwarp();
        // ---- phase 4 (lanes = queries): the t smallest of the boundary bin, ambiguity test, padding
        bool amb = false;
        if (q_status == 0) {
          const int cf = s_res[2 * lane], ct = s_res[2 * lane + 1];
          int* trow = s_u + lane * ts;
          const float* sv = s_sv + lane * KT_TAIL;
          int cnt_final = q_clo;
          if (q_rl) {
            amb = ct > 0 || cf != q_clo;
          } else if (ct != q_cnt1 || cf != q_clo) {
            amb = true;          // cannot happen (all sweeps classify identically); fp64 path if it ever does
          } else if (q_cnt1 == q_t) {
            float vm = -1.f;
            for (int i = 0; i < ct; ++i) vm = fmaxf(vm, sv[i]);
            amb = q_hi - vm <= q_dv;                    // the nearest rejected candidate has z >= the next bin edge
            cnt_final = k;
          } else {
            // rank by (z, arrival); keep rank < t.  v_t / v_n: last kept / first rejected value
            float v_t = -1.f, v_n = 3.0e38f;
            unsigned keep = 0u;
            for (int i = 0; i < ct; ++i) {
              const float vi = sv[i];
              int rank = 0;
              for (int j = 0; j < ct; ++j) {
                const float vj = sv[j];
                rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0;
              }
              if (rank < q_t) { keep |= 1u << i; v_t = fmaxf(v_t, vi); } else { v_n = fminf(v_n, vi); }
            }
            amb = v_n - v_t <= q_dv;
            int w = q_clo;
            for (int i = 0; i < ct; ++i)
              if ((keep >> i) & 1u) trow[w++] = trow[q_clo + i];
            cnt_final = k;
          }
          for (int c = cnt_final; c < k; ++c) trow[c] = -1;
        }
        const unsigned amb_m = __ballot_sync(KT_FULL, amb);
        const unsigned done = ready & ~amb_m;
        kt_push(amb_m, cs, (m < 255 ? m : 255) | (KT_FB_AMBIGUOUS << 8), lane, counters, fb_list);
        kt_push(crowded, cs, (m < 255 ? m : 255) | (KT_FB_CROWDED << 8), lane, counters, fb_list);
        __syncwarp();
        // ---- write the finished rows: lanes = (query, column) so that short chunks still fill the warp
        if (done) {
          const int lq = lane & (qw - 1), cc = lane / qw, cstep = 32 / qw;
          const int q = cs + lq;
          if ((done >> lq) & 1u) {
            int32_t* dst = ell_idx + (int64_t)(q >> 5) * k * DC_SLICE + (q & 31);
            for (int c = cc; c < k; c += cstep) dst[(int64_t)c * DC_SLICE] = s_u[lq * ts + c];
          }
        }
        __syncwarp();
        pend = still;
        if (pend) {
          if (m >= prm.mmax) {
            kt_push(pend, cs, (m + 1 < 255 ? m + 1 : 255) | (KT_FB_RING << 8), lane, counters, fb_list);
            pend = 0u;
          } else {
            m = m < 4 ? m + 1 : (m + (m >> 1) < prm.mmax ? m + (m >> 1) : prm.mmax);
          }
        }
      }
    }
  }
}

// the queries the cell kernel could not finish, one per thread, exact fp64 selection
__global__ void __launch_bounds__(KNN_THREADS, KNN_BLOCKS)
knn_thread_list_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                       const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, dc_grid g,
                       const int32_t* __restrict__ cell_start, int k, double r2cap, int max_ring,
                       const int2* __restrict__ fb_list, const int32_t* __restrict__ counters, int32_t* __restrict__ ell_idx) {
  __shared__ unsigned short hist[KNN_BINS][KNN_THREADS];
  const int count = counters[1];
  // entry i goes to lane i / n_warps of warp i % n_warps: a short list of heavy queries is spread over all warps
  // instead of filling the first few with 32 heavy queries each
  const int n_warps = gridDim.x * (blockDim.x >> 5);
  const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int i = warp + (threadIdx.x & 31) * n_warps; i < count; i += 32 * n_warps) {
    const int2 e = fb_list[i];
    const int q = e.x;
    int32_t* out_j = ell_idx + (int64_t)(q >> 5) * k * DC_SLICE + (q & 31);
    const dc_point pq = dc_ld_point(Q + q);
    int c0, c1, c2;
    dc_key_coords(g, qkeys[q], c0, c1, c2);
    int cnt = 0;
    const int ring = e.y & 0xff;        // (bits 8..: why the cell kernel gave up, for statistics)
    const int first = ring < 1 ? 1 : (ring > max_ring ? max_ring : ring);
    knn_thread_query(P, pkeys, n, g, cell_start, pq, c0, c1, c2, k, r2cap, max_ring, first, &hist[0][threadIdx.x],
                     [&](int j, double d2) {
                       out_j[(int64_t)cnt * DC_SLICE] = j;
                       ++cnt;
                     });
    for (int c = cnt; c < k; ++c) out_j[(int64_t)c * DC_SLICE] = -1;
  }
}

// ---------------------------------------------------------------------------------------------
// Recorded path (dc_knn_recorded): one query per thread like dc_knn, but ONE distance pass.
//
// The histogram pass of dc_knn already visits every candidate of the block and knows the bin of each one inside the
// bound; dc_knn then walks all rows again (loads, fp64 distances, bins) to emit.  Here the first pass records the bin
// of every visited candidate -- one byte, 255 = outside the bound, four candidates per 32-bit word, one word per
// iteration of the four-wide scan loop -- in a thread-private array (local memory), and the later passes walk the same
// rows reading the record instead of the map: only the candidates OF the boundary bin (a handful) are re-read to be
// ranked by (fp64 d2, original index) with the very arithmetic of dc_knn.  The word index is the thread's own iteration
// counter: the lanes of a warp that sit in the same cell walk identical rows, so their counters agree and a warp's
// access touches as many 128-byte lines as it has cells (about three), like the candidate loads themselves.  (A record
// of (index, bin) pairs of the in-range candidates only -- shorter, but at lane-private positions, one line per lane
// and access -- was no faster than re-walking the rows.)  Same result as dc_knn bit for bit, same order inside a row.
// Iterations beyond KR_WORDS (long scans: many rings, dense cells) are not recorded and are recomputed by the later
// passes, so there is no cliff; queries with more than 8 candidates tied in the boundary sub-bin (exact ties,
// duplicates) go to a list that knn_thread_list_kernel finishes.
// ---------------------------------------------------------------------------------------------
#define KR_WORDS 256

// rows of knn_scan in the same order, one callback per four-wide iteration: f(j, hi) with j the first candidate
template <typename F>
__device__ __forceinline__ void knn_rows(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                         const int32_t* __restrict__ cell_start, int c0, int c1, int c2, int rho, F&& f) {
  knn_rows_t<4>(g, pkeys, n, cell_start, c0, c1, c2, rho, f);
}

// the same rows, sixteen candidates per callback
template <typename F>
__device__ __forceinline__ void knn_rows16(const dc_grid& g, const uint64_t* __restrict__ pkeys, int64_t n,
                                           const int32_t* __restrict__ cell_start, int c0, int c1, int c2, int rho, F&& f) {
  knn_rows_t<16>(g, pkeys, n, cell_start, c0, c1, c2, rho, f);
}

// bins of the four candidates j .. j+3 of a row ending at hi (255 = outside the bound / past the end): what the first
// pass records, recomputed for the iterations a long scan could not record
__device__ __forceinline__ unsigned int knn_word(const dc_point* __restrict__ P, const dc_point& pq, int j, int hi, double bound2,
                                              double scale1) {
  const int lastj = hi - 1;
  const int j1 = j + 1 < lastj ? j + 1 : lastj, j2 = j + 2 < lastj ? j + 2 : lastj, j3 = j + 3 < lastj ? j + 3 : lastj;
  const dc_point p0 = dc_ld_point(P + j), p1 = dc_ld_point(P + j1);
  const dc_point p2 = dc_ld_point(P + j2), p3 = dc_ld_point(P + j3);
  const double d0 = dc_dist2(p0, pq), d1 = dc_dist2(p1, pq), d2 = dc_dist2(p2, pq), d3 = dc_dist2(p3, pq);
  unsigned int w = 0xffffffffu;
  if (d0 < bound2) w ^= (unsigned int)(knn_bin(d0, scale1) ^ 255);
  if (j + 1 < hi && d1 < bound2) w ^= (unsigned int)(knn_bin(d1, scale1) ^ 255) << 8;
  if (j + 2 < hi && d2 < bound2) w ^= (unsigned int)(knn_bin(d2, scale1) ^ 255) << 16;
  if (j + 3 < hi && d3 < bound2) w ^= (unsigned int)(knn_bin(d3, scale1) ^ 255) << 24;
  return w;
}

// How the first pass is kept lean (every step timed on the 8.37 M point bench map, profiles/r2_knn_experiments.md):
//  * histogram columns at byte offset 4 lane + 2 warp: the 32 lanes of a warp sit in 32 different banks whatever their
//    bins (a column at 2 tid puts two lanes into every bank: a 2-way conflict on every counter access);
//  * no branch around the histogram update: the counter of the candidate's bin is read, incremented and stored under a
//    predicate, the record byte is selected (`if (inside) {...}` around a shared-memory read-modify-write compiles to
//    BSSY / BRA / BSYNC per candidate, with a third of the lanes inside on lidar maps).  Counters wrap instead of
//    saturating: a query with more than 65535 candidates inside the bound goes to the list knn_thread_list_kernel
//    finishes;
//  * the four candidates of a step are read at pj, pj + 1, pj + 2, pj + 3 without clamping the last ones to the end of
//    the row (one address computation): the caller provides KNN_PAD = 3 readable records behind the map, and slots
//    past the end of the row are masked;
//  * x, y, z are loaded without the tag (a 128-bit and a 64-bit load: six registers per candidate instead of the
//    eight a 256-bit load pins), which lets all four loads of a step be in flight inside the 64-register budget;
//  * the radius a block of rho rings is guaranteed to cover is measured from the QUERY (knn_face), so fewer queries
//    need a second ring;
//  * the emit pass stores with predicates and only collects the positions of the (<= 8) candidates of the boundary bin;
//    their distances and tags are read afterwards, eight independent loads with the warp converged.
// ---------------------------------------------------------------------------------------------

// sub-bin (second histogram level) of candidate j of bin b1: out of line, rare
__device__ __noinline__ int knn_sub_bin(const dc_point* __restrict__ P, int j, double qx, double qy, double qz, double scale1, int b1) {
  const dc_point pj = dc_ld_point(P + j);
  dc_point pq;
  pq.x = qx; pq.y = qy; pq.z = qz; pq.tag = 0;
  const double d2 = dc_dist2(pj, pq);
  int bb = __double2int_rz((d2 * scale1 - (double)b1) * (double)KNN_BINS);
  return bb < 0 ? 0 : (bb > KNN_BINS - 1 ? KNN_BINS - 1 : bb);
}

// x, y, z of a record without its tag
__device__ __forceinline__ dc_point knn_ld_xyz(const dc_point* p) {
  dc_point r;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  asm("ld.global.nc.f64 %0, [%1+16];" : "=d"(r.z) : "l"(p));
  r.tag = 0;
  return r;
}

// 16-bit shared-memory accesses at a 32-bit shared address (the kernel below keeps ONE opaque register with the address
// of the thread's histogram column: a generic pointer to it was re-derived -- S2R, S2UR, ULEA, ... -- at every use)
__device__ __forceinline__ unsigned int knn_lds16(unsigned int a) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void knn_sts16(unsigned int a, unsigned int v) {
  asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((unsigned short)v) : "memory");
}
#define KNN_COL (KNN_THREADS * 2u)      // bytes between two counters of a histogram column

__global__ void __launch_bounds__(KNN_THREADS, KNN_BLOCKS)
knn_record_kernel(const dc_point* __restrict__ P, const uint64_t* __restrict__ pkeys, int64_t n,
                  const dc_point* __restrict__ Q, const uint64_t* __restrict__ qkeys, int64_t nq, dc_grid g,
                  const int32_t* __restrict__ cell_start, int k, double r2cap, int max_ring,
                  int2* __restrict__ fb_list, int32_t* __restrict__ counters, int32_t* __restrict__ ell_idx) {
  static_assert(KNN_THREADS == 64, "the interleaved histogram columns are laid out for two warps per block");
  __shared__ unsigned short hist[KNN_BINS][KNN_THREADS];
  unsigned int rec[KR_WORDS];
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  if (nq <= 0 || (q >> 5) > ((nq - 1) >> 5)) return;
  int32_t* out_j = ell_idx + (q >> 5) * (int64_t)k * DC_SLICE + lane;
  int cnt = 0;
  bool fallback = false;
  if (q < nq) {
#ifdef DC_KNN_STATS
    const long long stats_t0 = clock64();
#endif
    // shared address of this thread's histogram column (counter b at hs + b * KNN_COL)
    unsigned int hs = (unsigned int)__cvta_generic_to_shared(&hist[0][2 * lane + (threadIdx.x >> 5)]);
    asm volatile("" : "+r"(hs));      // opaque: kept in a register instead of being recomputed (8 instructions) in every step
    const dc_point pq = dc_ld_point(Q + q);
    int c0, c1, c2;
    dc_key_coords(g, qkeys[q], c0, c1, c2);
    const double slack_cell = g.cell * (1.0 - 1e-9);
    const double face = knn_face(g, pq, c0, c1, c2);
    // ---- 1. ring growth + level-1 histogram + record
    int rho = 1;
    double bound2, scale1;
    unsigned int n_in;
    int n_words;
    for (;;) {
      if (rho > max_ring) rho = max_ring;
      const bool last = rho >= max_ring;
      const double reach = rho * slack_cell + face;
      bound2 = last ? r2cap : fmin(reach * reach, r2cap);
      scale1 = (double)KNN_BINS / bound2;
#pragma unroll
      for (int b = 0; b < KNN_BINS; ++b) knn_sts16(hs + b * KNN_COL, 0u);
      n_words = 0;
      unsigned int n_out = 0u;
      knn_rows(g, pkeys, n, cell_start, c0, c1, c2, rho, [&](int j, int hi) {
        const dc_point* pj = P + j;
        const dc_point p0 = knn_ld_xyz(pj), p1 = knn_ld_xyz(pj + 1), p2 = knn_ld_xyz(pj + 2), p3 = knn_ld_xyz(pj + 3);
        const double d0 = dc_dist2(p0, pq), d1 = dc_dist2(p1, pq), d2 = dc_dist2(p2, pq), d3 = dc_dist2(p3, pq);
        unsigned int w = 0u;
        auto visit = [&](double dd, bool valid, unsigned int sel) {
          const bool in = valid && dd < bound2;
          const int b = knn_bin(dd, scale1);
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .b16 v;\n\tsetp.ne.s32 p, %1, 0;\n\tld.shared.u16 v, [%0];\n\t"
              "add.u16 v, v, 1;\n\t@p st.shared.u16 [%0], v;\n\t}"
              :: "r"(hs + (unsigned int)b * KNN_COL), "r"((int)in) : "memory");
          w = __byte_perm(w, (unsigned int)(in ? b : 255), sel);
        };
        visit(d0, true, 0x3214u);
        visit(d1, j + 1 < hi, 0x3240u);
        visit(d2, j + 2 < hi, 0x3410u);
        visit(d3, j + 3 < hi, 0x4210u);
        n_out += (unsigned int)__popc(w & 0x80808080u);      // bit 7 of a record byte = outside the bound / past the row
        if (n_words < KR_WORDS) rec[n_words] = w;
        ++n_words;
      });
      n_in = 4u * (unsigned int)n_words - n_out;
      if (n_in >= (unsigned int)k || last) break;
      rho = rho < 4 ? rho + 1 : rho * 2;
    }
    if (n_in > 65535u) {
      fallback = true;                       // a 16-bit counter may have wrapped
    } else if (n_in <= (unsigned int)k) {
      // everything inside the bound is a neighbour (fewer than k exist within r / in the map)
      int it = 0;
      knn_rows(g, pkeys, n, cell_start, c0, c1, c2, rho, [&](int j, int hi) {
        const unsigned int w = it < KR_WORDS ? rec[it] : knn_word(P, pq, j, hi, bound2, scale1);
        ++it;
        if (w == 0xffffffffu) return;
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_)
          if (((w >> (8 * s_)) & 255u) != 255u) out_j[(int64_t)(cnt++) * DC_SLICE] = j + s_;
      });
    } else {
      // ---- level 1: bin of the k-th distance
      // (four counters per step: one dependent shared-memory load per bin was 3 % of the kernel; n_in > k, so the
      // scan always ends inside the histogram)
      unsigned int c_lo = 0u, cnt1 = 0u;
      int b1 = 0;
      for (; b1 < KNN_BINS; b1 += 4) {
        const unsigned int v0 = knn_lds16(hs + b1 * KNN_COL), v1 = knn_lds16(hs + (b1 + 1) * KNN_COL),
                           v2 = knn_lds16(hs + (b1 + 2) * KNN_COL), v3 = knn_lds16(hs + (b1 + 3) * KNN_COL);
        const unsigned int s1 = c_lo + v0, s2 = s1 + v1, s3 = s2 + v2, s4 = s3 + v3;
        if (s4 < (unsigned int)k) { c_lo = s4; continue; }
        if (s1 >= (unsigned int)k) { cnt1 = v0; }
        else if (s2 >= (unsigned int)k) { c_lo = s1; cnt1 = v1; b1 += 1; }
        else if (s3 >= (unsigned int)k) { c_lo = s2; cnt1 = v2; b1 += 2; }
        else { c_lo = s3; cnt1 = v3; b1 += 3; }
        break;
      }
      int b2 = KNN_BINS;
      unsigned int cnt2 = cnt1;
      const bool lvl2 = cnt1 > 8u && c_lo + cnt1 > (unsigned int)k;
      if (lvl2) {
        // ---- 2. level-2 histogram over the recorded candidates of bin b1
#pragma unroll
        for (int b = 0; b < KNN_BINS; ++b) knn_sts16(hs + b * KNN_COL, 0u);
        int it = 0;
        knn_rows(g, pkeys, n, cell_start, c0, c1, c2, rho, [&](int j, int hi) {
          const unsigned int w = it < KR_WORDS ? rec[it] : knn_word(P, pq, j, hi, bound2, scale1);
          ++it;
          if (w == 0xffffffffu) return;
#pragma unroll
          for (int s_ = 0; s_ < 4; ++s_) {
            if ((int)((w >> (8 * s_)) & 255u) != b1) continue;
            const int bb = knn_sub_bin(P, j + s_, pq.x, pq.y, pq.z, scale1, b1);
            const unsigned int v = knn_lds16(hs + bb * KNN_COL);
            knn_sts16(hs + bb * KNN_COL, v == 65535u ? v : v + 1u);
          }
        });
        for (b2 = 0; b2 < KNN_BINS; ++b2) {
          cnt2 = knn_lds16(hs + b2 * KNN_COL);
          if (c_lo + cnt2 >= (unsigned int)k) break;
          c_lo += cnt2;
        }
      }
      const unsigned int t = (unsigned int)k - c_lo;
      const bool take_all = (t == cnt2);
      if (!take_all && cnt2 > 8u) {
        fallback = true;                     // more than 8 candidates tied in the boundary sub-bin: repeated selection, rare
      } else {
        // ---- 3. emit from the record.  Bins below `thr` are neighbours outright; candidates of bin b1 that need a
        // second look (sub-bin under the second level, rank among the boundary candidates) only leave their position
        // in the thread's dead histogram column (two counters per entry)
        const bool plain_b1 = take_all && !lvl2;
        int thr = plain_b1 ? b1 + 1 : b1;
        asm volatile("" : "+r"(thr));        // opaque: one register instead of a select at every use
        int nb = 0;
        int it = 0;
        // four words of the record per step, loaded together: the record comes back from L2 (480 KB per SM do not
        // stay in L1) and one dependent load per word left the pass waiting on it
        knn_rows16(g, pkeys, n, cell_start, c0, c1, c2, rho, [&](int jg, int hi) {
          const int left = hi - jg;
          const int nw = left >= 16 ? 4 : (left + 3) >> 2;
          unsigned int w0, w1 = 0xffffffffu, w2 = 0xffffffffu, w3 = 0xffffffffu;
          if (it + 3 < KR_WORDS) {
            w0 = rec[it]; w1 = rec[it + 1]; w2 = rec[it + 2]; w3 = rec[it + 3];      // words past nw: ignored below
          } else {
            w0 = it < KR_WORDS ? rec[it] : knn_word(P, pq, jg, hi, bound2, scale1);
            if (nw > 1) w1 = it + 1 < KR_WORDS ? rec[it + 1] : knn_word(P, pq, jg + 4, hi, bound2, scale1);
            if (nw > 2) w2 = it + 2 < KR_WORDS ? rec[it + 2] : knn_word(P, pq, jg + 8, hi, bound2, scale1);
            if (nw > 3) w3 = it + 3 < KR_WORDS ? rec[it + 3] : knn_word(P, pq, jg + 12, hi, bound2, scale1);
          }
          it += nw;
          for (int i = 0; i < nw; ++i) {
            const unsigned int w = i == 0 ? w0 : (i == 1 ? w1 : (i == 2 ? w2 : w3));
            if (w == 0xffffffffu) continue;
            const int j0 = jg + 4 * i;
            const int e0 = (int)(w & 255u), e1 = (int)((w >> 8) & 255u), e2 = (int)((w >> 16) & 255u), e3 = (int)(w >> 24);
            const bool has_b1 = !plain_b1 && (e0 == b1 || e1 == b1 || e2 == b1 || e3 == b1);
            if (lvl2 && has_b1) {
              // second level (rare): slot by slot, so that the row keeps the order of the walk
#pragma unroll 1
              for (int s_ = 0; s_ < 4; ++s_) {
                const int b = (int)((w >> (8 * s_)) & 255u);
                const int j = j0 + s_;
                if (b < b1) {
                  out_j[(int64_t)(cnt++) * DC_SLICE] = j;
                } else if (b == b1) {
                  const int bb = knn_sub_bin(P, j, pq.x, pq.y, pq.z, scale1, b1);
                  if (bb < b2 || (bb == b2 && take_all)) {
                    out_j[(int64_t)(cnt++) * DC_SLICE] = j;
                  } else if (bb == b2 && nb < 8) {
                    knn_sts16(hs + 2 * nb * KNN_COL, (unsigned int)j);
                    knn_sts16(hs + (2 * nb + 1) * KNN_COL, (unsigned int)j >> 16);
                    ++nb;
                  }
                }
              }
            } else {
              if (e0 < thr) out_j[(int64_t)cnt * DC_SLICE] = j0;
              cnt += e0 < thr ? 1 : 0;
              if (e1 < thr) out_j[(int64_t)cnt * DC_SLICE] = j0 + 1;
              cnt += e1 < thr ? 1 : 0;
              if (e2 < thr) out_j[(int64_t)cnt * DC_SLICE] = j0 + 2;
              cnt += e2 < thr ? 1 : 0;
              if (e3 < thr) out_j[(int64_t)cnt * DC_SLICE] = j0 + 3;
              cnt += e3 < thr ? 1 : 0;
              if (has_b1) {
                // the (<= 8) candidates of the boundary bin: positions only, ranked after the walk
#pragma unroll
                for (int s_ = 0; s_ < 4; ++s_) {
                  if ((int)((w >> (8 * s_)) & 255u) == b1 && nb < 8) {
                    knn_sts16(hs + 2 * nb * KNN_COL, (unsigned int)(j0 + s_));
                    knn_sts16(hs + (2 * nb + 1) * KNN_COL, (unsigned int)(j0 + s_) >> 16);
                    ++nb;
                  }
                }
              }
            }
          }
        });
        if (!take_all) {
          // rank the boundary candidates by (d2, original index): eight independent loads, the warp converged
          double bd[8];
          int bj[8], bt[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool have = i < nb;
            const int j = have ? (int)(knn_lds16(hs + 2 * i * KNN_COL) | (knn_lds16(hs + (2 * i + 1) * KNN_COL) << 16)) : 0;
            const dc_point pj = dc_ld_point(P + j);
            bd[i] = have ? dc_dist2(pj, pq) : INFINITY;
            bj[i] = have ? j : 0x7fffffff;
            bt[i] = have ? (int)pj.tag : 0x7fffffff;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            unsigned int rank = 0u;
#pragma unroll
            for (int m = 0; m < 8; ++m) rank += knn_less(bd[m], bt[m], bd[i], bt[i]) ? 1u : 0u;
            if (i < nb && rank < t) out_j[(int64_t)(cnt++) * DC_SLICE] = bj[i];
          }
        }
      }
    }
    if (fallback) {
      const int slot = atomicAdd(counters + 1, 1);
      fb_list[slot] = make_int2((int)q, rho);
    }
#ifdef DC_KNN_STATS
    atomicAdd(counters + 2 + (rho < 9 ? rho : 9), 1);          // final ring of the query (tools/knn_ring_stats.py)
    atomicAdd((unsigned long long*)(counters + 14), (unsigned long long)n_words);
    // thread cycles by final ring, in the last 128 bytes of the (otherwise unused) fallback list
    atomicAdd((unsigned long long*)(fb_list + nq) - 16 + (rho < 9 ? rho : 9), (unsigned long long)(clock64() - stats_t0));
#endif
  }
  if (!fallback)
    for (int c = cnt; c < k; ++c) out_j[(int64_t)c * DC_SLICE] = -1;
}

__global__ void knn_record_init_kernel(int32_t* counters) { if (threadIdx.x < 16) counters[threadIdx.x] = 0; }

extern "C" int dc_knn_recorded(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                               const dc_grid_spec* spec, const int32_t* cell_start, int k, double r, int32_t* ell_idx,
                               void* temp, size_t* temp_bytes, void* stream) {
  if (!temp_bytes) return dc_set_error(DC_ERR_ARG, "dc_knn_recorded: temp_bytes is NULL");
  if (k < 1) return dc_set_error(DC_ERR_ARG, "dc_knn_recorded: k must be positive");
  if (nq > 2147483647LL || n > 2147483647LL) return dc_set_error(DC_ERR_OVERFLOW, "dc_knn_recorded: more than 2^31-1 points");
  // workspace: 64-byte header (counters) | fallback list int2[nq]
  const size_t need = 64 + (size_t)(nq > 0 ? nq : 1) * 8;
  if (!temp) { *temp_bytes = need; return DC_OK; }
  if (*temp_bytes < need) return dc_set_error(DC_ERR_ARG, "dc_knn_recorded: workspace too small");
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
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* counters = (int32_t*)temp;
  int2* fb = (int2*)((char*)temp + 64);
  knn_record_init_kernel<<<1, 32, 0, st>>>(counters);
#ifdef DC_KNN_STATS
  DC_CUDA_CHECK(cudaMemsetAsync((char*)(fb + nq) - 128, 0, 128, st));
#endif
  DC_LAUNCH_CHECK();
  const int64_t n_slices = (nq + DC_SLICE - 1) / DC_SLICE;
  const int blocks = dc_blocks(n_slices * DC_SLICE, KNN_THREADS);
  knn_record_kernel<<<blocks, KNN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, nq, g, cell_start, k,
                                                    r2cap, max_ring, fb, counters, ell_idx);
  DC_LAUNCH_CHECK();
  int dev = 0, sms = 148;
  DC_CUDA_CHECK(cudaGetDevice(&dev));
  DC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  knn_thread_list_kernel<<<sms * 4, KNN_THREADS, 0, st>>>((const dc_point*)P, pkeys, n, (const dc_point*)Q, qkeys, g, cell_start,
                                                         k, r2cap, max_ring, fb, counters, ell_idx);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

struct kt_is_cell_start {
  const uint64_t* keys;
  int sub_bits;
  __host__ __device__ bool operator()(int i) const { return i == 0 || (keys[i] >> sub_bits) != (keys[i - 1] >> sub_bits); }
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
  kt_is_cell_start pred{qkeys, spec ? spec->sub_bits : 0};
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
  prm.tile_stride = (k + KT_TAIL) | 1;                          // odd: row-per-lane accesses are conflict-free
  const int union_words = 32 * prm.tile_stride > 32 * KT_HSTRIDE ? 32 * prm.tile_stride : 32 * KT_HSTRIDE;
  prm.warp_words = KT_WARP_FIXED + union_words;
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
  const size_t smem = (size_t)KT_WARPS * prm.warp_words * sizeof(int);
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
    while (pos > 0 && knn_less(dd, (int)P[jj].tag, d[pos - 1], (int)P[j[pos - 1]].tag)) { d[pos] = d[pos - 1]; j[pos] = j[pos - 1]; --pos; }
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

