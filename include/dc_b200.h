/* depth_correction_b200 -- C ABI of the B200 (sm_100a) map-consistency hot path.
 *
 * The reference (ctu-vras/depth_correction) is pure Python with no FFI layer; its boundary for this
 * path is the Python API.  Each entry point below names the reference
 * call (file:line under /root/reference/src/depth_correction/) whose native work it replaces.
 * The Python mirror in depth_correction_b200/ binds these symbols with ctypes;
 * there are no torch types in any signature.
 *
 * Conventions
 *   - every pointer is DEVICE memory unless the name ends in _host; buffers are caller-allocated
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it unless noted
 *   - return value: 0 = ok, otherwise a DC_ERR_* code; dc_last_error() gives a message (thread-local)
 *   - functions with (temp, temp_bytes): call with temp == NULL to get the size in *temp_bytes
 *   - dtype: DC_F32 / DC_F64 storage of caller tensors; all arithmetic is fp64
 *   - "sorted space": positions after sorting the map by grid cell; `order[s]` = original row of
 *     sorted position s.  Graphs are sliced-ELL in sorted space: rows are grouped in slices of 32,
 *     slice t holds width_t = (slice_ptr[t+1]-slice_ptr[t])/32 columns, entry (row, c) lives at
 *     ell_idx[slice_ptr[row/32] + c*32 + row%32], -1 = no neighbour.
 */
#ifndef DC_B200_H
#define DC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DC_OK 0
#define DC_ERR_CUDA 1
#define DC_ERR_ARG 2
#define DC_ERR_OVERFLOW 3

#define DC_F32 0
#define DC_F64 1

/* model kinds (model.py:149-286) */
#define DC_MODEL_NONE 0
#define DC_MODEL_POLYNOMIAL 1        /* d' = d - sum_k w_k g^e_k            (model.py:194-205) */
#define DC_MODEL_SCALED_POLYNOMIAL 2 /* d' = d (1 - sum_k w_k g^e_k)        (model.py:250-261) */
#define DC_MAX_TERMS 8

/* loss kinds (loss.py:216-370) and flags */
#define DC_LOSS_MIN_EIGVAL 0
#define DC_LOSS_TRACE 1
#define DC_FLAG_NORMALIZATION 1 /* lambda0 / clamp(sum lambda, 1e-6)   (loss.py:253-254) */
#define DC_FLAG_SQRT 2          /* sqrt after relu                     (loss.py:286-287) */
#define DC_FLAG_RAW 4           /* per-point value before relu/sqrt (general path: inliers, offsets) */

/* per-point flag bits in the packed scan records */
#define DC_PT_MODEL_MASK 1u /* depth is corrected by the model (DepthCloud.mask of the scan, model.py:254-260) */
#define DC_PT_LOSS_MASK 2u  /* point contributes a loss term (mask argument of the losses, loss.py:245-248) */

typedef struct dc_grid_spec {
  double origin[3]; /* min corner, xyz */
  double cell;      /* cell edge length */
  int32_t dims[3];  /* cells along x, y, z */
  int32_t axis[3];  /* axis[0] = fastest varying axis of the cell key ... axis[2] = slowest */
  int32_t sub_bits; /* 0 or 6: low key bits = Morton code of the point's 4x4x4 sub-cell, so that points of one cell are
                       stored in a spatially coherent order (consecutive queries of a warp are neighbours in space);
                       the cell of a key is key >> sub_bits */
} dc_grid_spec;

const char* dc_last_error(void);
int dc_version(void);

/* ---------------------------------------------------------------------------------------------
 * Kernel 1: neighbour search.  Replaces nearest_neighbors() = scipy cKDTree build + query /
 * query_ball_point on the host (nearest_neighbors.py:22-80), called from
 * DepthCloud.update_neighbors (depth_cloud.py:210-215).
 * ------------------------------------------------------------------------------------------- */

/* min/max over rows of pts[n,3]: out6 = {minx,miny,minz,maxx,maxy,maxz}; bad_count = #non-finite rows */
int dc_bounds(const void* pts, int dtype, int64_t n, double* out6, int32_t* bad_count, void* stream);

/* cell key of every row (linear index with spec->axis ordering, shifted left by spec->sub_bits with the sub-cell Morton
 * code below it) and ids = 0..n-1 */
int dc_cell_keys(const void* pts, int dtype, int64_t n, const dc_grid_spec* spec_host, uint64_t* keys, int32_t* ids,
                 void* stream);

/* the same for MANY clouds stacked along the slowest key axis (cloud s = rows first[s] .. first[s+1] of pts) so that one
 * search serves all of them without ever pairing points of different clouds: cloud s owns the cell layers
 * [s * period, s * period + period - guard) of that axis, guard >= the largest ring of cells the search will visit
 * (ceil(r / cell)); spec->dims of the slowest axis must be n_clouds * period.  Batched form of the per-scan searches
 * of local_feature_cloud (preproc.py:50, train.py:97-104). */
int dc_cell_keys_stacked(const void* pts, int dtype, int64_t n, const dc_grid_spec* spec_host, const int64_t* first,
                         int n_clouds, int period, int guard, uint64_t* keys, int32_t* ids, void* stream);

/* stable radix sort of (key, id) pairs on key bits [0, end_bit) */
int dc_sort_pairs(const uint64_t* keys_in, uint64_t* keys_out, const int32_t* ids_in, int32_t* ids_out, int64_t n,
                  int end_bit, void* temp, size_t* temp_bytes, void* stream);

/* sorted[s] = {double(pts[order[s]]), tag = order[s]}   (32-byte records); inv_order[order[s]] = s when given */
int dc_gather_points(const void* pts, int dtype, const int32_t* order, int64_t n, void* sorted_points,
                     int32_t* inv_order, void* stream);

/* dense table cell_start[c] = first sorted position whose cell (key >> sub_bits) is >= c, c in [0, n_cells]
 * (optional accelerator) */
int dc_cell_table(const uint64_t* keys_sorted, int64_t n, int64_t n_cells, int sub_bits, int32_t* cell_start, void* stream);

/* radius mode, pass 1: counts[q] = #{p : |p - q|^2 <= r^2} (fp64, same summation order as cKDTree),
 * slice_width[t] = max count in slice t.  Queries must be sorted by the same grid (self query: Q == P). */
int dc_radius_count(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                    const dc_grid_spec* spec_host, const int32_t* cell_start, double r, int32_t* counts,
                    int32_t* slice_width, void* stream);

/* slice_ptr[t] = 32 * exclusive_sum(slice_width)[t], t in [0, n_slices]  (int64) */
int dc_ell_offsets(const int32_t* slice_width, int64_t n_slices, int64_t* slice_ptr, void* temp, size_t* temp_bytes,
                   void* stream);

/* radius mode, pass 2: fill ell_idx (sorted-space indices, ascending; -1 padding) */
int dc_radius_fill(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                   const dc_grid_spec* spec_host, const int32_t* cell_start, double r, const int64_t* slice_ptr,
                   int32_t* ell_idx, void* stream);

/* kNN (r <= 0) or kNN within r (strict <, like cKDTree's distance_upper_bound): fixed-width ELL
 * (slice_ptr[t] = 32*k*t) holding the k nearest of every query in NO particular order (the step kernels
 * only need the set); ell_d2 (optional) receives the squared distances.  Exact ties at the k-th distance
 * are broken by the smaller ORIGINAL index (the tag of the map record), so the result does not depend on
 * the cell size.  One query per thread, rows re-walked by the emit pass; the only variant that can return distances.
 * The host code calls dc_knn_recorded (below) when it needs the lists only. */
int dc_knn(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
           const dc_grid_spec* spec_host, const int32_t* cell_start, int k, double r, int32_t* ell_idx, double* ell_d2,
           void* stream);
/* The same search, one WARP per occupied query cell (k <= 128; an independent second implementation, used as a cross-check): the candidate block of a cell is
 * staged once in registers as fp32 offsets from the cell centre, the cell's queries are streamed through it with a
 * shared-memory histogram select, and every query whose fp32 classification is not provably the fp64 one (plus cells
 * whose block exceeds the register slots or needs more than four rings) is finished by the fp64 one-thread-per-query
 * kernel of dc_knn.  Same result as dc_knn bit for bit.  Replaces cKDTree.query (nearest_neighbors.py:48-49).
 * temp: 64 + 12 nq bytes + select scratch (two-phase size query). */
int dc_knn_cells(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                 const dc_grid_spec* spec_host, const int32_t* cell_start, int k, double r, int32_t* ell_idx, void* temp,
                 size_t* temp_bytes, void* stream);
/* The same search with ONE distance pass per query: the histogram pass records (index, bin) of every candidate inside the
 * bound in a thread-private list and the emit pass replays the record, re-reading only the candidates of the boundary
 * bin (one byte per visited candidate; iterations beyond the 256-word record are recomputed in place); queries with
 * > 8 exact ties at the k-th place are finished by the kernel of dc_knn.  Same rows as dc_knn, entry by entry.
 * PRECONDITION: the n records of P are followed by DC_KNN_PAD (3) more READABLE records (content ignored): the first
 * pass reads four consecutive records per step without clamping the last ones to the end of a row.
 * Replaces cKDTree.query (nearest_neighbors.py:48-49).  temp: 64 + 8 nq bytes (two-phase size query). */
#define DC_KNN_PAD 3
int dc_knn_recorded(const void* P, const uint64_t* pkeys, int64_t n, const void* Q, const uint64_t* qkeys, int64_t nq,
                    const dc_grid_spec* spec_host, const int32_t* cell_start, int k, double r, int32_t* ell_idx, void* temp,
                    size_t* temp_bytes, void* stream);
/* order every kNN row by (d^2, original index = tag of the map record) in place: cKDTree.query returns
 * distance-sorted rows (nearest_neighbors.py:48); needed only when the reference layout is exported */
int dc_knn_sort_rows(const void* P, int64_t n, int k, int32_t* ell_idx, double* ell_d2, int64_t nq, void* stream);
/* squared distances of a kNN graph recomputed from the records (bit-identical to the ones the selection
 * compared); the training path calls dc_knn with ell_d2 == NULL and never needs them */
int dc_knn_distances(const void* P, const void* Q, int k, const int32_t* ell_idx, int64_t nq, double* ell_d2,
                     void* stream);

/* export to the reference layout: out[order_q[row], c] = order_p[ell(row, c)] (int64, -1 padding), K columns */
int dc_ell_to_padded(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t nq,
                     const int32_t* order_p, const int32_t* order_q, int K, int64_t* out, void* stream);
/* distances: out[order_q[row], c] = sqrt(d2) (fp64; +inf where missing) */
int dc_ell_to_dist(int k, const double* ell_d2, const int32_t* ell_idx, int64_t nq, const int32_t* order_q, double* out,
                   void* stream);
/* ascending sort of every row of an [n,K] int64 matrix with -1 kept last (radius mode row order) */
int dc_sort_rows(int64_t* rows, int64_t n, int K, void* temp, size_t* temp_bytes, void* stream);

/* import a reference-layout graph (neighbors int64 [n,K], -1 = missing, original order) into sorted space.
 * pass 1 (ell_idx == NULL): slice_width; pass 2: fill.  inv_order[orig] = sorted position. */
int dc_padded_to_ell(const int64_t* neighbors, int64_t n, int K, const int32_t* order, const int32_t* inv_order,
                     int32_t* slice_width, const int64_t* slice_ptr, int32_t* ell_idx, void* stream);

/* transposed graph (who lists me as a neighbour), needed by the backward pass for asymmetric (kNN) graphs:
 * step 1 writes one (dst<<32 | src) pair per valid edge and the total; the caller sorts the pairs with
 * dc_sort_keys; step 2 computes in-degrees and slice widths; step 3 fills the transposed ELL. */
int dc_graph_edges(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t n_rows,
                   const int64_t* edge_offset, uint64_t* pairs, void* stream);
int dc_graph_degrees(const int64_t* slice_ptr, const int32_t* ell_idx, int64_t n_rows,
                     int32_t* out_degree, void* stream);
/* stable radix sort on key bits [begin_bit, end_bit): edge pairs arrive ordered by src, so sorting the
 * dst bits alone (begin_bit = 32) yields (dst, src) order in 3-4 passes */
int dc_sort_keys(const uint64_t* keys_in, uint64_t* keys_out, int64_t n, int begin_bit, int end_bit, void* temp,
                 size_t* temp_bytes, void* stream);
int dc_exclusive_sum_i32_i64(const int32_t* in, int64_t* out, int64_t n, void* temp, size_t* temp_bytes, void* stream);
int dc_transpose_widths(const uint64_t* pairs_sorted, int64_t n_edges, int64_t n_cols, int32_t* in_degree,
                        int32_t* slice_width, void* stream);
int dc_transpose_fill(const uint64_t* pairs_sorted, int64_t n_edges, int64_t n_cols, const int64_t* slice_ptr_t,
                      int32_t* ell_idx_t, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Kernels 2/3: fused fixed-graph step.  Replaces, per training iteration,
 *   model(cloud) + cloud.transform(pose) + DepthCloud.concatenate      (preproc.py:80-119)
 *   update_points / update_mean / update_cov / update_eig              (depth_cloud.py:122-128,291-399)
 *   min_eigval_loss / trace_loss + reduce                              (loss.py:125-150,216-370)
 *   loss.backward() down to model.w, model.exponent and the poses     (train.py:306-312)
 * Scan records are packed once, in sorted space:
 *   rec_dir[s]  = {dir.x, dir.y, dir.z, depth}      (float4 / double4)
 *   rec_vp[s]   = {vp.x, vp.y, vp.z, inc_angle}     (float4 / double4)
 *   rec_meta[s] = scan_id << 2 | DC_PT_* flags      (uint32)
 * poses: fp64 [S,12] row-major 3x4 (already corrected, see dc_pose_compose).
 * ------------------------------------------------------------------------------------------- */

/* pack rows of one scan: row i of the scan is global row `first + i` and lands at inv_order[first + i]
 * (sorted space), or at first + i when inv_order == NULL (original order, used by dc_step_chain) */
int dc_pack_records(const void* vps, const void* dirs, const void* depth, const void* inc_angles,
                    const uint8_t* model_mask, const uint8_t* loss_mask, int dtype, int64_t first, int64_t count,
                    int scan_id, const int32_t* inv_order, void* rec_dir, void* rec_vp, uint32_t* rec_meta,
                    void* stream);
/* batched forms over all scans in one launch.  scan_ptr_table: device array [n_scans][5] of uint64 device
 * addresses {vps, dirs, depth, inc_angles, model_mask} (0 = absent); first: int64 [n_scans+1] global row of the
 * first point of every scan.  dc_pack_records_batched writes the sorted-space copy (through inv_order) and, when
 * rec_*_o are given, the original-order copy; dc_world_points_batched writes out[n,3] fp64 with poses fp64 [S,16]. */
int dc_pack_records_batched(const void* scan_ptr_table, const int64_t* first, int n_scans, int64_t n, int dtype,
                            const int32_t* inv_order, void* rec_dir, void* rec_vp, uint32_t* rec_meta, void* rec_dir_o,
                            void* rec_vp_o, uint32_t* rec_meta_o, void* stream);
int dc_world_points_batched(const void* scan_ptr_table, const int64_t* first, int n_scans, int64_t n, int dtype,
                            const double* poses, double* out, void* stream);
/* overwrite the DC_PT_LOSS_MASK bit from a global-order mask (NULL = all true) */
int dc_set_loss_mask(const uint8_t* loss_mask, int64_t n, const int32_t* order, uint32_t* rec_meta, void* stream);

/* pass A: corrected world points, 32-byte fp64 records in sorted space */
int dc_step_points(const void* rec_dir, const void* rec_vp, const uint32_t* rec_meta, int dtype, int64_t n,
                   const double* poses, int n_scans, int model_kind, const double* w, const double* exponent, int n_terms,
                   void* points_out, void* stream);

/* pass B: neighbourhood mean / covariance / eigen / loss.
 * outputs (any may be NULL): loss_pp[n] per-point loss (0 where DC_PT_LOSS_MASK is clear, unless RAW),
 * stash[n,8] = {mean xyz, v0 xyz, alpha, beta} for the backward pass, eigvals[n,3], partial sums ->
 * loss_sum[0] = sum of per-point losses over the loss mask, loss_sum[1] = number of masked points.
 * partials: scratch of 2*grid doubles + one uint32 counter (zeroed by the caller once). */
int dc_step_forward(const void* points, const uint32_t* rec_meta, int64_t n, const int64_t* slice_ptr,
                    const int32_t* ell_idx, int loss_kind, int flags, double* loss_pp, double* stash, double* eigvals,
                    double* loss_sum, void* partials, size_t partials_bytes, void* stream);

/* passes B + C1 in ONE kernel, for the mean / sum reductions (no DC_FLAG_RAW): the upstream gradient of every loss
 * term is then a single scalar that dc_step_chain's caller applies afterwards, so the thread that has just finished the
 * eigen epilogue walks its index column again and adds A_i (p_j - m_i) to g_sorted32[j] (float32 [n,4], SORTED space,
 * zeroed inside this call) with one 16-byte vector reduction per edge.  No stash, no second read of the index array
 * from HBM.  Replaces update_mean/cov/eig + loss (depth_cloud.py:291-399, loss.py:216-370) AND the neighbour part of
 * loss.backward() (train.py:306-312).  loss_pp may be NULL. */
int dc_step_forward_scatter(const void* points, const uint32_t* rec_meta, int64_t n, const int64_t* slice_ptr,
                            const int32_t* ell_idx, int loss_kind, int flags, double* loss_pp, void* g_sorted32,
                            double* loss_sum, void* partials, size_t partials_bytes, void* stream);

/* pass C1: g_j = sum_{i : j in N(i)} u_i A_i (p_j - m_i), a gather over the TRANSPOSED graph (sorted space),
 * written to the point's ORIGINAL row: g_out fp64 [n,3].  upstream_pp: optional per-point upstream gradient
 * in sorted space (NULL = 1 for every row).  No atomics. */
int dc_step_backward(const void* points, int64_t n, const int64_t* slice_ptr_t, const int32_t* ell_idx_t,
                     const double* stash, const double* upstream_pp, const int32_t* order, double* g_out, void* stream);

/* pass C1, scatter form: row i adds u_i A_i (p_j - m_i) to g_sorted[j] (SORTED space, zeroed by the caller) for
 * every j of its own list with L2 reductions -- no transposed graph needed.  Used for the first backward passes on
 * a kNN graph, before building the transpose pays off.  g_dtype = DC_F64: g_sorted fp64 [n,3], three reductions per
 * edge; DC_F32: g_sorted float32 [n,4] (16-byte aligned), ONE vector reduction per edge (3x the edge rate, fp32
 * accumulation -- meant for large maps, where the rounding averages out in the chain stage). */
int dc_step_backward_scatter(const void* points, int64_t n, const int64_t* slice_ptr, const int32_t* ell_idx,
                             const double* stash, const double* upstream_pp, void* g_sorted, int g_dtype, void* stream);

/* pass C2 + C3: chain g (fp64 [n,3] or, g_dtype = DC_F32, float32 [n,4]; row of original point i: g[g_index[i]], or g[i] when g_index == NULL) through p = R_s (vp + d' dir) + t_s to dw[n_terms],
 * dexponent[n_terms] (NULL to skip) and dposes[S,12]; outputs are ACCUMULATED (caller zeroes).
 * Records are the dc_pack_records layout packed with inv_order == NULL (original order).  The block table
 * aligns blocks to scans: block b covers rows [block_start[b], block_start[b] + block_count[b]) of scan
 * block_scan[b]; blocks of scan s are [scan_block_first[s], scan_block_first[s+1]).  partials: scratch of
 * n_blocks * (12 + 2*DC_MAX_TERMS) doubles.  Block reductions + fixed-order final sums: deterministic. */
int dc_step_chain(const void* g, int g_dtype, const int32_t* g_index, const void* rec_dir, const void* rec_vp, const uint32_t* rec_meta, int dtype,
                  const int32_t* block_scan, const int64_t* block_start, const int32_t* block_count, int n_blocks,
                  const int32_t* scan_block_first, const double* poses, int n_scans, int model_kind, const double* w,
                  const double* exponent, int n_terms, double* partials, double* dw, double* dexponent, double* dposes,
                  void* stream);

/* SE(3) correction T_s = P_s * Delta(delta_s): create_corrected_poses (eval.py:68-82) +
 * xyz_axis_angle_to_matrix (transform.py:68-78).  poses fp64 [S,16]; deltas fp64 [n_deltas,6] with
 * n_deltas == S or 1 (shared, PoseCorrection.common / .sequence); out fp64 [S,12]. */
int dc_pose_compose(const double* poses, const double* deltas, int n_scans, int n_deltas, double* out, void* stream);
int dc_pose_compose_backward(const double* poses, const double* deltas, int n_scans, int n_deltas, const double* dout,
                             double* ddeltas, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Unfused feature kernels on caller-ordered tensors, behind DepthCloud.update_mean / update_cov /
 * update_eig / update_normals / update_incidence_angles (depth_cloud.py:291-424) and their autograd.
 * neighbors int64 [n,K] (-1 = missing), weights float32 [n,K] or NULL (= valid mask).
 * ------------------------------------------------------------------------------------------- */
int dc_features(const void* points, int dtype, int64_t n, const int64_t* neighbors, const float* weights, int K,
                void* mean, void* cov, void* stream);
/* Per-point features of many clouds at once, second half (the first is dc_step_forward on the stacked graph with
 * points = the sorted records of the search): stash fp64 [n,8] and eigvals_sorted fp64 [n,3] per SORTED row ->
 * eigvals / mean / normals [n,3] and inc_angles [n] in the caller's order and dtype (any output may be NULL).
 * Replaces update_mean / update_eig / update_normals / update_incidence_angles of every scan
 * (depth_cloud.py:291-295, 376-424) called from local_feature_cloud (preproc.py:50). */
int dc_local_features_finish(const double* stash, const double* eigvals_sorted, const int32_t* order, const void* dirs, int dtype,
                             int64_t n, int use_normal_sign, void* eigvals, void* mean, void* normals, void* inc_angles,
                             void* stream);
/* Feature masks in one launch: filter_valid_neighbors / filter_eigenvalue(s) / filter_eigenvalue_ratio(s) /
 * within_bounds (filters.py:85-113, 184-254), as composed by local_feature_cloud (preproc.py:53-62) and
 * global_cloud_mask (preproc.py:130-142).  vals: [n, stride] values of `dtype` (eigvals: stride 3; a scalar field such
 * as a dispersion: stride 1), bounds_host: n_bounds (<= 16) records {kind, a, b, lo, hi} of doubles on the HOST with
 * kind 0: lo <= vals[i,a] <= hi, kind 1: lo <= vals[i,a] / vals[i,b] <= hi; bounds are inclusive, compared in `dtype`
 * like torch does, -inf / +inf / NaN disable a side; valid_counts (int64 [n], may be NULL) >= min_valid.
 * init != 0: mask = result; init == 0: mask &= result.  mask: uint8 / bool [n]. */
int dc_feature_mask(const void* vals, int dtype, int64_t n, int stride, const double* bounds_host, int n_bounds,
                    const int64_t* valid_counts, int64_t min_valid, int init, uint8_t* mask, void* stream);
int dc_features_backward(const void* points, int dtype, int64_t n, const int64_t* neighbors, const float* weights, int K,
                         const void* gmean, const void* gcov, void* gpoints, void* stream);
int dc_eigh3(const void* cov, int dtype, int64_t n, void* eigvals, void* eigvecs, void* stream);
int dc_eigh3_backward(const void* eigvals, const void* eigvecs, int dtype, int64_t n, const void* geigvals,
                      const void* geigvecs, void* gcov, void* stream);
/* points of one scan in the map frame, out[n,3] fp64 = R (vp + depth dir) + t with pose = fp64 4x4 row-major
 * (device): cloud.transform(pose).to_points() for the initial global cloud (preproc.py:108-119,180).  vps may be NULL. */
int dc_world_points(const void* vps, const void* dirs, const void* depth, int dtype, int64_t n, const double* pose,
                    double* out, void* stream);
/* DepthCloud.from_points (depth_cloud.py:592-638): dirs[n,3] = (points - vps) / depth where depth > 0,
 * depth[n] = |points - vps|, vps_out[n,3] = vps (zeros when vps == NULL; vps_out may be NULL).  Cloud dtype. */
int dc_from_points(const void* points, const void* vps, int dtype, int64_t n, void* dirs, void* depth, void* vps_out,
                   void* stream);
/* normals = -sign(dirs . v0) v0, inc = arccos(|dirs . n|) or arccos(-dirs . n)  (depth_cloud.py:401-424) */
int dc_normals_angles(const void* dirs, const void* eigvecs, int dtype, int64_t n, int use_normal_sign, void* normals,
                      void* inc_angles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Per-scan preprocessing filters.
 *
 * filter_grid (filters.py:24-82): one survivor per occupied voxel of edge grid_res.
 *   dc_voxel_keys: position t of the sequence (seq[t], or n-1-t when reversed, or t) -> 63-bit voxel key of that
 *     point (floor(x / grid_res) per axis in the cloud's dtype, biased 21-bit fields), ids[t] = t; *bad counts
 *     points whose cell index does not fit (|index| >= 2^20) or is NaN.
 *   (dc_sort_pairs on 63 bits groups the voxels; positions stay ascending inside a group.)
 *   dc_voxel_pick: for the last entry of every group: out_val = point index of the survivor (the LAST position
 *     of the voxel in the sequence, as the reference's dict keeps the last value), out_key = first position of
 *     the voxel in the sequence (dict order) or the survivor's index (preserve_order); 2^32 / -1 elsewhere;
 *     *count += number of voxels.  A second dc_sort_pairs on 33 bits of out_key orders the survivors.
 * filter_shadow_points (filters.py:257-309): keep[i] = min_k a_ik >= angle_lo && max_k a_ik <= angle_hi with
 *   a_ik = angle at x_i between (vp_i - x_i) and (x_nbr - x_i) over the direction-space neighbours; invalid
 *   neighbours (weight != 1) count as (angle_lo + angle_hi) / 2.  angle_min / angle_max [n] optional outputs.
 * ------------------------------------------------------------------------------------------- */
int dc_voxel_keys(const void* points, int dtype, int64_t n, double grid_res, const int32_t* seq, int reversed,
                  uint64_t* keys, int32_t* ids, int32_t* bad, void* stream);
int dc_voxel_pick(const uint64_t* keys_sorted, const int32_t* ids_sorted, int64_t n, const int32_t* seq, int reversed,
                  int preserve_order, uint64_t* out_key, int32_t* out_val, int32_t* count, void* stream);
int dc_shadow_mask(const void* points, const void* vps, int dtype, const int64_t* dir_neighbors,
                   const float* dir_neighbor_weights, int64_t n, int K, double angle_lo, double angle_hi, uint8_t* keep,
                   void* angle_min, void* angle_max, void* stream);

/* neighbourhood statistics of global_cloud_mask (depth_cloud.py:330-354): mean_depth[n] = sum_k w d[nbr] / sum_k w,
 * mean_vp_dist[n] = sum_k w |vp[nbr] - mean_vp| / sum_k w (either output may be NULL); neighbors int64 [n,K],
 * weights float32 [n,K] or NULL (= neighbors >= 0). */
int dc_neighbor_stats(const void* depth, const void* vps, int dtype, const int64_t* neighbors, const float* weights,
                      int64_t n, int K, void* mean_depth, void* mean_vp_dist, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ICP-style losses between two consecutive scans (loss.py:373-565).
 * Correspondences: either explicit index lists sel1 / sel2 (int64 [m]; the reference's precomputed `masks`) or,
 * with sel1 == NULL, (t, nn[t]) for t < m = n1, kept when dist[t] <= threshold (nn / dist: nearest neighbour of
 * every point of cloud 1 in cloud 2 from dc_knn with k = 1, fp64 distances).  Points are rounded to float32 first
 * (loss.py:424-425), arithmetic in fp64.  normals_dtype: dtype of the normal arrays (may differ from the points').
 *   dc_icp_forward: out4 = {sum_t |n1.(p2-p1)| |n1|, sum_t |n2.(p2-p1)| |n2|, count, sum_t dist[t]} (point-to-plane)
 *                   or {sum_t |p2-p1|, 0, count, sum_t dist[t]} (point-to-point); partials: >= 32*blocks+16 bytes
 *                   (blocks = ceil(m/256)), the trailing 16 bytes zero before the first call.
 *   dc_icp_backward: gradients of coef2[0] * out4[0] + coef2[1] * out4[1] (coef2: 2 doubles on the device),
 *                   ACCUMULATED into fp64 g_points1 [n1,3], g_points2 [n2,3], g_normals1, g_normals2 (any may be NULL).
 * dc_f64_sort_keys / dc_f64_from_sort_keys: order-preserving uint64 keys of fp64 values (NaN last, *n_nan counts
 *   them) and back, for the inlier threshold torch.nanquantile(dists, ratio) via dc_sort_keys.
 * ------------------------------------------------------------------------------------------- */
int dc_icp_forward(const void* points1, const void* points2, int dtype, const void* normals1, const void* normals2,
                   int normals_dtype, const int64_t* nn, const double* dist, double threshold, const int64_t* sel1,
                   const int64_t* sel2, int64_t m, int point_to_plane, double* out4, void* partials, size_t partials_bytes,
                   void* stream);
int dc_icp_backward(const void* points1, const void* points2, int dtype, const void* normals1, const void* normals2,
                    int normals_dtype, const int64_t* nn, const double* dist, double threshold, const int64_t* sel1,
                    const int64_t* sel2, int64_t m, int point_to_plane, const double* coef2, double* g_points1,
                    double* g_points2, double* g_normals1, double* g_normals2, void* stream);
int dc_f64_sort_keys(const double* x, int64_t n, uint64_t* keys, int32_t* n_nan, void* stream);
int dc_f64_from_sort_keys(const uint64_t* keys, int64_t n, double* x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU slab exchange (no counterpart in the single-process reference).  Slab g of
 * n_ranks owns b[g] <= x < b[g+1] along the split axis (inner_boundaries = b[1] .. b[n_ranks-1], a DEVICE array:
 * the slab plan never leaves the device);
 * a point is sent to every slab with b[g] - halo <= x < b[g+1] + halo.
 *   dc_route_count : gmin / gmax (uint8 [n]) = first / last destination of every point of world_points (fp64 [n,3]),
 *                    counts int32 [n_ranks] = records per destination; scan_counts (optional) int32 [n_ranks][scan_stride]
 *                    = records per (destination, GLOBAL scan id scan_ids[s] < scan_stride) with first / scan_ids / n_scans
 *                    as in dc_route_pack, so that a receiver learns the size of every scan it will hold from the count
 *                    exchange alone
 *   dc_route_pack  : send_f [M,8] (cloud dtype: vp.xyz, dir.xyz, depth, inc_angle) and send_i int32 [M,4] (scan id,
 *                    row, model mask, owned) contiguous per destination (dest_offset int64 [n_ranks] = exclusive sum of
 *                    counts; cursor int32 [n_ranks] scratch), read through the scan pointer table of
 *                    dc_pack_records_batched; order inside a destination is arbitrary
 *   dc_route_keys  : keys[t] = scan id << 32 | row, ids[t] = t of received index records (for dc_sort_pairs)
 *   dc_route_unpack: received records gathered in `order` into vps / dirs [m,3], depth / inc [m], mask / owned
 *                    uint8 [m], gid int64 [m,2] = (scan id, row)
 * ------------------------------------------------------------------------------------------- */
int dc_route_count(const double* world_points, int axis, int64_t n, const double* inner_boundaries, int n_ranks, double halo,
                   const int64_t* first, const int32_t* scan_ids, int n_scans, uint8_t* gmin, uint8_t* gmax, int32_t* counts,
                   int32_t* scan_counts, int scan_stride, void* stream);
int dc_route_pack(const void* scan_ptr_table, const int64_t* first, const int32_t* scan_ids, int n_scans, int64_t n, int dtype,
                  const double* world_points, int axis, const double* inner_boundaries, int n_ranks, double halo,
                  const uint8_t* gmin, const uint8_t* gmax, const int64_t* dest_offset, int32_t* cursor, void* send_f,
                  int32_t* send_i, void* stream);
int dc_route_keys(const int32_t* recv_i, int64_t m, uint64_t* keys, int32_t* ids, void* stream);
/* hist[b] = number of points with floor-toward-zero((x_axis - a0) * scale) == b (clamped to 0 .. n_bins-1): the rank-local
 * part of the histogram SlabPartitioner.plan all-reduces to cut slabs of equal point count */
int dc_axis_histogram(const double* world_points, int axis, int64_t n, double a0, double scale, int n_bins, int32_t* hist,
                      void* stream);
int dc_route_unpack(const void* recv_f, const int32_t* recv_i, const int32_t* order, int64_t m, int dtype, void* vps, void* dirs,
                    void* depth, void* inc, uint8_t* mask, uint8_t* owned, int64_t* gid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DC_B200_H */
