/* isoext_b200.h -- C-ABI of libisoext_b200.so: the drop-in boundary of the B200-native
 * iso-surface extraction core.
 *
 * It replaces the reference's nanobind module `isoext_ext` (src/isoext_ext.cu:93-384 of
 * GuangyanCai/isoext v0.5.1): every entry point below names the reference interface it stands in
 * for.  Conventions:
 *   - plain C types only; every pointer named d_* / values / workspace / scratch / V / F is a
 *     DEVICE pointer allocated by the caller (the Python host layer uses torch storage);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it; functions that return
 *     counts synchronise that stream before returning (the reference is fully synchronous on the
 *     legacy default stream);
 *   - return value 0 = success, negative = error (ISOEXT_E_*); isoext_last_error() gives the
 *     message (thread-local).  The reference's std::runtime_error maps to these codes and the
 *     host layer re-raises RuntimeError with the same message text;
 *   - two-phase protocol (count, then emit) so that the caller allocates exact-size outputs;
 *   - grids are addressed as a local slab of a global grid (x_offset, X_global) so that the same
 *     entry points serve the single-GPU path (x_offset = 0, X_global = X) and the slab-sharded
 *     multi-GPU path.
 * Index conventions are the reference's: value index = x*Y*Z + y*Z + z (src/utils.cu:21-24),
 * cell index = x*(Y-1)*(Z-1) + y*(Z-1) + z (include/utils.cuh:43-47), corner/edge numbering of
 * include/shared_luts.cuh:3-38, `shape` = POINTS per axis (src/grid/uniform.cu:11-12).
 */
#ifndef ISOEXT_B200_H
#define ISOEXT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISOEXT_OK 0
#define ISOEXT_E_INVALID (-1)   /* bad argument                                           */
#define ISOEXT_E_CUDA (-2)      /* CUDA runtime error                                     */
#define ISOEXT_E_WORKSPACE (-3) /* workspace / scratch too small                          */
#define ISOEXT_E_CAPACITY (-4)  /* entry capacity exceeded: counts_out[0] = needed size   */
#define ISOEXT_E_METHOD (-5)    /* unknown MC method (src/mc/base.cu:23-25)               */

#define ISOEXT_METHOD_NAGAE 0    /* include/mc/nagae.cuh    (default, src/isoext_ext.cu:101) */
#define ISOEXT_METHOD_LORENSEN 1 /* include/mc/lorensen.cuh                                  */

/* ---- library ------------------------------------------------------------------------------ */
const char *isoext_last_error(void);
const char *isoext_build_info(void);
int isoext_abi_version(void);

/* Position of grid plane i along one axis: fma(i / (res-1), amax - amin, amin) in float32, i.e. the
 * per-axis term of get_vtx_pos_op (include/utils.cuh:71-79).  Host-side; no GPU needed. */
float isoext_axis_position(int64_t i, int64_t res, float amin, float amax);

/* Tuning knob of the volume-streaming kernel (development only): low byte = variant, next byte =
 * launched blocks per SM; 0 = default. */
int isoext_debug_set_signbits_variant(int v);
/* Development tuning knobs (key 0: blocks per SM of the chained-scan kernels); value 0 = default. */
int isoext_debug_set_tuning(int key, int value);

/* Development: per-kernel CUDA-event timing of every launch of this library. */
int isoext_debug_detail_enable(int on);
int isoext_debug_detail_report(char *buf, int buf_size);
int isoext_debug_detail_timeline(char *buf, int buf_size);

/* ---- measurement hooks (no reference counterpart; used by bench.py only) ---------------------
 * begin(): reset the launch counter and start recording a CUDA-event pair around every launch of the
 * volume-streaming kernel (the roofline's dominant kernel) on the stream it is launched on.
 * end(): total / count of those durations and the number of kernel launches since begin(). */
int isoext_profile_begin(void);
int isoext_profile_end(double *stream_ms_total, int64_t *stream_launches, int64_t *kernel_launches);

/* ---- UniformGrid.get_points  (src/grid/uniform.cu:22-30, include/utils.cuh:62-80) ----------
 * out: (X,Y,Z,3) f32.  Positions use the global index: pos_x = fma((x+x_offset)/(X_global-1), ...). */
int isoext_grid_points_dense(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                             const float *aabb_min, const float *aabb_max, float *d_out, void *stream);

/* ---- marching_cubes on a UniformGrid  (src/mc/mc.cu:17-68, src/isoext_ext.cu:95-109) --------
 * workspace_bytes: size of the phase-1 workspace for a slab of X*Y*Z points and at most
 * cap_entries active points.  scratch_bytes: size of the phase-2 scratch for n_candidates. */
size_t isoext_mc_dense_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t cap_entries);
size_t isoext_mc_dense_scratch_bytes(int64_t n_candidates);

/* Phase 1.  values: (X,Y,Z) f32, 32-byte aligned, read exactly once from HBM.
 * emit_x_lo/hi: local cell-x range [lo,hi) whose triangles are emitted (whole slab: 0, X-1);
 * cells outside it still decide which shared-plane vertices exist (slab halo layers).
 * counts_out[0..3] = active entries S, triangles T, vertex candidates Vc (>= welded vertex count),
 * n_big = candidates in x-plane buckets too large for the shared-memory sort (pass it to emit). */
int isoext_mc_dense_count(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float level, int method,
                          int64_t emit_x_lo, int64_t emit_x_hi, void *workspace, size_t workspace_bytes,
                          int64_t cap_entries, const void *sdf_program, void *stream, int64_t *counts_out);

/* Phase 2 (same arguments + the phase-1 workspace untouched in between).
 * V: capacity Vc x 3 f32, receives the welded vertices in the reference's order (lexicographic
 * (x,y,z), src/utils.cu:49-55).  F: T x 3 int32 ids into V, in ascending-cell x LUT order.
 * x_lo_threshold / x_hi_threshold classify V by x position for slab ownership
 * (-INFINITY / +INFINITY on a single GPU).
 * counts_out[0..2] = welded vertices, #vertices with x < lo, #vertices with x < hi. */
int isoext_mc_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                         const float *aabb_min, const float *aabb_max, float level, int method,
                         int64_t emit_x_lo, int64_t emit_x_hi, void *workspace, size_t workspace_bytes,
                         int64_t cap_entries, void *scratch, size_t scratch_bytes, int64_t n_candidates, int64_t n_big,
                         float x_lo_threshold, float x_hi_threshold, float *V, int32_t *F, const void *sdf_program, void *stream,
                         int64_t *counts_out);

/* ---- get_intersection on a UniformGrid  (src/its.cu:93-159, src/isoext_ext.cu:329-343) -------
 * The Intersection the host layer builds keeps caller-owned arrays instead of the reference's
 * (cell, edge) -> uint2 edge list: `entries` (the ordered active-point list, 8 B each), `row_start`
 * (X*Y+2 u32), `cellslot` / `its_off` (per entry: rank among active cells, CSR offset), `isout` (per
 * entry: v_lo <= v_hi of its 3 owned edges, src/its.cu:78).  points / normals are the (I,3) arrays of
 * include/its.cuh:6-26 in the same order (active cells ascending, edges 0..11).
 * Phase 1: counts_out[0..2] = entries S, active cells, intersections I. */
size_t isoext_its_dense_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t cap_entries);
int isoext_its_dense_count(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                           const float *aabb_min, const float *aabb_max, float level, void *workspace,
                           size_t workspace_bytes, int64_t cap_entries, void *entries, uint32_t *row_start,
                           uint32_t *cellslot, uint32_t *its_off, const void *sdf_program, void *stream, int64_t *counts_out);
/* Phase 2: points (I x 3), normals (I x 3; written only if compute_normals), isout (S bytes),
 * cell_offsets (n_cells+1 u32) and cell_indices (n_cells i64) as in include/its.cuh:9-12. */
int isoext_its_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float level, int compute_normals,
                          const void *entries, int64_t n_entries, const uint32_t *cellslot, const uint32_t *its_off,
                          int64_t n_cells, int64_t n_its, float *points, float *normals, unsigned char *isout,
                          uint32_t *cell_offsets, int64_t *cell_indices, const void *sdf_program, void *stream);
/* compute_intersection_normals (src/its.cu:270-284): trilinear central differences of each cell,
 * in the reference's float32 rounding pattern. */
int isoext_its_dense_normals(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                             const float *aabb_min, const float *aabb_max, const void *entries, int64_t n_entries,
                             const uint32_t *cellslot, const uint32_t *its_off, const float *points, float *normals,
                             const void *sdf_program, void *stream);

/* ---- dual_contouring on a UniformGrid  (src/dc.cu:161-218, src/isoext_ext.cu:345-378) --------
 * Phase 1: QEF (src/dc.cu:14-80) + pseudo-inverse solve (replaces src/batched_la.cu:104-179) + clip
 * (src/dc.cu:82-99) -> dual_v (n_cells x 3); quads from the owned sign-change edges (replaces
 * Grid::get_dual_quads, src/grid/uniform.cu:60-83).  counts_out[0..1] = quads Q, used dual vertices Vc.
 * Phase 2: V (capacity Vc x 3, position-sorted + welded like src/dc.cu:204-211), F (2Q x 3 int32:
 * shorter-diagonal split, src/dc.cu:139-155), optional quads_out (Q x 4 active-cell slots, oriented).
 * counts_out[0] = welded vertices. */
size_t isoext_dc_dense_workspace_bytes(int64_t n_entries, int64_t n_cells);
size_t isoext_dc_dense_scratch_bytes(int64_t n_candidates);
/* GPU sparse-grid population from a dense field: replaces the Python chunk loop of the reference's
 * recipe (tests/conftest.py:39-61 of the reference; doc/grids.ipynb runs 10,738 iterations of it at 1024^3).  Step 1 is
 * isoext_its_dense_count (entries, cellslot; counts_out[1] = crossing cells); this is step 2: the ascending cell ids
 * and the (n_cells, 8) corner values in the corner order of include/utils.cuh:32-60. */
int isoext_band_from_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                                const float *aabb_min, const float *aabb_max, const void *entries, int64_t n_entries,
                                const uint32_t *cellslot, int64_t *d_cell_idx, float *d_values8, const void *sdf_program,
                                void *stream);
/* Slab support (no reference counterpart): emit_x_lo / emit_x_hi = local point planes [lo, hi) whose
 * sign-change edges emit their quad (0 .. X on one GPU); dual vertices of ALL cells of the slab are welded, and the
 * emit phase returns how many welded vertices lie below x_lo_threshold / x_hi_threshold (ownership by position,
 * same rule as isoext_mc_dense_emit): counts_out = {V, n_lo, n_hi}.  Pass -inf / +inf on one GPU. */
int isoext_dc_dense_count(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                          const float *aabb_max, int64_t emit_x_lo, int64_t emit_x_hi, const void *entries,
                          int64_t n_entries, const uint32_t *row_start, const uint32_t *cellslot, const uint32_t *its_off,
                          int64_t n_cells, const float *points, const float *normals, float reg, float svd_tol,
                          float *dual_v, void *workspace, size_t workspace_bytes, void *stream, int64_t *counts_out);
int isoext_dc_dense_emit(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                         const float *aabb_max, int64_t emit_x_lo, int64_t emit_x_hi, float x_lo_threshold,
                         float x_hi_threshold, const void *entries, int64_t n_entries, const uint32_t *row_start,
                         const uint32_t *cellslot, const unsigned char *isout, int64_t n_cells, const float *dual_v,
                         void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes,
                         int64_t n_candidates, float *V, int32_t *F, int32_t *quads_out, void *stream,
                         int64_t *counts_out);

/* ---- SparseGrid  (src/grid/sparse.cu, src/isoext_ext.cu:170-303) ------------------------------
 * A sparse grid is a sorted list of n cell indices (int64 here; the reference's int32 API caps the grid
 * at INT_MAX points) plus (n,8) f32 corner values in the Morton corner order of include/utils.cuh:32-60.
 * X,Y,Z = points per axis of the enclosing uniform grid. */
/* get_points / get_points_by_cell_indices (src/grid/sparse.cu:21-43,144-148): out = (n,8,3) f32 */
int isoext_sparse_points(int64_t X, int64_t Y, int64_t Z, const float *aabb_min, const float *aabb_max,
                         const int64_t *cell_idx, int64_t n, float *d_out, void *stream);
/* filter_cell_indices (src/grid/sparse.cu:150-179): keep[i] = 1 iff cell i's case is not 0/255 */
int isoext_sparse_crossing(const float *values8, int64_t n, float level, unsigned char *keep, void *stream);
/* ---- cell-list maintenance and index utilities (csrc/setops.cu) ---------------------------------
 * workspace: isoext_setops_workspace_bytes(max number of ids of the call).  Each call synchronises the
 * stream once to return its count.
 * add_cells (src/grid/sparse.cu:71-97: thrust sort + unique + set_union): out = sorted unique ids of `ids`
 * (the host layer passes old list ++ new ids); out has room for n. */
size_t isoext_setops_workspace_bytes(int64_t n);
int isoext_ids_sort_unique(const int64_t *ids, int64_t n, int64_t *d_out, void *workspace, size_t workspace_bytes,
                           void *stream, int64_t *n_out);
/* remove_cells (src/grid/sparse.cu:99-126: thrust set_difference): out = a (sorted unique) minus b (any order). */
int isoext_ids_difference(const int64_t *a, int64_t na, const int64_t *b, int64_t nb, int64_t *d_out, void *workspace,
                          size_t workspace_bytes, void *stream, int64_t *n_out);
/* the copy_if of filter_cell_indices (src/grid/sparse.cu:170-178): stable compaction of 4- or 8-byte items. */
int isoext_compact_flagged(const void *src, int elem_bytes, const unsigned char *keep, int64_t n, void *d_out,
                           void *workspace, size_t workspace_bytes, void *stream, int64_t *n_out);
/* UniformGrid::get_cells (src/grid/uniform.cu:42-51, include/utils.cuh:32-60): (X-1,Y-1,Z-1,8) corner point ids,
 * uint32 (wide = 0, the reference's type) or int64 (wide = 1, needed above 2^32 points). */
int isoext_grid_cells_dense(int64_t X, int64_t Y, int64_t Z, int wide, void *d_out, void *stream);
/* welded vertices per cell layer along x (no reference counterpart: input of the slab balancer, dist.py). */
int isoext_vertex_layer_histogram(const float *V, int64_t n, float aabb_min_x, float aabb_max_x, int64_t layers,
                                  uint32_t *d_hist, void *stream);
/* marching_cubes on a SparseGrid: phase 1 counts_out[0..1] = T, Vc; phase 2 counts_out[0] = V */
size_t isoext_mc_sparse_workspace_bytes(int64_t n);
size_t isoext_sparse_scratch_bytes(int64_t n_candidates, int64_t X, int64_t Y);
/* Slab support (the sorted cell list is partitioned by x layer): only the cells [emit_begin, emit_end) of
 * the list emit triangles (0 .. n on one GPU); the vertices of ALL cells are welded, and the emit phase returns
 * counts_out = {V, # with x < x_lo_threshold, # with x < x_hi_threshold} (ownership by position). */
int isoext_mc_sparse_count(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_begin,
                           int64_t emit_end, void *workspace, size_t workspace_bytes, void *stream, int64_t *counts_out);
int isoext_mc_sparse_emit(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                          const float *aabb_min, const float *aabb_max, float level, int method, float x_lo_threshold,
                          float x_hi_threshold, void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes,
                          int64_t n_candidates, float *V, int32_t *F, void *stream, int64_t *counts_out);
/* get_intersection on a SparseGrid: cinfo / cellslot / its_off are caller-owned (n+1) u32 arrays kept by
 * the Intersection.  Phase 1 counts_out[0..1] = active cells, intersections.  emit mode: 0 points,
 * 1 points + normals, 2 normals only (compute_intersection_normals). */
int isoext_its_sparse_count(const float *values8, int64_t n, float level, uint32_t *cinfo, uint32_t *cellslot,
                            uint32_t *its_off, void *workspace, size_t workspace_bytes, void *stream, int64_t *counts_out);
int isoext_its_sparse_emit(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, float level, int mode, const uint32_t *cinfo,
                           const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, int64_t n_its, float *points,
                           float *normals, uint32_t *cell_offsets, int64_t *cell_indices, void *stream);
/* dual_contouring on a SparseGrid: neighbour cells by binary search in cell_idx (replaces the dense
 * X*Y*Z idx_map of src/grid/sparse.cu:223-243).  Phase 1 counts_out[0..1] = Q, Vc; phase 2 [0..2] = V, # with
 * x < x_lo_threshold, # with x < x_hi_threshold.
 * Slabs (new capability): only the cells [emit_begin, emit_end) of the list emit quads (0 .. n on one GPU), every
 * quad of the list marks its cells as used (the welded vertex set does not depend on the emit range), and the welded
 * vertices are classified by x position against the two thresholds (-inf / +inf on one GPU). */
size_t isoext_dc_sparse_workspace_bytes(int64_t n, int64_t X);
int isoext_dc_sparse_count(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, const uint32_t *cinfo, const uint32_t *cellslot,
                           const uint32_t *its_off, const float *points, const float *normals, float reg, float svd_tol,
                           int64_t emit_begin, int64_t emit_end, float *dual_v, void *workspace, size_t workspace_bytes,
                           void *stream, int64_t *counts_out);
int isoext_dc_sparse_emit(const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z, const float *aabb_min,
                          const float *aabb_max, const uint32_t *cinfo, const uint32_t *cellslot, const float *dual_v,
                          void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes,
                          int64_t n_candidates, float x_lo_threshold, float x_hi_threshold, float *V, int32_t *F,
                          int32_t *quads_out, void *stream, int64_t *counts_out);

/* Single-call fast path of marching_cubes on a UniformGrid: both phases enqueued back to back, ONE
 * stream synchronisation.  The caller supplies capacities (typically the sizes of the previous extraction
 * of this grid): V holds cand_cap rows, F tri_cap rows, scratch = isoext_mc_dense_scratch_bytes(cand_cap).
 * Returns 0 with counts_out[0..7] = {S, T, Vc, n_big, V, n_lo, n_hi, n_radix}; returns 1 ("not completed",
 * counts_out[0..3] and [7] valid, outputs undefined) if a capacity was exceeded -- entries, candidates,
 * triangles, or big_cap (candidates in oversized sort buckets that the second bucket level was sized for;
 * 0 = not enqueued) -- or if n_radix > 0 candidates needed the radix last resort of the sort while
 * radix == 0 (not enqueued): the caller then uses isoext_mc_dense_count + isoext_mc_dense_emit and passes
 * radix = 1 next time.
 * Halo split (slabs): if halo_event (a cudaEvent_t) is not NULL, the first halo_planes_lo and the last halo_planes_hi
 * x planes of `values` are still being filled on another stream, which records halo_event when done: the volume
 * stream over the planes in between starts immediately and only the halo planes wait for the event. */
int isoext_mc_dense_run(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                        const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_x_lo,
                        int64_t emit_x_hi, void *workspace, size_t workspace_bytes, int64_t cap_entries, void *scratch,
                        size_t scratch_bytes, int64_t cand_cap, int64_t tri_cap, int64_t big_cap, int radix,
                        float x_lo_threshold, float x_hi_threshold, int64_t halo_planes_lo, int64_t halo_planes_hi,
                        void *halo_event, float *V, int32_t *F, const void *sdf_program, void *stream,
                        int64_t *counts_out);

/* Slab-local -> global vertex ids after the per-rank counts have been all-gathered (new capability;
 * the reference is single-GPU).  id < n_lo -> base_mine - (n_lo - id); n_lo <= id < n_hi ->
 * base_mine + (id - n_lo); id >= n_hi -> base_next + (id - n_hi). */
int isoext_relabel_faces(int32_t *F, int64_t n_ids, int64_t n_lo, int64_t n_hi, int64_t base_mine,
                         int64_t base_next, void *stream);

/* ---- NVLink / NVSwitch peer transport of the slab-sharded path (no reference counterpart: the
 * reference is single-GPU; replaces NCCL send/recv + all_gather on one node) ---------------------
 * Every rank allocates its value slab and a SYNC block of isoext_peer_sync_words() u64 with
 * isoext_peer_alloc (cudaMalloc + cudaIpcGetMemHandle), ships the 64-byte handles to the other ranks
 * (any host channel) and maps theirs with isoext_peer_open.  SYNC words: [0] READY epoch, [1] DONE epoch,
 * [2 + 4*s ..] ring of 32 count slots {epoch, n_own, n_tri, -}.  err_mapped points to host-mapped
 * (pinned) memory: a kernel that waits longer than 20 s for a peer sets it to 1 and stops waiting. */
int isoext_peer_sync_words(void);
int isoext_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64);
int isoext_peer_free(void *d_ptr);
int isoext_peer_open(const unsigned char *handle64, void **d_ptr);
int isoext_peer_close(void *d_ptr);
/* *d_flag = value with release semantics at system scope, after everything enqueued before on `stream`. */
int isoext_peer_publish(uint64_t *d_flag, uint64_t value, void *stream);
/* Wait on the stream until *flag >= want for each non-null flag (peer memory). */
int isoext_peer_wait(const uint64_t *d_flag_a, const uint64_t *d_flag_b, uint64_t want, uint32_t *err_mapped, void *stream);
/* One kernel: wait until the owners of the two source segments have published READY >= epoch, then copy
 * n0 / n1 floats from their slabs (peer memory) into the local halo planes. */
int isoext_peer_halo_pull(float *d_dst0, const float *peer_src0, int64_t n0, const uint64_t *peer_ready0, float *d_dst1,
                          const float *peer_src1, int64_t n1, const uint64_t *peer_ready1, uint64_t epoch,
                          uint32_t *err_mapped, void *stream);
/* Publish this rank's (owned vertices, triangles) of `epoch` in its SYNC block. */
int isoext_peer_publish_counts(uint64_t *d_sync, uint64_t epoch, int64_t n_own, int64_t n_tri, void *stream);
/* isoext_relabel_faces with the bases computed on the device: base_mine = sum of n_own over the ranks
 * below `rank`, read from their SYNC blocks (peer_syncs: host array of >= rank device pointers).
 * d_bases_out (optional, 2 x int64): {vertex base, face base}. */
int isoext_relabel_faces_peer(int32_t *d_F, int64_t n_ids, int64_t n_lo, int64_t n_hi, int64_t n_own,
                              const uint64_t *const *peer_syncs, int rank, uint64_t epoch, int64_t *d_bases_out,
                              uint32_t *err_mapped, void *stream);

/* ---- analytic SDF programs (csrc/sdfprog.cuh) -------------------------------------
 * The reference's workflow evaluates isoext.sdf objects with torch on grid.get_points(), stores the field and extracts
 * it (src/isoext/sdf.py:69-221, tests/conftest.py:21-27).  For the built-in primitives and combinators the host layer
 * compiles the object tree to a postfix program (isoext_sdf_program_bytes() bytes of device memory, layout of
 * struct SdfProg) and every dense entry point above that takes `sdf_program` evaluates it wherever it would load a
 * value: pass values = NULL and the field never exists in HBM (the volume pass becomes a compute pass with Lipschitz
 * culling).  isoext_sdf_eval_dense materialises the field with the same device function (two-step path). */
size_t isoext_sdf_program_bytes(void);
int isoext_sdf_eval_dense(const void *sdf_program, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float *d_out, void *stream);

/* ---- gaussian_smooth (src/isoext/utils.py:5-39 is a dense k^3 conv3d) ----------------
 * Separable: three 1-D passes with clamped indices (= replicate padding).  taps_host: the k normalised 1-D weights
 * (k odd, <= 127; the host layer computes them with the reference's expressions).  field, tmp and out are three
 * distinct device buffers of X*Y*Z floats.  Agrees with the dense convolution to float32 rounding. */
int isoext_gaussian_smooth_separable(const float *field, int64_t X, int64_t Y, int64_t Z, const float *taps_host, int k,
                                     float *d_tmp, float *d_out, void *stream);

/* ---- mesh output (host code, csrc/meshio.cu) --------------------------------------
 * write_obj (src/isoext/utils.py:42-63 is a Python loop over v.tolist()): the same BYTES -- coordinates printed like
 * Python's repr(float(x)), one-based face ids -- formatted on all host cores.  v (nv x 3 f32) and f (nf x 3 i32,
 * zero-based) are HOST pointers.  An empty mesh leaves an empty file.  write_ply: binary little-endian PLY (extension). */
int isoext_write_obj(const char *path, const float *v_host, int64_t nv, const int32_t *f_host, int64_t nf);
int isoext_write_ply(const char *path, const float *v_host, int64_t nv, const int32_t *f_host, int64_t nf);

#ifdef __cplusplus
}
#endif
#endif /* ISOEXT_B200_H */
