// dc_dense.cu -- intersections, normals and dual contouring on a dense grid for sm_100a.
//
// Replaces get_intersection / compute_intersection_normals (src/its.cu:93-284), get_qef +
// BatchedLASolver::lsq_svd + fix_dual_v_op (src/dc.cu:14-99, src/batched_la.cu:104-179: cuSOLVER
// gesvdjBatched + 2x cuBLAS gemvStridedBatched + 2 thrust passes), Grid::get_dual_quads
// (src/grid/uniform.cu:60-83: sort_by_key + unique_by_key over 4x-duplicated edges) and
// get_triangles_op + welding (src/dc.cu:101-157,204-217) of the reference.
//
// The front end (sign bits -> ordered entry list + row_start) is shared with marching cubes
// (dense.cuh).  Everything after it is surface-sized:
//   k_its_scan     look-back scan over entries: cell slots + CSR offsets of the (cell,edge) crossings
//   k_its_emit     crossing points (bit-exact lerp), optional normals, is_out bits, CSR arrays
//   k_its_normals  normals only (when the caller supplied an Intersection without normals)
//   k_dc_solve     per cell: QEF accumulation in the reference's float32 rounding pattern, 3x3
//                  symmetric eigen-solve in registers (FP64 Jacobi), pseudo-inverse, clip to the cell
//   k_dc_quads     per owned sign-change edge: the 4 incident cells through row_start (each edge is
//                  owned by exactly one entry -> no sort/unique of duplicated edges), marks used cells
//   k_dc_scan      look-back scan: quad offsets, candidate ids of the used cells
//   k_dc_keys      sortable position keys of the used dual vertices
//   radix_sort96 + k_unique (weld.cuh): reference order = lexicographic positions
//                  (the layer-segmented sort of segsort.cuh was measured here: 103 vs 233 us on curved surfaces, but
//                  320 us on the 512^3 CSG of BASELINE configs[3], whose flat faces put 260 k dual vertices with
//                  almost-equal x into one layer and through the second bucket level -- the radix sort stays)
//   k_dc_faces     orientation flip, shorter-diagonal split, final ids
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "dense.cuh"
#include "radix.cuh"
#include "weld.cuh"
#include "dcmath.cuh"

namespace isx {

constexpr int IT_ITEMS = 8;
constexpr int IT_TILE = 256 * IT_ITEMS;

// ---- phase-1 workspace ------------------------------------------------------------------------
struct ItsWs {
    u32 *counters;
    u32 *bits;
    u64 *descA, *descC, *descI;
    unsigned char *span_cnt;   // entries per candidate span: count pass -> fill pass (dense.cuh), not initialised
    u32 *heavy_list;
};
static size_t carve_its_ws(Carver &c, const DenseParams &p, size_t cap, ItsWs *out) {
    ItsWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.bits = c.take<u32>(signbits_words(p.P));
    b.descA = c.take<u64>(compact_desc_count(p));
    b.descC = c.take<u64>(cap / IT_TILE + 2);
    b.descI = c.take<u64>(cap / IT_TILE + 2);
    b.span_cnt = c.take<unsigned char>(compact_span_bytes(p));
    b.heavy_list = c.take<u32>(compact_heavy_cap((u32) cap));
    if (out) *out = b;
    return c.bytes();
}

// ---------------------------------------------------------------------------------------------
// scan over entries: cellslot (rank among active-cell entries) and its_off (CSR offset)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_its_scan(u32 cap, u32 *__restrict__ counters, const uint2 *__restrict__ entries,
                                                  u32 *__restrict__ cellslot, u32 *__restrict__ its_off,
                                                  u64 *__restrict__ descC, u64 *__restrict__ descI) {
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preC, s_preI;
    const u32 S = counters[C_S];
    if (S > cap) return;
    const u32 ntiles = (S + IT_TILE - 1) / IT_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_B], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 s0 = tile * IT_TILE + threadIdx.x * IT_ITEMS;
        u32 isc[IT_ITEMS], ni[IT_ITEMS], sumC = 0, sumI = 0;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            isc[j] = 0;
            ni[j] = 0;
            if (s < S) {
                const u32 w = entries[s].y;
                if (ent_cell(w)) {
                    isc[j] = 1;
                    ni[j] = case_edge_count(ent_case(w));
                }
            }
            sumC += isc[j];
            sumI += ni[j];
        }
        u32 totC, totI;
        u32 exC = block_exclusive_scan(sumC, &totC, sw);
        u32 exI = block_exclusive_scan(sumI, &totI, sw);
        const u32 warp = threadIdx.x >> 5;
        if (warp == 0) {
            u32 pre = lookback_exclusive(descC, 1, tile, totC, 1u);
            if (threadIdx.x == 0) s_preC = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descI, 1, tile, totI, 1u);
            if ((threadIdx.x & 31) == 0) s_preI = pre;
        }
        __syncthreads();
        exC += s_preC;
        exI += s_preI;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            if (s < S) {
                cellslot[s] = isc[j] ? exC : 0xffffffffu;
                its_off[s] = exI;
                exC += isc[j];
                exI += ni[j];
                if (s == S - 1) {
                    counters[C_T] = exC;    // number of active cells
                    counters[C_I] = exI;    // number of intersections
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per active cell: crossing points in edge order 0..11 (src/its.cu:37-90), normals, is_out bits
// ---------------------------------------------------------------------------------------------
template <bool POINTS, bool NORMALS, bool IMPLICIT = false>
__global__ void __launch_bounds__(128) k_its_emit(const float *__restrict__ values, DenseParams p,
                                                  const uint2 *__restrict__ entries, u32 S, const u32 *__restrict__ cellslot,
                                                  const u32 *__restrict__ its_off, float *__restrict__ points,
                                                  float *__restrict__ normals, unsigned char *__restrict__ isout,
                                                  u32 *__restrict__ cell_offsets, i64 *__restrict__ cell_indices, u32 n_cells,
                                                  u32 n_its) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 w = e.y, r = e.x, z = ent_z(w);
        if (POINTS) {
            // is_out of the owned edges: v_lo <= v_hi (src/its.cu:78), independent of the level
            const u32 own = ent_own(w);
            u32 io = 0;
            if (own) {
                const u32 x = r / Y, y = r - x * Y;
                const float v0 = field_value<IMPLICIT>(values, p, x, y, z);
                if ((own & 1u) && v0 <= field_value<IMPLICIT>(values, p, x, y, z + 1)) io |= 1u;
                if ((own & 2u) && v0 <= field_value<IMPLICIT>(values, p, x, y + 1, z)) io |= 2u;
                if ((own & 4u) && v0 <= field_value<IMPLICIT>(values, p, x + 1, y, z)) io |= 4u;
            }
            isout[s] = (unsigned char) io;
        }
        if (!ent_cell(w)) continue;
        const u32 slot = cellslot[s];
        u32 o = its_off[s];
        if (POINTS) {
            const u32 x = r / Y, y = r - x * Y;
            cell_offsets[slot] = o;
            cell_indices[slot] = ((i64) (x + (u32) p.g.x_off) * (Y - 1) + y) * (Z - 1) + z;
            if (slot == n_cells - 1) cell_offsets[n_cells] = n_its;
        }
        CellData c;
        load_cell<IMPLICIT>(values, p, r, z, c);
        const u32 status = edge_mask_of_case(ent_case(w));
        const u32 o0 = o;
        if (POINTS) {
#pragma unroll
            for (int k = 0; k < 12; k++) {      // compile-time edge: the corner indices fold to registers
                if (!((status >> k) & 1u)) continue;
                float qx, qy, qz;
                cell_edge_point(c, k, p.level, qx, qy, qz);
                points[3 * (size_t) o + 0] = qx;
                points[3 * (size_t) o + 1] = qy;
                points[3 * (size_t) o + 2] = qz;
                o++;
            }
        }
        if (NORMALS) {
            // one rolled copy of the division-heavy normal evaluation instead of twelve inlined ones behind the edge
            // loop (the unrolled kernel stalled on instruction fetch; same change as k_sp_its_emit: 389 -> 266 us there)
            const u32 m = __popc(status);
#pragma unroll 1
            for (u32 j = 0; j < m; j++) {
                const size_t q = 3 * (size_t) (o0 + j);
                float nx, ny, nz;
                cell_normal(c, points[q], points[q + 1], points[q + 2], nx, ny, nz);
                normals[q] = nx; normals[q + 1] = ny; normals[q + 2] = nz;
            }
        }
    }
}

// GPU sparse-grid population (replaces the Python chunk loop of the reference's recipe,
// tests/conftest.py:39-61: get_potential_cell_indices -> get_points_by_cell_indices -> sdf -> filter_cell_indices ->
// add_cells -> set_values): every crossing cell of the dense field becomes a sparse cell with its 8 corner values in
// Morton corner order.  The ordered entry list already is the ascending cell list.
template <bool IMPLICIT>
__global__ void __launch_bounds__(256) k_band_emit(const float *__restrict__ values, DenseParams p, const uint2 *__restrict__ entries,
                                                   u32 S, const u32 *__restrict__ cellslot, i64 *__restrict__ cell_idx,
                                                   float *__restrict__ values8) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        if (!ent_cell(e.y)) continue;
        const u32 slot = cellslot[s], r = e.x, z = ent_z(e.y);
        const u32 x = r / Y, y = r - x * Y;
        cell_idx[slot] = ((i64) (x + (u32) p.g.x_off) * (Y - 1) + y) * (Z - 1) + z;
        CellData c;
        load_cell_values<IMPLICIT>(values, p, r, z, c);
        float4 *o = reinterpret_cast<float4 *>(values8 + 8 * (size_t) slot);
        o[0] = make_float4(c.v[0], c.v[1], c.v[2], c.v[3]);
        o[1] = make_float4(c.v[4], c.v[5], c.v[6], c.v[7]);
    }
}

// QEF + solve + clip, one thread per active cell (dcmath.cuh: qef_solve_clip).
__global__ void __launch_bounds__(128) k_dc_solve(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ cellslot, const u32 *__restrict__ its_off,
                                                  const float *__restrict__ points, const float *__restrict__ normals,
                                                  float reg, float svd_tol, float *__restrict__ dual_v) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 w = e.y;
        if (!ent_cell(w)) continue;
        const u32 r = e.x, z = ent_z(w), x = r / Y, y = r - x * Y;
        const u32 xg = x + (u32) p.g.x_off;
        const float lo[3] = {axis_pos(xg, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]), axis_pos(y, Y - 1, p.g.amin[1], p.g.asize[1]),
                             axis_pos(z, Z - 1, p.g.amin[2], p.g.asize[2])};
        const float hi[3] = {axis_pos(xg + 1, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]),
                             axis_pos(y + 1, Y - 1, p.g.amin[1], p.g.asize[1]), axis_pos(z + 1, Z - 1, p.g.amin[2], p.g.asize[2])};
        qef_solve_clip(points, normals, its_off[s], case_edge_count(ent_case(w)), reg, svd_tol, lo, hi,
                       dual_v + 3 * (size_t) cellslot[s]);
    }
}

// ---------------------------------------------------------------------------------------------
// quads: the 4 cells around an owned sign-change edge (src/shared_luts.cu:78-82 gives the cyclic
// order: cell = edge origin - offset_k).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 find_cell(const uint2 *__restrict__ entries, const u32 *__restrict__ row_start,
                                         const u32 *__restrict__ cellslot, u32 row, u32 z) {
    const u32 lo = row_start[row], hi = row_start[row + 1];
    const u32 i = row_lower_bound(entries, lo, hi, z);
    if (i >= hi || ent_z(entries[i].y) != z) return 0xffffffffu;
    return cellslot[i];   // 0xffffffff when the entry is not an active cell
}

// returns false if the quad must be skipped (a neighbour outside the cell grid or not active)
__device__ __forceinline__ bool quad_cells(const DenseParams &p, const uint2 *__restrict__ entries,
                                           const u32 *__restrict__ row_start, const u32 *__restrict__ cellslot, u32 r, u32 z,
                                           int axis, u32 q[4]) {
    const u32 X = (u32) p.g.X, Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    const u32 x = r / Y, y = r - x * Y;
    // offsets (dx,dy,dz) per k; axis: 0 = +z edge, 1 = +y edge, 2 = +x edge
    int dx[4], dy[4], dz[4];
    if (axis == 2) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 1; dz[1] = 0; dx[2] = 0; dy[2] = 1; dz[2] = 1; dx[3] = 0; dy[3] = 0; dz[3] = 1; }
    else if (axis == 1) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 0; dz[1] = 1; dx[2] = 1; dy[2] = 0; dz[2] = 1; dx[3] = 1; dy[3] = 0; dz[3] = 0; }
    else { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 1; dy[1] = 0; dz[1] = 0; dx[2] = 1; dy[2] = 1; dz[2] = 0; dx[3] = 0; dy[3] = 1; dz[3] = 0; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (x < (u32) dx[k] || y < (u32) dy[k] || z < (u32) dz[k]) return false;   // include/utils.cuh:129-131
        const u32 cx = x - dx[k], cy = y - dy[k], cz = z - dz[k];
        if (cx + 1 >= X || cy + 1 >= Y || cz + 1 >= Z) return false;               // (reference: undefined behaviour)
        q[k] = find_cell(entries, row_start, cellslot, cx * Y + cy, cz);
        if (q[k] == 0xffffffffu) return false;
    }
    return true;
}

// A dual vertex is part of the mesh iff some quad references it, i.e. iff one of its cell's sign-change edges has
// all 4 incident cells inside the cell grid (on a dense grid they are then active, because they share the edge).
// That is a function of the case and of the cell's GLOBAL position only, so every slab of a sharded grid marks
// exactly the cells the single-device run marks, whatever its halo (the reference marks through idx_map /
// get_triangles_op, src/dc.cu:184-203).  Edge e runs along `axis`; its origin is corner edge_c0(e).
__device__ __forceinline__ bool cell_is_used(const DenseParams &p, u32 xg, u32 y, u32 z, u32 cs) {
    const u32 Xg = (u32) p.g.Xg, Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    const u32 status = edge_mask_of_case(cs);
    bool used = false;
#pragma unroll
    for (int e = 0; e < 12; e++) {
        if (!((status >> e) & 1u)) continue;
        const int c0 = edge_c0(e), d = c0 ^ edge_c1(e);            // d: 4 = x edge, 2 = y edge, 1 = z edge
        const u32 ox = xg + ((c0 >> 2) & 1), oy = y + ((c0 >> 1) & 1), oz = z + (c0 & 1);
        const bool inx = ox >= 1u && ox + 2u <= Xg, iny = oy >= 1u && oy + 2u <= Y, inz = oz >= 1u && oz + 2u <= Z;
        used = used || (d == 4 ? (iny && inz) : (d == 2 ? (inx && inz) : (inx && iny)));
    }
    return used;
}

__global__ void __launch_bounds__(128) k_dc_quads(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ row_start, const u32 *__restrict__ cellslot,
                                                  unsigned char *__restrict__ qmask, unsigned char *__restrict__ used,
                                                  u32 *__restrict__ qslots) {
    const u32 Y = (u32) p.g.Y;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 own = ent_own(e.y);
        u32 m = 0;
        const u32 x = e.x / Y;
        if (ent_cell(e.y) && cellslot[s] != 0xffffffffu)
            used[cellslot[s]] = cell_is_used(p, x + (u32) p.g.x_off, e.x - x * Y, ent_z(e.y), ent_case(e.y)) ? 1 : 0;
        // quads are attributed to the slab that owns the plane of the edge's origin: local planes [emit_lo, emit_hi)
        if (own && x >= p.emit_lo && x < p.emit_hi) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (!((own >> a) & 1u)) continue;
                u32 q[4];
                if (quad_cells(p, entries, row_start, cellslot, e.x, ent_z(e.y), a, q)) {
                    m |= 1u << a;          // the face kernel reads the four slots back instead of searching again
#pragma unroll
                    for (int k = 0; k < 4; k++) qslots[(size_t) (4 * a + k) * S + s] = q[k];
                }
            }
        }
        qmask[s] = (unsigned char) m;
    }
}

// scan over entries: quad offsets + candidate ids of used cells
__global__ void __launch_bounds__(256) k_dc_scan(u32 S, u32 *__restrict__ counters, const u32 *__restrict__ cellslot,
                                                 const unsigned char *__restrict__ qmask, const unsigned char *__restrict__ used,
                                                 u32 *__restrict__ quad_off, u32 *__restrict__ cand_of_cell,
                                                 u64 *__restrict__ descQ, u64 *__restrict__ descU) {
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preQ, s_preU;
    const u32 ntiles = (S + IT_TILE - 1) / IT_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_D], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 s0 = tile * IT_TILE + threadIdx.x * IT_ITEMS;
        u32 nq[IT_ITEMS], us[IT_ITEMS], slot[IT_ITEMS], sumQ = 0, sumU = 0;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            nq[j] = 0; us[j] = 0; slot[j] = 0xffffffffu;
            if (s < S) {
                nq[j] = __popc((u32) qmask[s]);
                slot[j] = cellslot[s];
                if (slot[j] != 0xffffffffu) us[j] = used[slot[j]] ? 1u : 0u;
            }
            sumQ += nq[j];
            sumU += us[j];
        }
        u32 totQ, totU;
        u32 exQ = block_exclusive_scan(sumQ, &totQ, sw);
        u32 exU = block_exclusive_scan(sumU, &totU, sw);
        const u32 warp = threadIdx.x >> 5;
        if (warp == 0) {
            u32 pre = lookback_exclusive(descQ, 1, tile, totQ, 1u);
            if (threadIdx.x == 0) s_preQ = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descU, 1, tile, totU, 1u);
            if ((threadIdx.x & 31) == 0) s_preU = pre;
        }
        __syncthreads();
        exQ += s_preQ;
        exU += s_preU;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            if (s < S) {
                quad_off[s] = exQ;
                if (slot[j] != 0xffffffffu) cand_of_cell[slot[j]] = us[j] ? exU : 0xffffffffu;
                exQ += nq[j];
                exU += us[j];
                if (s == S - 1) {
                    counters[C_Q] = exQ;
                    counters[C_VC] = exU;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_dc_keys(u32 n_cells, const u32 *__restrict__ cand_of_cell, const float *__restrict__ dual_v,
                                                 u32 *__restrict__ kx, u32 *__restrict__ ky, u32 *__restrict__ kz) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
        const u32 id = cand_of_cell[c];
        if (id == 0xffffffffu) continue;
        kx[id] = float_key(dual_v[3 * (size_t) c + 0]);
        ky[id] = float_key(dual_v[3 * (size_t) c + 1]);
        kz[id] = float_key(dual_v[3 * (size_t) c + 2]);
    }
}

__global__ void __launch_bounds__(128) k_dc_faces(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ qslots, const u32 *__restrict__ cellslot,
                                                  const unsigned char *__restrict__ qmask, const unsigned char *__restrict__ isout,
                                                  const u32 *__restrict__ quad_off, const u32 *__restrict__ cand_of_cell,
                                                  const u32 *__restrict__ cand_rank, const float *__restrict__ dual_v,
                                                  int *__restrict__ F, int *__restrict__ quads_out) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const u32 m = qmask[s];
        if (!m) continue;
        const uint2 e = entries[s];
        const u32 io = isout[s];
        u32 qi = quad_off[s];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!((m >> a) & 1u)) continue;
            u32 q[4];
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = qslots[(size_t) (4 * a + k) * S + s];
            if (!((io >> a) & 1u)) {   // edge points inward: reverse the loop (src/dc.cu:123-127)
                const u32 t0 = q[0], t1 = q[1];
                q[0] = q[3]; q[1] = q[2]; q[2] = t1; q[3] = t0;
            }
            if (quads_out) {
                quads_out[4 * (size_t) qi + 0] = (int) q[0]; quads_out[4 * (size_t) qi + 1] = (int) q[1];
                quads_out[4 * (size_t) qi + 2] = (int) q[2]; quads_out[4 * (size_t) qi + 3] = (int) q[3];
            }
            const float *v0 = dual_v + 3 * (size_t) q[0], *v1 = dual_v + 3 * (size_t) q[1];
            const float *v2 = dual_v + 3 * (size_t) q[2], *v3 = dual_v + 3 * (size_t) q[3];
            int id[4];
#pragma unroll
            for (int k = 0; k < 4; k++) id[k] = (int) cand_rank[cand_of_cell[q[k]]];
            int *f = F + 6 * (size_t) qi;
            if (dist_ref(v0, v2) > dist_ref(v1, v3)) {   // split along v1-v3 (src/dc.cu:139-147)
                f[0] = id[1]; f[1] = id[3]; f[2] = id[0];
                f[3] = id[3]; f[4] = id[1]; f[5] = id[2];
            } else {                                      // split along v0-v2 (src/dc.cu:148-155)
                f[0] = id[2]; f[1] = id[0]; f[2] = id[1];
                f[3] = id[0]; f[4] = id[2]; f[5] = id[3];
            }
            qi++;
        }
    }
}

// ---- DC workspaces -----------------------------------------------------------------------------
struct DcWs {
    u32 *counters;
    unsigned char *qmask;   // S
    unsigned char *used;    // n_cells
    u32 *quad_off;          // S
    u32 *cand_of_cell;      // n_cells
    u32 *qslots;            // 12 planes of S: the 4 cell slots of the quad of (entry, axis), k_dc_quads -> k_dc_faces
    u64 *descQ, *descU;
};
static size_t carve_dc_ws(Carver &c, size_t S, size_t n_cells, DcWs *out) {
    DcWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.qmask = c.take<unsigned char>(S + 1);
    b.used = c.take<unsigned char>(n_cells + 1);
    b.quad_off = c.take<u32>(S + 1);
    b.cand_of_cell = c.take<u32>(n_cells + 1);
    b.descQ = c.take<u64>(S / IT_TILE + 2);
    b.descU = c.take<u64>(S / IT_TILE + 2);
    b.qslots = c.take<u32>(12 * (S + 1));
    if (out) *out = b;
    return c.bytes();
}
struct DcScratch {
    u32 *kx, *ky, *kz, *cand_rank;
    u64 *descV;
    RadixBuffers radix;
};
static size_t carve_dc_scratch(Carver &c, size_t nc, DcScratch *out) {
    DcScratch s;
    s.kx = c.take<u32>(nc);
    s.ky = c.take<u32>(nc);
    s.kz = c.take<u32>(nc);
    s.cand_rank = c.take<u32>(nc);
    s.descV = c.take<u64>(nc / UQ_TILE + 2);
    RadixBuffers::carve(c, nc, &s.radix);
    if (out) *out = s;
    return c.bytes();
}

int make_dense_params(i64 X, i64 Y, i64 Z, i64 x_off, i64 Xg, const float *amin, const float *amax, float level,
                      i64 emit_lo, i64 emit_hi, DenseParams *out);
int device_sms();

}   // namespace isx

using namespace isx;

extern "C" {

size_t isoext_its_dense_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t cap_entries) {
    DenseParams p;
    float z3[3] = {0, 0, 0}, o3[3] = {1, 1, 1};
    if (make_dense_params(X, Y, Z, 0, X, z3, o3, 0.f, 0, X - 1, &p) != OK) return 0;
    Carver c(nullptr);
    return carve_its_ws(c, p, (size_t) cap_entries, nullptr);
}

// Phase 1 of get_intersection: classify + compact + CSR scan.
//   entries (cap_entries+1 x uint2), row_start (X*Y+2 x u32), cellslot / its_off (cap_entries x u32):
//   caller-owned outputs that the Intersection keeps.
//   counts_out[0..2] = entries S, active cells, intersections I.
int isoext_its_dense_count(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                           const float *aabb_min, const float *aabb_max, float level, void *workspace, size_t workspace_bytes,
                           int64_t cap_entries, void *entries, uint32_t *row_start, uint32_t *cellslot, uint32_t *its_off,
                           const void *sdf_program, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, 0, X - 1, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (!p.sdf && (reinterpret_cast<uintptr_t>(values) & 31u) != 0) return fail(E_INVALID, "values must be 32-byte aligned");
    if (cap_entries < 1 || cap_entries >= ((i64) 1 << 29)) return fail(E_INVALID, "cap_entries out of range");
    Carver c(workspace);
    ItsWs b;
    if (carve_its_ws(c, p, (size_t) cap_entries, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const u32 cap = (u32) cap_entries;
    uint2 *ent = static_cast<uint2 *>(entries);
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descA, 0, compact_desc_count(p) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(row_start, 0, ((size_t) p.R + 2) * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descC, 0, ((size_t) cap / IT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descI, 0, ((size_t) cap / IT_TILE + 2) * sizeof(u64), stream));
    const int sms = device_sms();
    if (p.sdf) launch_sdf_bits(p, b.bits, stream);
    else launch_signbits(values, b.bits, span_sum_of(b.bits, p.P), p.P, level, stream);
    launch_compact(b.bits, p, ent, cap, row_start, b.descA, b.counters, b.span_cnt, b.heavy_list, stream);
    ISX_LAUNCH(k_its_scan, scan_blocks(sms), 256, 0, stream, cap, b.counters, ent, cellslot, its_off, b.descC, b.descI);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_S];
    counts_out[1] = h[C_T];
    counts_out[2] = h[C_I];
    if (h[C_S] > cap) return fail(E_CAPACITY, "entry capacity exceeded; retry with cap_entries >= counts_out[0]");
    return OK;
}

// Phase 2 of get_intersection: points (I x 3), normals (I x 3, only if compute_normals), is_out bits
// per entry, CSR cell_offsets (n_cells+1, u32) and cell_indices (n_cells, i64).
int isoext_its_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float level, int compute_normals, const void *entries,
                          int64_t n_entries, const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, int64_t n_its,
                          float *points, float *normals, unsigned char *isout, uint32_t *cell_offsets, int64_t *cell_indices,
                          const void *sdf_program, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, 0, X - 1, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (n_entries <= 0) return OK;
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const int sms = device_sms();
#define ISX_ITS_EMIT(N, I)                                                                                                     \
    ISX_LAUNCH((k_its_emit<true, N, I>), sms * 8, 128, 0, stream, values, p, ent, (u32) n_entries, cellslot, its_off, points, normals, \
               isout, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its)
    if (compute_normals) { if (p.sdf) ISX_ITS_EMIT(true, true); else ISX_ITS_EMIT(true, false); }
    else { if (p.sdf) ISX_ITS_EMIT(false, true); else ISX_ITS_EMIT(false, false); }
#undef ISX_ITS_EMIT
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// compute_intersection_normals (src/its.cu:270-284) for an existing Intersection.
int isoext_its_dense_normals(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                             const float *aabb_min, const float *aabb_max, const void *entries, int64_t n_entries,
                             const uint32_t *cellslot, const uint32_t *its_off, const float *points, float *normals,
                             const void *sdf_program, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X - 1, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (n_entries <= 0) return OK;
    if (p.sdf)
        ISX_LAUNCH((k_its_emit<false, true, true>), device_sms() * 8, 128, 0, stream, values, p, static_cast<const uint2 *>(entries),
                   (u32) n_entries, cellslot, its_off, const_cast<float *>(points), normals, nullptr, nullptr, nullptr, 0u, 0u);
    else
        ISX_LAUNCH((k_its_emit<false, true, false>), device_sms() * 8, 128, 0, stream, values, p, static_cast<const uint2 *>(entries),
                   (u32) n_entries, cellslot, its_off, const_cast<float *>(points), normals, nullptr, nullptr, nullptr, 0u, 0u);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// Second step of the sparse population from a dense field (first step: isoext_its_dense_count, whose counts_out[1]
// is the number of crossing cells): cell_idx (n_cells, ascending) and values8 (n_cells x 8).
int isoext_band_from_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                                const float *aabb_min, const float *aabb_max, const void *entries, int64_t n_entries,
                                const uint32_t *cellslot, int64_t *cell_idx, float *values8, const void *sdf_program, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (n_entries <= 0) return OK;
    if ((reinterpret_cast<uintptr_t>(values8) & 15u) != 0) return fail(E_INVALID, "values8 must be 16-byte aligned");
    if (p.sdf)
        ISX_LAUNCH(k_band_emit<true>, device_sms() * 8, 256, 0, stream, values, p, static_cast<const uint2 *>(entries), (u32) n_entries,
                   cellslot, cell_idx, values8);
    else
        ISX_LAUNCH(k_band_emit<false>, device_sms() * 8, 256, 0, stream, values, p, static_cast<const uint2 *>(entries), (u32) n_entries,
                   cellslot, cell_idx, values8);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// ---- analytic programs (sdfprog.cuh) ---------------------------------------------------------------
size_t isoext_sdf_program_bytes(void) { return sizeof(SdfProg); }
// Materialise the field of a program on a grid / slab: out = (X, Y, Z) f32.  The two-step path, and the reference point of the
// fused path (same device function).
int isoext_sdf_eval_dense(const void *sdf_program, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!sdf_program) return fail(E_INVALID, "no program");
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    i64 want = (p.P + 255) / 256;
    const i64 cap = (i64) device_sms() * 32;
    ISX_LAUNCH(k_sdf_fill, (int) (want > cap ? cap : want), 256, 0, stream, p, out);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

size_t isoext_dc_dense_workspace_bytes(int64_t n_entries, int64_t n_cells) {
    Carver c(nullptr);
    return carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, nullptr);
}
size_t isoext_dc_dense_scratch_bytes(int64_t n_candidates) {
    Carver c(nullptr);
    return carve_dc_scratch(c, (size_t) (n_candidates > 0 ? n_candidates : 1), nullptr);
}

// Phase 1 of dual_contouring: dual vertices (n_cells x 3, clipped) + quad / candidate counts.
//   emit_x_lo / emit_x_hi: local point planes [lo, hi) whose sign-change edges emit their quad (0 .. X on one GPU;
//   a slab emits the planes it owns and welds the dual vertices of its ghost layers as well).
//   counts_out[0..1] = quads Q, used dual vertices Vc.
int isoext_dc_dense_count(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                          const float *aabb_max, int64_t emit_x_lo, int64_t emit_x_hi, const void *entries, int64_t n_entries,
                          const uint32_t *row_start,
                          const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, const float *points,
                          const float *normals, float reg, float svd_tol, float *dual_v, void *workspace,
                          size_t workspace_bytes, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, emit_x_lo, emit_x_hi, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = 0;
    if (n_entries <= 0 || n_cells <= 0) return OK;
    Carver c(workspace);
    DcWs b;
    if (carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const u32 S = (u32) n_entries;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descQ, 0, ((size_t) S / IT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descU, 0, ((size_t) S / IT_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_dc_solve, sms * 8, 128, 0, stream, p, ent, S, cellslot, its_off, points, normals, reg, svd_tol, dual_v);
    ISX_LAUNCH(k_dc_quads, sms * 8, 128, 0, stream, p, ent, S, row_start, cellslot, b.qmask, b.used, b.qslots);
    ISX_LAUNCH(k_dc_scan, scan_blocks(sms), 256, 0, stream, S, b.counters, cellslot, b.qmask, b.used, b.quad_off, b.cand_of_cell, b.descQ, b.descU);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_Q];
    counts_out[1] = h[C_VC];
    return OK;
}

// Phase 2: V (capacity Vc x 3), F (2Q x 3 int32), optional quads_out (Q x 4 cell slots, oriented).
//   counts_out[0..2] = welded vertices, # with x < x_lo_threshold, # with x < x_hi_threshold (slab ownership; pass
//   -inf / +inf on one GPU).
int isoext_dc_dense_emit(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                         const float *aabb_max, int64_t emit_x_lo, int64_t emit_x_hi, float x_lo_threshold, float x_hi_threshold,
                         const void *entries, int64_t n_entries, const uint32_t *row_start,
                         const uint32_t *cellslot, const unsigned char *isout, int64_t n_cells, const float *dual_v,
                         void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes, int64_t n_candidates,
                         float *V, int32_t *F, int32_t *quads_out, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, emit_x_lo, emit_x_hi, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = counts_out[2] = 0;
    if (n_candidates <= 0) return OK;
    Carver c(workspace);
    DcWs b;
    if (carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    Carver cs(scratch);
    DcScratch s;
    if (carve_dc_scratch(cs, (size_t) n_candidates, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const u32 S = (u32) n_entries, nc = (u32) n_candidates;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(s.descV, 0, ((size_t) nc / UQ_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_dc_keys, sms * 8, 256, 0, stream, (u32) n_cells, b.cand_of_cell, dual_v, s.kx, s.ky, s.kz);
    ISX_CUDA(radix_sort96(s.kx, s.ky, s.kz, nc, s.radix, stream));
    ISX_LAUNCH(k_unique, scan_blocks(sms), 256, 0, stream, nc, s.radix.perm[0], s.kx, s.ky, s.kz, s.cand_rank, V, b.counters, s.descV,
               host_float_key(x_lo_threshold), host_float_key(x_hi_threshold));
    ISX_LAUNCH(k_dc_faces, sms * 8, 128, 0, stream, p, ent, S, b.qslots, cellslot, b.qmask, isout, b.quad_off, b.cand_of_cell,
               s.cand_rank, dual_v, F, quads_out);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_V];
    counts_out[1] = h[C_NLO];
    counts_out[2] = h[C_NHI];
    return OK;
}

}   // extern "C"
