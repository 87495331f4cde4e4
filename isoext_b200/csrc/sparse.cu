// sparse.cu -- SparseGrid kernels for sm_100a: corner positions, crossing-cell filter, marching cubes,
// intersections and dual contouring over a sorted list of active cells with (N,8) corner values.
//
// Replaces the reference's SparseGrid paths (src/grid/sparse.cu:21-69,128-246 + the generic MC/DC
// drivers): no (N,8) `cells` iota, no (N,8,3) `points` materialisation, and -- for dual contouring --
// no dense X*Y*Z `idx_map` (src/grid/sparse.cu:223): neighbour cells are found by binary search in the
// sorted cell list.  Cell indices are int64 on this side of the ABI (the reference's int32 API caps
// the grid at INT_MAX points; the host layer accepts both).
//
// A sparse cell carries its own 8 corner values, which need not agree with its neighbours', so
// vertices are welded by position only (exactly the reference's semantics): every (cell, edge) vertex
// that a kept triangle references becomes a sort candidate.
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "dense.cuh"
#include "radix.cuh"
#include "weld.cuh"
#include "segsort.cuh"
#include "dcmath.cuh"

#define ISX_LUT_QUAL static __device__
#include "mc_luts.inc"

namespace isx {

struct SparseParams {
    Geom g;          // X,Y,Z = points per axis of the enclosing uniform grid
    i64 cy, cz;      // cells per axis along y and z  (Y-1, Z-1)
    float level;
    float eps2[3];   // plain-cell test of marching cubes (dense.cuh: cell_is_plain)
};

__device__ __forceinline__ void cell_coords(const SparseParams &p, i64 idx, u32 &x, u32 &y, u32 &z) {
    z = (u32) (idx % p.cz);
    const i64 t = idx / p.cz;
    y = (u32) (t % p.cy);
    x = (u32) (t / p.cy);
}

__device__ __forceinline__ void load_sparse_cell(const float *__restrict__ values8, const SparseParams &p, u32 s, i64 idx, CellData &c) {
    const float4 lo = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s));
    const float4 hi = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s) + 1);
    c.v[0] = lo.x; c.v[1] = lo.y; c.v[2] = lo.z; c.v[3] = lo.w;
    c.v[4] = hi.x; c.v[5] = hi.y; c.v[6] = hi.z; c.v[7] = hi.w;
    u32 x, y, z;
    cell_coords(p, idx, x, y, z);
    c.px[0] = axis_pos(x, (u32) p.g.X - 1, p.g.amin[0], p.g.asize[0]);
    c.px[1] = axis_pos(x + 1, (u32) p.g.X - 1, p.g.amin[0], p.g.asize[0]);
    c.py[0] = axis_pos(y, (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]);
    c.py[1] = axis_pos(y + 1, (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]);
    c.pz[0] = axis_pos(z, (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2]);
    c.pz[1] = axis_pos(z + 1, (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2]);
}

__device__ __forceinline__ u32 case_of_values(const float *v, float level) {
    u32 c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c |= (u32) (__fsub_rn(v[i], level) < 0.0f) << i;
    return c;
}

// ---- get_points / get_points_by_cell_indices (src/grid/sparse.cu:21-36) ---------------------------
__global__ void __launch_bounds__(256) k_sp_points(SparseParams p, const i64 *__restrict__ cell_idx, i64 n, float *__restrict__ out) {
    for (i64 t = (i64) blockIdx.x * blockDim.x + threadIdx.x; t < n * 8; t += (i64) gridDim.x * blockDim.x) {
        const i64 s = t >> 3;
        const u32 k = (u32) t & 7u;
        u32 x, y, z;
        cell_coords(p, cell_idx[s], x, y, z);
        float *o = out + 3 * t;
        o[0] = axis_pos(x + ((k >> 2) & 1u), (u32) p.g.X - 1, p.g.amin[0], p.g.asize[0]);
        o[1] = axis_pos(y + ((k >> 1) & 1u), (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]);
        o[2] = axis_pos(z + (k & 1u), (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2]);
    }
}

// ---- filter_cell_indices (src/grid/sparse.cu:150-179): 1 where the cell crosses the level ---------
__global__ void __launch_bounds__(256) k_sp_crossing(const float *__restrict__ values8, i64 n, float level, unsigned char *__restrict__ keep) {
    for (i64 s = (i64) blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (i64) gridDim.x * blockDim.x) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * s));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * s) + 1);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const u32 c = case_of_values(v, level);
        keep[s] = (c != 0u && c != 255u) ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// marching cubes
// ---------------------------------------------------------------------------------------------
constexpr int SPT_ITEMS = 8;
constexpr int SPT_TILE = 256 * SPT_ITEMS;

// cinfo = case | trimask << 8 | usedmask << 16
// Cells outside [emit_begin, emit_end) (the ghost cells of a slab) keep their used-edge mask -- their vertices are
// welded so that the slab can number by position -- but emit no triangle.
__global__ void __launch_bounds__(128) k_sp_mc_classify(const float *__restrict__ values8, const i64 *__restrict__ cell_idx,
                                                        SparseParams p, u32 n, int method, u32 *__restrict__ cinfo, u32 emit_begin,
                                                        u32 emit_end) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        CellData c;
        load_sparse_cell(values8, p, s, cell_idx[s], c);
        const u32 cs = case_of_values(c.v, p.level);
        if (cs == 0u || cs == 255u) {
            cinfo[s] = cs;
            continue;
        }
        const u32 status = edge_mask_of_case(cs);
        const u64 word = method == 0 ? kTriWords_nagae[cs] : kTriWords_lorensen[cs];
        const u32 nt = (u32) (word >> 60);
        u32 mask = 0, used = 0;
        if (cell_is_plain(c, status, p.level, p.eps2)) {
            // no crossing is near a cell corner, so no two crossing points can be bit-equal: every LUT triangle is kept and
            // every sign-change edge is used, without evaluating a position (the fast path of the dense k_cell_tris)
            mask = (1u << nt) - 1u;
            used = status;
            if (s < emit_begin || s >= emit_end) mask = 0;
            cinfo[s] = cs | (mask << 8) | (used << 16);
            continue;
        }
        float ex[12], ey[12], ez[12];
#pragma unroll
        for (int k = 0; k < 12; k++)
            if ((status >> k) & 1u) cell_edge_point(c, k, p.level, ex[k], ey[k], ez[k]);
        for (u32 k = 0; k < nt; k++) {
            const u32 a = (u32) (word >> (12 * k)) & 15u, b = (u32) (word >> (12 * k + 4)) & 15u, d = (u32) (word >> (12 * k + 8)) & 15u;
            const bool ab = ex[a] != ex[b] || ey[a] != ey[b] || ez[a] != ez[b];
            const bool ad = ex[a] != ex[d] || ey[a] != ey[d] || ez[a] != ez[d];
            const bool bd = ex[b] != ex[d] || ey[b] != ey[d] || ez[b] != ez[d];
            if (ab && ad && bd) {
                mask |= 1u << k;
                used |= (1u << a) | (1u << b) | (1u << d);
            }
        }
        if (s < emit_begin || s >= emit_end) mask = 0;
        cinfo[s] = cs | (mask << 8) | (used << 16);
    }
}

// Candidate de-duplication along z.  Every (cell, edge) vertex is a candidate of the position sort because cells carry
// their own corner values; but the list is sorted by cell id, so the cell below (x, y, z-1) -- if present -- is the
// PREVIOUS list element, and it shares the four edges of this cell's low-z face (this cell's e3, e7, e8, e11 are its
// e1, e5, e9, e10: same end points, same direction, hence the same arithmetic).  Where the two cells hold bit-equal
// corner values on such an edge and the lower cell emits it, this cell refers to the lower cell's candidate instead
// of emitting a bit-identical one: a third fewer keys to sort.  Cells with different values keep their own vertex
// (positional welding decides, as in the reference).
//   cinfo after this pass = case | trimask << 8 | EMITTED edges << 16 | dup4 << 28   (dup4: e3, e7, e8, e11)
__device__ __forceinline__ u32 sp_dup_edges(u32 dup4) {
    return ((dup4 & 1u) << 3) | ((dup4 & 2u) << 6) | ((dup4 & 4u) << 6) | ((dup4 & 8u) << 8);
}
__global__ void __launch_bounds__(256) k_sp_mc_dedupe(const float *__restrict__ values8, const i64 *__restrict__ cell_idx, SparseParams p,
                                                      u32 n, u32 *__restrict__ cinfo) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        if (s == 0) continue;
        const u32 w = cinfo[s];
        const u32 used = (w >> 16) & 0xfffu;
        if (!(used & 0x988u)) continue;                             // none of e3, e7, e8, e11 in use
        const i64 id = cell_idx[s];
        if (cell_idx[s - 1] + 1 != id || id % p.cz == 0) continue;  // previous element is not the cell below in the same row
        const u32 usedp = (cinfo[s - 1] >> 16) & 0xfffu;            // (its low-face bits may change concurrently, e1/e5/e9/e10 never do)
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s) + 1);
        const float4 plo = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) (s - 1)));
        const float4 phi = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) (s - 1)) + 1);
        // bit-equality of my corners 0, 2, 4, 6 with the lower cell's corners 1, 3, 5, 7
        const bool c0 = __float_as_uint(lo.x) == __float_as_uint(plo.y), c2 = __float_as_uint(lo.z) == __float_as_uint(plo.w);
        const bool c4 = __float_as_uint(hi.x) == __float_as_uint(phi.y), c6 = __float_as_uint(hi.z) == __float_as_uint(phi.w);
        u32 dup4 = 0;
        if (((used >> 3) & 1u) && ((usedp >> 1) & 1u) && c0 && c2) dup4 |= 1u;    // e3  (0-2) == lower e1  (1-3)
        if (((used >> 7) & 1u) && ((usedp >> 5) & 1u) && c4 && c6) dup4 |= 2u;    // e7  (4-6) == lower e5  (5-7)
        if (((used >> 8) & 1u) && ((usedp >> 9) & 1u) && c0 && c4) dup4 |= 4u;    // e8  (0-4) == lower e9  (1-5)
        if (((used >> 11) & 1u) && ((usedp >> 10) & 1u) && c2 && c6) dup4 |= 8u;  // e11 (2-6) == lower e10 (3-7)
        if (dup4) cinfo[s] = (w & ~(sp_dup_edges(dup4) << 16)) | (dup4 << 28);
    }
}

// scan over cells of (kept triangles, used edges)
__global__ void __launch_bounds__(256) k_sp_scan2(u32 n, u32 *__restrict__ counters, const u32 *__restrict__ cinfo, int shiftA,
                                                  u32 maskA, int shiftB, u32 maskB, u32 *__restrict__ offA, u32 *__restrict__ offB,
                                                  u64 *__restrict__ descA, u64 *__restrict__ descB, int cntA, int cntB, int ticket) {
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preA, s_preB;
    const u32 ntiles = (n + SPT_TILE - 1) / SPT_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[ticket], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 s0 = tile * SPT_TILE + threadIdx.x * SPT_ITEMS;
        u32 a[SPT_ITEMS], b[SPT_ITEMS], sumA = 0, sumB = 0;
#pragma unroll
        for (int j = 0; j < SPT_ITEMS; j++) {
            const u32 s = s0 + j;
            a[j] = 0; b[j] = 0;
            if (s < n) {
                const u32 w = cinfo[s];
                a[j] = __popc((w >> shiftA) & maskA);
                b[j] = __popc((w >> shiftB) & maskB);
            }
            sumA += a[j];
            sumB += b[j];
        }
        u32 totA, totB;
        u32 exA = block_exclusive_scan(sumA, &totA, sw);
        u32 exB = block_exclusive_scan(sumB, &totB, sw);
        const u32 warp = threadIdx.x >> 5;
        if (warp == 0) {
            u32 pre = lookback_exclusive(descA, 1, tile, totA, 1u);
            if (threadIdx.x == 0) s_preA = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descB, 1, tile, totB, 1u);
            if ((threadIdx.x & 31) == 0) s_preB = pre;
        }
        __syncthreads();
        exA += s_preA;
        exB += s_preB;
#pragma unroll
        for (int j = 0; j < SPT_ITEMS; j++) {
            const u32 s = s0 + j;
            if (s < n) {
                offA[s] = exA;
                offB[s] = exB;
                exA += a[j];
                exB += b[j];
                if (s == n - 1) {
                    counters[cntA] = exA;
                    counters[cntB] = exB;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) k_sp_mc_keys(const float *__restrict__ values8, const i64 *__restrict__ cell_idx,
                                                    SparseParams p, u32 n, const u32 *__restrict__ cinfo,
                                                    const u32 *__restrict__ cand_off, u32 *__restrict__ kx, u32 *__restrict__ ky,
                                                    u32 *__restrict__ kz) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 used = (cinfo[s] >> 16) & 0xfffu;   // the edges this cell emits (k_sp_mc_dedupe)
        if (!used) continue;
        CellData c;
        load_sparse_cell(values8, p, s, cell_idx[s], c);
        u32 id = cand_off[s];
#pragma unroll
        for (int k = 0; k < 12; k++) {
            if (!((used >> k) & 1u)) continue;
            float qx, qy, qz;
            cell_edge_point(c, k, p.level, qx, qy, qz);
            kx[id] = float_key(qx);
            ky[id] = float_key(qy);
            kz[id] = float_key(qz);
            id++;
        }
    }
}

__global__ void __launch_bounds__(256) k_sp_mc_faces(u32 n, int method, const u32 *__restrict__ cinfo, const u32 *__restrict__ tri_off,
                                                     const u32 *__restrict__ cand_off, const u32 *__restrict__ cand_rank,
                                                     int *__restrict__ F) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 w = cinfo[s];
        const u32 mask = (w >> 8) & 0xffu;
        if (!mask) continue;
        const u32 cs = w & 0xffu, emitted = (w >> 16) & 0xfffu, base = cand_off[s];
        const u32 dup = sp_dup_edges(w >> 28);             // edges whose vertex is the lower cell's candidate (k_sp_mc_dedupe)
        u32 pemit = 0, pbase = 0;
        if (dup) {
            pemit = (cinfo[s - 1] >> 16) & 0xfffu;
            pbase = cand_off[s - 1];
        }
        const u64 word = method == 0 ? kTriWords_nagae[cs] : kTriWords_lorensen[cs];
        const u32 nt = (u32) (word >> 60);
        size_t o = 3 * (size_t) tri_off[s];
        for (u32 k = 0; k < nt; k++) {
            if (!((mask >> k) & 1u)) continue;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const u32 e = (u32) (word >> (12 * k + 4 * j)) & 15u;
                u32 cand;
                if ((dup >> e) & 1u) {
                    const u32 ep = e == 3u ? 1u : (e == 7u ? 5u : (e == 8u ? 9u : 10u));   // the same edge in the lower cell
                    cand = pbase + __popc(pemit & ((1u << ep) - 1u));
                } else {
                    cand = base + __popc(emitted & ((1u << e) - 1u));
                }
                F[o + j] = (int) cand_rank[cand];
            }
            o += 3;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// intersections (src/its.cu:93-159 on a SparseGrid) -- active = cells with a sign-change edge
// ---------------------------------------------------------------------------------------------
// cinfo for its/DC = case (8 bits); cells with case 0/255 are inactive.
__global__ void __launch_bounds__(256) k_sp_case(const float *__restrict__ values8, u32 n, float level, u32 *__restrict__ cinfo) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s));
        const float4 hi = __ldg(reinterpret_cast<const float4 *>(values8 + 8 * (size_t) s) + 1);
        const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        const u32 c = case_of_values(v, level);
        // bit 8 marks "active"; bits 16.. hold the sign-change edge mask (popcount = #intersections)
        cinfo[s] = (c != 0u && c != 255u) ? (c | 0x100u | (edge_mask_of_case(c) << 16)) : c;
    }
}

template <bool POINTS, bool NORMALS>
__global__ void __launch_bounds__(128) k_sp_its_emit(const float *__restrict__ values8, const i64 *__restrict__ cell_idx,
                                                     SparseParams p, u32 n, const u32 *__restrict__ cinfo,
                                                     const u32 *__restrict__ cellslot, const u32 *__restrict__ its_off,
                                                     float *__restrict__ points, float *__restrict__ normals,
                                                     u32 *__restrict__ cell_offsets, i64 *__restrict__ cell_indices, u32 n_cells,
                                                     u32 n_its) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 w = cinfo[s];
        if (!(w & 0x100u)) continue;
        const u32 slot = cellslot[s];
        u32 o = its_off[s];
        if (POINTS) {
            cell_offsets[slot] = o;
            cell_indices[slot] = (i64) s;   // the reference's sparse cell_indices are slots of the sparse list
            if (slot == n_cells - 1) cell_offsets[n_cells] = n_its;
        }
        CellData c;
        load_sparse_cell(values8, p, s, cell_idx[s], c);
        const u32 status = (w >> 16) & 0xfffu;
        const u32 o0 = o;
        if (POINTS) {
#pragma unroll
            for (int k = 0; k < 12; k++) {      // the edge is a compile-time constant here (corner indices fold to registers)
                if (!((status >> k) & 1u)) continue;
                float qx, qy, qz;
                cell_edge_point(c, k, p.level, qx, qy, qz);
                points[3 * (size_t) o] = qx; points[3 * (size_t) o + 1] = qy; points[3 * (size_t) o + 2] = qz;
                o++;
            }
        }
        if (NORMALS) {
            // One copy of the (division-heavy, ~300 instructions) normal evaluation in a rolled loop over the cell's
            // intersections, instead of twelve inlined copies behind the unrolled edge loop: the kernel was stalling on
            // instruction fetch (ncu: no_inst among the top stall reasons on every hot line).
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

// ---------------------------------------------------------------------------------------------
// dual contouring over the sparse list
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_sp_dc_solve(const i64 *__restrict__ cell_idx, SparseParams p, u32 n,
                                                     const u32 *__restrict__ cinfo, const u32 *__restrict__ cellslot,
                                                     const u32 *__restrict__ its_off, const float *__restrict__ points,
                                                     const float *__restrict__ normals, float reg, float svd_tol,
                                                     float *__restrict__ dual_v) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 w = cinfo[s];
        if (!(w & 0x100u)) continue;
        u32 x, y, z;
        cell_coords(p, cell_idx[s], x, y, z);
        const float lo[3] = {axis_pos(x, (u32) p.g.X - 1, p.g.amin[0], p.g.asize[0]), axis_pos(y, (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]),
                             axis_pos(z, (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2])};
        const float hi[3] = {axis_pos(x + 1, (u32) p.g.X - 1, p.g.amin[0], p.g.asize[0]),
                             axis_pos(y + 1, (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]),
                             axis_pos(z + 1, (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2])};
        qef_solve_clip(points, normals, its_off[s], __popc(w >> 16), reg, svd_tol, lo, hi, dual_v + 3 * (size_t) cellslot[s]);
    }
}

// slot (in the sparse list) of the cell with index `key`, or 0xffffffff
__device__ __forceinline__ u32 find_sparse_cell(const i64 *__restrict__ cell_idx, u32 n, i64 key) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (cell_idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < n && cell_idx[lo] == key) ? lo : 0xffffffffu;
}
// same, inside the list range [lo, hi) (one x plane of cells: plane_start[x] .. plane_start[x + 1])
__device__ __forceinline__ u32 find_sparse_cell_in(const i64 *__restrict__ cell_idx, u32 lo, u32 hi, i64 key) {
    const u32 end = hi;
    while (lo < hi) {
        const u32 mid = (lo + hi) >> 1;
        if (cell_idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < end && cell_idx[lo] == key) ? lo : 0xffffffffu;
}
// same, for a key below cell_idx[s] that lies a short way back in the list (a neighbour in the same x plane: one cell
// or one row back): gallop backwards from s, then bisect the last stride
__device__ __forceinline__ u32 find_sparse_cell_back(const i64 *__restrict__ cell_idx, u32 s, i64 key) {
    u32 hi = s, step = 1;          // invariant: cell_idx[hi] > key
    u32 lo;
    while (true) {
        if (step > hi) { lo = 0; break; }
        const u32 probe = hi - step;
        const i64 v = cell_idx[probe];
        if (v == key) return probe;
        if (v < key) { lo = probe + 1; break; }
        hi = probe;
        step <<= 1;
    }
    return find_sparse_cell_in(cell_idx, lo, hi, key);
}
// first list position of every x plane of cells: plane_start[x] = lower_bound(cell_idx, x * cy * cz), x = 0 .. X-1
// (X - 1 cell planes; the last entry is n)
static __global__ void __launch_bounds__(256) k_sp_plane_start(const i64 *__restrict__ cell_idx, u32 n, SparseParams p,
                                                               u32 *__restrict__ plane_start) {
    const u32 planes = (u32) p.g.X;   // entries 0 .. X-1
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < planes; x += gridDim.x * blockDim.x) {
        const i64 key = (i64) x * p.cy * p.cz;
        u32 lo = 0, hi = n;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if (cell_idx[mid] < key) lo = mid + 1; else hi = mid;
        }
        plane_start[x] = lo;
    }
}

// The edge (p_lo, axis) is handled by the cell whose corner 0 is p_lo (its edges e0 = +z, e3 = +y,
// e8 = +x): that cell is always one of the 4 cells around the edge, so if it is absent the reference
// skips the quad anyway.  Quad order = sorted (p_lo, p_hi) = ascending cell, then +z, +y, +x.
// q[k] = sparse-list slots of the 4 cells in the reference's cyclic order; false -> skip.
__device__ __forceinline__ bool sparse_quad_cells(const SparseParams &p, const i64 *__restrict__ cell_idx, u32 n,
                                                  const u32 *__restrict__ cinfo, const u32 *__restrict__ plane_start, u32 s, int axis,
                                                  u32 q[4]) {
    u32 x, y, z;
    cell_coords(p, cell_idx[s], x, y, z);
    int dx[4], dy[4], dz[4];
    if (axis == 2) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 1; dz[1] = 0; dx[2] = 0; dy[2] = 1; dz[2] = 1; dx[3] = 0; dy[3] = 0; dz[3] = 1; }
    else if (axis == 1) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 0; dz[1] = 1; dx[2] = 1; dy[2] = 0; dz[2] = 1; dx[3] = 1; dy[3] = 0; dz[3] = 0; }
    else { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 1; dy[1] = 0; dz[1] = 0; dx[2] = 1; dy[2] = 1; dz[2] = 0; dx[3] = 0; dy[3] = 1; dz[3] = 0; }
    q[0] = s;
    // the neighbours of the previous x plane lie inside [plane_start[x-1], plane_start[x]): 11 probes instead of 21 at
    // 2.4 M cells; those of the own plane are one cell or one row back: a short gallop from s
    u32 pl = 0, ph = 0;
    if (x >= 1 && axis != 2) { pl = plane_start[x - 1]; ph = plane_start[x]; }
#pragma unroll
    for (int k = 1; k < 4; k++) {
        if (x < (u32) dx[k] || y < (u32) dy[k] || z < (u32) dz[k]) return false;
        const i64 key = ((i64) (x - dx[k]) * p.cy + (y - dy[k])) * p.cz + (z - dz[k]);
        const u32 t = dx[k] ? find_sparse_cell_in(cell_idx, pl, ph, key) : find_sparse_cell_back(cell_idx, s, key);
        if (t == 0xffffffffu || !(cinfo[t] & 0x100u)) return false;
        q[k] = t;
    }
    return true;
}

// sign-change flags of the three edges at corner 0: e0 (v0-v1, +z), e3 (v0-v2, +y), e8 (v0-v4, +x)
__device__ __forceinline__ u32 own_edges_of_case(u32 cs) {
    const u32 b0 = cs & 1u;
    return (b0 ^ ((cs >> 1) & 1u)) | ((b0 ^ ((cs >> 2) & 1u)) << 1) | ((b0 ^ ((cs >> 4) & 1u)) << 2);
}

// dinfo = qmask (3 bits) | isout (3 bits) << 4
__global__ void __launch_bounds__(128) k_sp_dc_quads(const float *__restrict__ values8, const i64 *__restrict__ cell_idx,
                                                     SparseParams p, u32 n, const u32 *__restrict__ cinfo,
                                                     const u32 *__restrict__ plane_start, u32 *__restrict__ dinfo,
                                                     unsigned char *__restrict__ used, u32 emit_begin, u32 emit_end,
                                                     u32 *__restrict__ qnb) {
    // Every quad of the list marks its four cells as used (so that the set of welded dual vertices of a slab does not
    // depend on which quads the slab emits); only the cells [emit_begin, emit_end) keep their quads (the owned cells
    // of a slab; 0 .. n on one GPU).
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 w = cinfo[s];
        const bool emits = s >= emit_begin && s < emit_end;
        u32 m = 0, io = 0;
        if (w & 0x100u) {
            const u32 own = own_edges_of_case(w & 0xffu);
            if (own) {
                const float v0 = values8[8 * (size_t) s];
                if (v0 <= values8[8 * (size_t) s + 1]) io |= 1u;   // +z: corner 1
                if (v0 <= values8[8 * (size_t) s + 2]) io |= 2u;   // +y: corner 2
                if (v0 <= values8[8 * (size_t) s + 4]) io |= 4u;   // +x: corner 4
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    if (!((own >> a) & 1u)) continue;
                    u32 q[4];
                    if (sparse_quad_cells(p, cell_idx, n, cinfo, plane_start, s, a, q)) {
                        if (emits) {        // the face kernel reads the three neighbours back instead of searching again
                            m |= 1u << a;
                            qnb[(size_t) (3 * a + 0) * n + s] = q[1];
                            qnb[(size_t) (3 * a + 1) * n + s] = q[2];
                            qnb[(size_t) (3 * a + 2) * n + s] = q[3];
                        }
                        used[q[0]] = 1; used[q[1]] = 1; used[q[2]] = 1; used[q[3]] = 1;
                    }
                }
            }
        }
        dinfo[s] = m | (io << 4);
    }
}

// fold the used flags into dinfo (bit 8) so that the generic two-counter scan can run on it
__global__ void __launch_bounds__(256) k_sp_fold_used(u32 n, const unsigned char *__restrict__ used, u32 *__restrict__ dinfo) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x)
        if (used[s]) dinfo[s] |= 0x100u;
}

__global__ void __launch_bounds__(256) k_sp_dc_keys(u32 n, const u32 *__restrict__ dinfo, const u32 *__restrict__ cand_off,
                                                    const u32 *__restrict__ cellslot, const float *__restrict__ dual_v,
                                                    u32 *__restrict__ kx, u32 *__restrict__ ky, u32 *__restrict__ kz) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        if (!(dinfo[s] & 0x100u)) continue;
        const u32 id = cand_off[s];
        const float *v = dual_v + 3 * (size_t) cellslot[s];
        kx[id] = float_key(v[0]);
        ky[id] = float_key(v[1]);
        kz[id] = float_key(v[2]);
    }
}

__global__ void __launch_bounds__(128) k_sp_dc_faces(const i64 *__restrict__ cell_idx, SparseParams p, u32 n,
                                                     const u32 *__restrict__ cinfo, const u32 *__restrict__ qnb,
                                                     const u32 *__restrict__ dinfo,
                                                     const u32 *__restrict__ quad_off, const u32 *__restrict__ cand_off,
                                                     const u32 *__restrict__ cellslot, const u32 *__restrict__ cand_rank,
                                                     const float *__restrict__ dual_v, int *__restrict__ F, int *__restrict__ quads_out) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        const u32 di = dinfo[s], m = di & 7u, io = (di >> 4) & 7u;
        if (!m) continue;
        u32 qi = quad_off[s];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!((m >> a) & 1u)) continue;
            u32 q[4] = {s, qnb[(size_t) (3 * a + 0) * n + s], qnb[(size_t) (3 * a + 1) * n + s], qnb[(size_t) (3 * a + 2) * n + s]};
            if (!((io >> a) & 1u)) {
                const u32 t0 = q[0], t1 = q[1];
                q[0] = q[3]; q[1] = q[2]; q[2] = t1; q[3] = t0;
            }
            const float *v[4];
            int id[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                v[k] = dual_v + 3 * (size_t) cellslot[q[k]];
                id[k] = (int) cand_rank[cand_off[q[k]]];
                if (quads_out) quads_out[4 * (size_t) qi + k] = (int) cellslot[q[k]];
            }
            int *f = F + 6 * (size_t) qi;
            if (dist_ref(v[0], v[2]) > dist_ref(v[1], v[3])) {
                f[0] = id[1]; f[1] = id[3]; f[2] = id[0];
                f[3] = id[3]; f[4] = id[1]; f[5] = id[2];
            } else {
                f[0] = id[2]; f[1] = id[0]; f[2] = id[1];
                f[3] = id[0]; f[4] = id[2]; f[5] = id[3];
            }
            qi++;
        }
    }
}

// ---- workspaces ----------------------------------------------------------------------------------
struct SpWs {
    u32 *counters;
    u32 *cinfo, *offA, *offB;        // per cell
    u64 *descA, *descB;
};
static size_t carve_sp_ws(Carver &c, size_t n, SpWs *out) {
    SpWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.cinfo = c.take<u32>(n + 1);
    b.offA = c.take<u32>(n + 1);
    b.offB = c.take<u32>(n + 1);
    b.descA = c.take<u64>(n / SPT_TILE + 2);
    b.descB = c.take<u64>(n / SPT_TILE + 2);
    if (out) *out = b;
    return c.bytes();
}
struct SpScratch {
    u32 *kx, *ky, *kz, *cand_rank;
    u64 *descV;
    GenSort gs;   // segmented sort by grid layer (segsort.cuh); its radix buffers serve the fallback
};
static size_t carve_sp_scratch(Carver &c, size_t nc, u32 x_planes, u32 Y, SpScratch *out) {
    SpScratch s;
    s.kx = c.take<u32>(nc);
    s.ky = c.take<u32>(nc);
    s.kz = c.take<u32>(nc);
    s.cand_rank = c.take<u32>(nc);
    s.descV = c.take<u64>(nc / UQ_TILE + 2);
    GenSort::carve(c, nc, x_planes, Y, &s.gs);
    if (out) *out = s;
    return c.bytes();
}
// sort (segmented by grid layer, radix sort as the fallback) + weld, then `faces()` enqueues the face kernel;
// leaves the counter block in h
constexpr u32 SP_SEG_MAX = 16u << 20;
static bool sp_use_seg(u32 nc) { return g_tuning[3] == 100 || (g_tuning[3] != 101 && nc < SP_SEG_MAX); }
template <typename FacesFn>
static int sort_weld_faces(const SpScratch &s, u32 nc, const Geom &g, u32 *counters, float x_lo_threshold, float x_hi_threshold, float *V,
                           cudaStream_t stream, u32 *h, bool use_seg, FacesFn faces) {
    const int sms = device_sms();
    for (int attempt = use_seg ? 0 : 1; attempt < 2; attempt++) {
        if (attempt == 0) {
            ISX_CUDA(gen_sort_run(s.kx, s.ky, s.kz, nc, g, s.gs, counters, stream));
            ISX_LAUNCH(k_unique, scan_blocks(sms), 256, 0, stream, nc, s.gs.seg.perm, s.gs.seg.skx, s.gs.seg.sky, s.gs.seg.skz, s.cand_rank, V,
                       counters, s.descV, host_float_key(x_lo_threshold), host_float_key(x_hi_threshold), nullptr, 0xffffffffu, true, gen_sort_gate());
        } else {   // NaN keys, or a piece the two bucket levels could not cut below the shared-memory capacity
            if (use_seg) {
                ISX_CUDA(gen_sort_reset_for_radix(counters, stream));
                ISX_CUDA(cudaMemsetAsync(s.descV, 0, ((size_t) nc / UQ_TILE + 2) * sizeof(u64), stream));
            }
            ISX_CUDA(radix_sort96(s.kx, s.ky, s.kz, nc, s.gs.seg.radix, stream));
            ISX_LAUNCH(k_unique, scan_blocks(sms), 256, 0, stream, nc, s.gs.seg.radix.perm[0], s.kx, s.ky, s.kz, s.cand_rank, V, counters,
                       s.descV, host_float_key(x_lo_threshold), host_float_key(x_hi_threshold));
        }
        faces();
        ISX_CUDA(cudaGetLastError());
        ISX_CUDA(cudaMemcpyAsync(h, counters, C_COUNT * sizeof(u32), cudaMemcpyDeviceToHost, stream));
        ISX_CUDA(cudaStreamSynchronize(stream));
        if (!h[C_ABORT] && !h[C_RADIX]) break;
    }
    return OK;
}
struct SpDcWs {
    u32 *counters;
    u32 *dinfo, *quad_off, *cand_off;
    unsigned char *used;
    u64 *descA, *descB;
    u32 *plane_start;   // X + 1
    u32 *qnb;           // 9 planes of n: list positions of the 3 other cells of the quad of (cell, axis), k_sp_dc_quads -> k_sp_dc_faces
};
static size_t carve_sp_dc_ws(Carver &c, size_t n, size_t X, SpDcWs *out) {
    SpDcWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.plane_start = c.take<u32>(X + 1);
    b.dinfo = c.take<u32>(n + 1);
    b.quad_off = c.take<u32>(n + 1);
    b.cand_off = c.take<u32>(n + 1);
    b.used = c.take<unsigned char>(n + 1);
    b.descA = c.take<u64>(n / SPT_TILE + 2);
    b.descB = c.take<u64>(n / SPT_TILE + 2);
    b.qnb = c.take<u32>(9 * n);
    if (out) *out = b;
    return c.bytes();
}

static int make_sparse_params(i64 X, i64 Y, i64 Z, const float *amin, const float *amax, float level, i64 n, SparseParams *out) {
    if (X < 2 || Y < 2 || Z < 2) return fail(E_INVALID, "sparse grid shape must be at least 2 points per axis");
    if (X > 0xffffffffLL || Y > 0xffffffffLL || Z > 0xffffffffLL) return fail(E_INVALID, "grid axis too long");
    if (n < 0 || n >= ((i64) 1 << 28)) return fail(E_INVALID, "too many sparse cells for one call (>= 2^28)");
    SparseParams p;
    p.g.X = X; p.g.Y = Y; p.g.Z = Z; p.g.x_off = 0; p.g.Xg = X;
    for (int a = 0; a < 3; a++) { p.g.amin[a] = amin[a]; p.g.asize[a] = amax[a] - amin[a]; }
    p.cy = Y - 1; p.cz = Z - 1;
    p.level = level;
    const i64 res[3] = {X, Y, Z};
    for (int a = 0; a < 3; a++) p.eps2[a] = plain_eps2(p.g.amin[a], p.g.asize[a], res[a]);
    *out = p;
    return OK;
}

int device_sms();
}   // namespace isx

using namespace isx;

static int grid_for(i64 n, int threads, int max_blocks) {
    i64 want = (n + threads - 1) / threads;
    return (int) (want < 1 ? 1 : (want > max_blocks ? max_blocks : want));
}

extern "C" {

// get_points / get_points_by_cell_indices: out = (n, 8, 3) f32
int isoext_sparse_points(int64_t X, int64_t Y, int64_t Z, const float *aabb_min, const float *aabb_max, const int64_t *cell_idx,
                         int64_t n, float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, 0.f, 0, &p);
    if (rc != OK) return rc;
    if (n <= 0) return OK;
    ISX_LAUNCH(k_sp_points, grid_for(n * 8, 256, device_sms() * 16), 256, 0, stream, p, cell_idx, n, out);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// filter_cell_indices: keep[i] = 1 iff the (n,8) values of cell i straddle the level
int isoext_sparse_crossing(const float *values8, int64_t n, float level, unsigned char *keep, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return OK;
    ISX_LAUNCH(k_sp_crossing, grid_for(n, 256, device_sms() * 16), 256, 0, stream, values8, n, level, keep);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

size_t isoext_mc_sparse_workspace_bytes(int64_t n) {
    Carver c(nullptr);
    return carve_sp_ws(c, (size_t) (n > 0 ? n : 1), nullptr);
}
size_t isoext_sparse_scratch_bytes(int64_t n_candidates, int64_t X, int64_t Y) {
    Carver c(nullptr);
    return carve_sp_scratch(c, (size_t) (n_candidates > 0 ? n_candidates : 1), (u32) X, (u32) Y, nullptr);
}

// marching_cubes on a SparseGrid, phase 1: counts_out[0..1] = triangles T, vertex candidates Vc.
// Only the cells [emit_begin, emit_end) of the list emit triangles (0 .. n on one GPU; the owned cells of a slab).
int isoext_mc_sparse_count(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_begin,
                           int64_t emit_end, void *workspace, size_t workspace_bytes, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (method != 0 && method != 1) return fail(E_METHOD, "Unknown method");
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, level, n, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = 0;
    if (n == 0) return OK;
    Carver c(workspace);
    SpWs b;
    if (carve_sp_ws(c, (size_t) n, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descA, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descB, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    stream_timer_mark(stream);
    if (emit_begin < 0 || emit_end > n || emit_begin > emit_end) return fail(E_INVALID, "emit range out of bounds");
    ISX_LAUNCH(k_sp_mc_classify, grid_for(n, 128, sms * 16), 128, 0, stream, values8, cell_idx, p, (u32) n, method, b.cinfo,
               (u32) emit_begin, (u32) emit_end);
    stream_timer_mark(stream);
    if (g_tuning[7] != 1) ISX_LAUNCH(k_sp_mc_dedupe, grid_for(n, 256, sms * 16), 256, 0, stream, values8, cell_idx, p, (u32) n, b.cinfo);
    ISX_LAUNCH(k_sp_scan2, scan_blocks(sms), 256, 0, stream, (u32) n, b.counters, b.cinfo, 8, 0xffu, 16, 0xfffu, b.offA, b.offB, b.descA,
               b.descB, (int) C_T, (int) C_VC, (int) C_TICKET_B);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_T];
    counts_out[1] = h[C_VC];
    return OK;
}

// phase 2: V (capacity Vc x 3), F (T x 3); counts_out[0..2] = welded vertices, # with x < x_lo_threshold, # with
// x < x_hi_threshold (slab ownership by position; -inf / +inf on one GPU)
int isoext_mc_sparse_emit(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                          const float *aabb_min, const float *aabb_max, float level, int method, float x_lo_threshold,
                          float x_hi_threshold, void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes,
                          int64_t n_candidates, float *V, int32_t *F, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, level, n, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = counts_out[2] = 0;
    if (n == 0 || n_candidates <= 0) return OK;
    if (n_candidates >= ((i64) 1 << 31)) return fail(E_INVALID, "too many vertex candidates");
    Carver c(workspace);
    SpWs b;
    if (carve_sp_ws(c, (size_t) n, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    Carver cs(scratch);
    SpScratch s;
    if (carve_sp_scratch(cs, (size_t) n_candidates, (u32) X, (u32) Y, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    const u32 nc = (u32) n_candidates;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(s.descV, 0, ((size_t) nc / UQ_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_sp_mc_keys, grid_for(n, 128, sms * 16), 128, 0, stream, values8, cell_idx, p, (u32) n, b.cinfo, b.offB, s.kx, s.ky, s.kz);
    u32 h[C_COUNT];
    // (cell, edge) vertices minus the duplicates along z (k_sp_mc_dedupe), segmented by grid layer up to SP_SEG_MAX
    // candidates, global radix sort beyond (measured: 1.10 vs 1.25 ms at 6.4 M candidates, 30.2 vs 21.1 ms at 100 M);
    // g_tuning[3] = 100 / 101 force one or the other
    rc = sort_weld_faces(s, nc, p.g, b.counters, x_lo_threshold, x_hi_threshold, V, stream, h, sp_use_seg(nc), [&]() {
        ISX_LAUNCH(k_sp_mc_faces, grid_for(n, 256, sms * 16), 256, 0, stream, (u32) n, method, b.cinfo, b.offA, b.offB, s.cand_rank, F);
    });
    if (rc != OK) return rc;
    counts_out[0] = h[C_V];
    counts_out[1] = h[C_NLO];
    counts_out[2] = h[C_NHI];
    return OK;
}

// get_intersection on a SparseGrid, phase 1.  cinfo / cellslot / its_off: caller-owned (n+1) u32 arrays.
//   counts_out[0..1] = active cells, intersections I
int isoext_its_sparse_count(const float *values8, int64_t n, float level, uint32_t *cinfo, uint32_t *cellslot, uint32_t *its_off,
                            void *workspace, size_t workspace_bytes, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    counts_out[0] = counts_out[1] = 0;
    if (n <= 0) return OK;
    if (n >= ((i64) 1 << 28)) return fail(E_INVALID, "too many sparse cells for one call (>= 2^28)");
    Carver c(workspace);
    SpWs b;
    if (carve_sp_ws(c, (size_t) n, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descA, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descB, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_sp_case, grid_for(n, 256, sms * 16), 256, 0, stream, values8, (u32) n, level, cinfo);
    ISX_LAUNCH(k_sp_scan2, scan_blocks(sms), 256, 0, stream, (u32) n, b.counters, cinfo, 8, 1u, 16, 0xfffu, cellslot, its_off, b.descA, b.descB,
               (int) C_T, (int) C_I, (int) C_TICKET_B);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_T];
    counts_out[1] = h[C_I];
    return OK;
}

// phase 2 (mode 0: points only, 1: points + normals, 2: normals only from existing points)
int isoext_its_sparse_emit(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, float level, int mode, const uint32_t *cinfo,
                           const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, int64_t n_its, float *points,
                           float *normals, uint32_t *cell_offsets, int64_t *cell_indices, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, level, n, &p);
    if (rc != OK) return rc;
    if (n <= 0) return OK;
    const int blocks = grid_for(n, 128, device_sms() * 16);
    if (mode == 0)
        ISX_LAUNCH((k_sp_its_emit<true, false>), blocks, 128, 0, stream, values8, cell_idx, p, (u32) n, cinfo, cellslot, its_off, points,
                   normals, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its);
    else if (mode == 1)
        ISX_LAUNCH((k_sp_its_emit<true, true>), blocks, 128, 0, stream, values8, cell_idx, p, (u32) n, cinfo, cellslot, its_off, points,
                   normals, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its);
    else
        ISX_LAUNCH((k_sp_its_emit<false, true>), blocks, 128, 0, stream, values8, cell_idx, p, (u32) n, cinfo, cellslot, its_off, points,
                   normals, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

size_t isoext_dc_sparse_workspace_bytes(int64_t n, int64_t X) {
    Carver c(nullptr);
    return carve_sp_dc_ws(c, (size_t) (n > 0 ? n : 1), (size_t) (X > 0 ? X : 1), nullptr);
}

// dual_contouring on a SparseGrid, phase 1: dual_v (n_cells x 3); counts_out[0..1] = quads Q, used vertices Vc
int isoext_dc_sparse_count(const float *values8, const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z,
                           const float *aabb_min, const float *aabb_max, const uint32_t *cinfo, const uint32_t *cellslot,
                           const uint32_t *its_off, const float *points, const float *normals, float reg, float svd_tol,
                           int64_t emit_begin, int64_t emit_end, float *dual_v, void *workspace, size_t workspace_bytes,
                           void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, 0.f, n, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = 0;
    if (n <= 0) return OK;
    if (emit_begin < 0 || emit_end > n || emit_begin > emit_end) return fail(E_INVALID, "emit range out of bounds");
    Carver c(workspace);
    SpDcWs b;
    if (carve_sp_dc_ws(c, (size_t) n, (size_t) X, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const int sms = device_sms();
    const int blocks = grid_for(n, 128, sms * 16);
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.used, 0, (size_t) n + 1, stream));
    ISX_CUDA(cudaMemsetAsync(b.descA, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descB, 0, ((size_t) n / SPT_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_sp_dc_solve, blocks, 128, 0, stream, cell_idx, p, (u32) n, cinfo, cellslot, its_off, points, normals, reg, svd_tol, dual_v);
    ISX_LAUNCH(k_sp_plane_start, grid_for(X, 256, sms * 4), 256, 0, stream, cell_idx, (u32) n, p, b.plane_start);
    ISX_LAUNCH(k_sp_dc_quads, blocks, 128, 0, stream, values8, cell_idx, p, (u32) n, cinfo, b.plane_start, b.dinfo, b.used, (u32) emit_begin,
               (u32) emit_end, b.qnb);
    ISX_LAUNCH(k_sp_fold_used, grid_for(n, 256, sms * 16), 256, 0, stream, (u32) n, b.used, b.dinfo);
    ISX_LAUNCH(k_sp_scan2, scan_blocks(sms), 256, 0, stream, (u32) n, b.counters, b.dinfo, 0, 7u, 8, 1u, b.quad_off, b.cand_off, b.descA, b.descB,
               (int) C_Q, (int) C_VC, (int) C_TICKET_D);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_Q];
    counts_out[1] = h[C_VC];
    return OK;
}

int isoext_dc_sparse_emit(const int64_t *cell_idx, int64_t n, int64_t X, int64_t Y, int64_t Z, const float *aabb_min,
                          const float *aabb_max, const uint32_t *cinfo, const uint32_t *cellslot, const float *dual_v, void *workspace,
                          size_t workspace_bytes, void *scratch, size_t scratch_bytes, int64_t n_candidates, float x_lo_threshold,
                          float x_hi_threshold, float *V, int32_t *F, int32_t *quads_out, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    SparseParams p;
    int rc = make_sparse_params(X, Y, Z, aabb_min, aabb_max, 0.f, n, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = counts_out[2] = 0;
    if (n <= 0 || n_candidates <= 0) return OK;
    Carver c(workspace);
    SpDcWs b;
    if (carve_sp_dc_ws(c, (size_t) n, (size_t) X, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    Carver cs(scratch);
    SpScratch s;
    if (carve_sp_scratch(cs, (size_t) n_candidates, (u32) X, (u32) Y, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    const u32 nc = (u32) n_candidates;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(s.descV, 0, ((size_t) nc / UQ_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_sp_dc_keys, grid_for(n, 256, sms * 16), 256, 0, stream, (u32) n, b.dinfo, b.cand_off, cellslot, dual_v, s.kx, s.ky, s.kz);
    u32 h[C_COUNT];
    // dual vertices: one candidate per cell, segmented by grid layer (0.39 ms vs 0.76 ms for the global radix sort at
    // 2.4 M; at 38.7 M the radix sort is the faster one: 29.1 vs 38.0 ms for the whole call)
    rc = sort_weld_faces(s, nc, p.g, b.counters, x_lo_threshold, x_hi_threshold, V, stream, h, sp_use_seg(nc), [&]() {
        ISX_LAUNCH(k_sp_dc_faces, grid_for(n, 128, sms * 16), 128, 0, stream, cell_idx, p, (u32) n, cinfo, b.qnb, b.dinfo, b.quad_off, b.cand_off,
                   cellslot, s.cand_rank, dual_v, F, quads_out);
    });
    if (rc != OK) return rc;
    counts_out[0] = h[C_V];
    counts_out[1] = h[C_NLO];
    counts_out[2] = h[C_NHI];
    return OK;
}

}   // extern "C"
