// mc_dense.cu -- dense-grid marching cubes for sm_100a.
//
// Replaces mc::marching_cubes on a UniformGrid (src/mc/mc.cu:17-68, src/mc/nagae.cu:34-102,
// src/mc/lorensen.cu:34-103, src/utils.cu:32-59 of the reference) with:
//   k_signbits   (dense.cuh)  one streaming read of the field, 4 B/voxel            [HBM-bound]
//   k_compact    (dense.cuh)  bit-parallel active-point compaction, look-back scan
//   k_cell_tris               per active cell: edge crossings, LUT triangles, degenerate drop,
//                             marks the edge-owned vertex slots that a kept triangle references
//   k_scan_entries            look-back scan: triangle offsets + vertex candidate ids
//   k_cand_pos                one position per used edge slot (edge-keyed welding, no hash pass)
//   radix_sort96 (radix.cuh)  reference vertex order = lexicographic (x,y,z) of positions
//   k_unique                  positional welding of candidates that snapped onto a shared corner
//   k_emit_faces              LUT-driven face emission with final vertex ids
// Every arithmetic step that decides a bit of the output uses explicit _rn intrinsics.
#include <cmath>
#include <cstdlib>
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "dense.cuh"
#include "radix.cuh"
#include "weld.cuh"
#include "segsort.cuh"

#include <cstring>

#define ISX_LUT_QUAL __device__
#include "mc_luts.inc"

namespace isx {

__device__ __forceinline__ u64 tri_word(int method, u32 cs) {
    return method == 0 ? kTriWords_nagae[cs] : kTriWords_lorensen[cs];
}

constexpr int SE_ITEMS = 8;   // consecutive entries per thread in k_scan_entries
constexpr int SE_TILE = 256 * SE_ITEMS;

struct McBuffers {
    // phase 1 (sized by the grid and the entry capacity)
    u32 *bits;
    u32 *row_start;
    u64 *descA;
    u32 *counters;
    uint2 *entries;
    u32 *nb;            // 3 per entry: lower bounds in rows (x,y+1), (x+1,y), (x+1,y+1)
    unsigned char *ntri, *trimask;
    unsigned char *used;   // 3 per entry
    u32 *tri_off, *cand_info;
    u64 *descT, *descU, *descV;
    unsigned char *span_cnt;   // entries per candidate 128-point span (written by the count pass, read by the fill pass)
    u32 *heavy_list;           // rows filled cooperatively (k_rowfill_heavy)
    size_t zero_bytes;       // bytes from `counters` that one memset clears at the start of a call
    float4 *vals4;  // per entry that owns an edge: the values at its point and at the +z / +y / +x neighbours (k_cell_tris -> k_cand_pos)
    u32 *bdelta;   // per entry: sort bucket (layer offset + sub-bucket) of its 3 owned edge vertices, one byte each
    SegHead seg;             // bucket histogram / offsets over the sort buckets (sort_buckets(p))
};

static size_t carve_mc(Carver &c, const DenseParams &p, size_t cap, McBuffers *out) {
    McBuffers b;
    // --- everything that must be zero at the start of a call is contiguous: ONE memset per call ---
    b.counters = c.take<u32>(C_COUNT);
    b.descA = c.take<u64>(compact_desc_count(p));
    b.descT = c.take<u64>(cap / SE_TILE + 2);
    b.descU = c.take<u64>(cap / SE_TILE + 2);
    b.descV = c.take<u64>(3 * cap / UQ_TILE + 2);
    SegHead::carve(c, (size_t) sort_buckets(p), &b.seg);
    b.row_start = c.take<u32>((size_t) p.R + 2);   // accumulates the per-row entry counts before it is scanned
    b.used = c.take<unsigned char>(3 * (cap + 2));
    b.zero_bytes = (size_t) ((char *) (b.used + 3 * (cap + 2)) - (char *) b.counters);
    // --- the rest is fully overwritten before it is read ---
    b.bits = c.take<u32>(signbits_words(p.P));
    b.entries = c.take<uint2>(cap + 1);
    b.nb = c.take<u32>(3 * cap);
    b.ntri = c.take<unsigned char>(cap);
    b.trimask = c.take<unsigned char>(cap);
    b.tri_off = c.take<u32>(cap);
    b.cand_info = c.take<u32>(cap + 1);
    b.bdelta = c.take<u32>(cap + 1);
    b.vals4 = c.take<float4>(cap + 1);
    b.heavy_list = c.take<u32>(compact_heavy_cap((u32) cap));
    b.span_cnt = c.take<unsigned char>(compact_span_bytes(p));
    if (out) *out = b;
    return c.bytes();
}

struct McScratch {
    u32 *kx, *ky, *kz;
    u32 *cand_rank;
    SegScratch seg;
};

static size_t carve_mc_scratch(Carver &c, size_t nc, McScratch *out) {
    McScratch s;
    s.kx = c.take<u32>(nc);
    s.ky = c.take<u32>(nc);
    s.kz = c.take<u32>(nc);
    s.cand_rank = c.take<u32>(nc);
    SegScratch::carve(c, nc, &s.seg);
    if (out) *out = s;
    return c.bytes();
}

// Sort bucket of the vertices on the entry's owned edges (segsort.cuh).  Any function of the position that is
// monotone in the lexicographic (x, y, z) order will do; this one is exact float compares only:
//   layer  b = #{i : px[i] <= x'} - 1; for a vertex owned by plane x it is x-1 (an in-plane vertex whose x rounded
//            just below px[x]), x, or x+1 (an x-edge vertex that landed on the next plane): code = delta+1 in {0,1,2}
//   inside a layer: vertices with x' == px[b] exactly come first, cut into gy groups by y (thresholds are the
//            y-plane positions py[g*ystep]); then the vertices with x' > px[b], cut into gx groups by
//            floor((x' - px[b]) * gx / (px[b+1] - px[b])) (monotone in x').
// Encoding: one byte per owned axis (0: +z edge, 1: +y, 2: +x) = code | sub << 2.
__device__ __forceinline__ u32 sub_bucket(const DenseParams &p, i64 xb /* global layer index */, float xv, float yv, u32 y) {
    if (xb < 0) return 0u;
    const u32 resx = (u32) p.g.Xg - 1;
    const float pb = axis_pos((u32) xb, resx, p.g.amin[0], p.g.asize[0]);
    if (xv == pb) {
        if (p.gy == 1) return 0u;
        const u32 g0 = y / p.ystep;
        const u32 resy = (u32) p.g.Y - 1;
        if (g0 >= 1 && yv < axis_pos(g0 * p.ystep, resy, p.g.amin[1], p.g.asize[1])) return g0 - 1;
        if (g0 + 1 < p.gy && yv >= axis_pos((g0 + 1) * p.ystep, resy, p.g.amin[1], p.g.asize[1])) return g0 + 1;
        return g0;
    }
    if (p.gx == 1) return p.gy;
    const float pb1 = axis_pos((u32) xb + 1, resx, p.g.amin[0], p.g.asize[0]);
    const float f = __fmul_rn(__fsub_rn(xv, pb), __fdiv_rn((float) p.gx, __fsub_rn(pb1, pb)));
    u32 k = f > 0.f ? (u32) f : 0u;   // NaN / negative -> 0; monotone
    if (k > p.gx - 1) k = p.gx - 1;
    return p.gy + k;
}

// v0 / vz / vy / vx: the values at the entry's point and at its +z / +y / +x neighbours (only those of owned edges are read)
__device__ __forceinline__ u32 owned_buckets(const DenseParams &p, u32 r, u32 own, float v0, float vz, float vy, float vx) {
    if (!own) return 0u;
    const u32 Y = (u32) p.g.Y;
    const u32 x = r / Y, y = r - x * Y;
    const u32 xg = x + (u32) p.g.x_off;
    const float px0 = axis_pos(xg, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
    const float py0 = axis_pos(y, Y - 1, p.g.amin[1], p.g.asize[1]);
    u32 out = 0;
    if (own & 1u) {
        const float t = edge_t(v0, vz, p.level);
        const float xv = lerp_ref(t, px0, px0);
        const u32 code = xv < px0 ? 0u : 1u;
        out |= code | (sub_bucket(p, (i64) xg + code - 1, xv, lerp_ref(t, py0, py0), y) << 2);
    }
    if (own & 2u) {
        const float t = edge_t(v0, vy, p.level);
        const float xv = lerp_ref(t, px0, px0);
        const float py1 = axis_pos(y + 1, Y - 1, p.g.amin[1], p.g.asize[1]);
        const u32 code = xv < px0 ? 0u : 1u;
        out |= (code | (sub_bucket(p, (i64) xg + code - 1, xv, lerp_ref(t, py0, py1), y) << 2)) << 8;
    }
    if (own & 4u) {
        const float t = edge_t(v0, vx, p.level);
        const float px1 = axis_pos(xg + 1, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
        const float xv = lerp_ref(t, px0, px1);
        const u32 code = xv >= px1 ? 2u : (xv < px0 ? 0u : 1u);
        out |= (code | (sub_bucket(p, (i64) xg + code - 1, xv, lerp_ref(t, py0, py0), y) << 2)) << 16;
    }
    return out;
}
// bucket id from the byte of one owned axis
__device__ __forceinline__ u32 bucket_of(u32 x, u32 byte, u32 nsub) { return (x + (byte & 3u)) * nsub + ((byte >> 2) & 63u); }

// ---------------------------------------------------------------------------------------------
// K3: one thread per entry that is a valid cell.
// ---------------------------------------------------------------------------------------------
// The kernel is bound by the latency of its dependent loads (entry -> row bounds -> binary searches -> neighbour
// entries, and the value gathers from HBM), so the value loads are issued first and overlap the searches.
template <bool IMPLICIT>
__global__ void __launch_bounds__(128) k_cell_tris(const float *__restrict__ values, DenseParams p, int method,
                                                   const uint2 *__restrict__ entries, const u32 *__restrict__ row_start,
                                                   u32 cap, const u32 *__restrict__ counters, u32 *__restrict__ nb,
                                                   unsigned char *__restrict__ ntri, unsigned char *__restrict__ trimask,
                                                   unsigned char *__restrict__ used, u32 *__restrict__ bdelta,
                                                   float4 *__restrict__ vals4) {
    pdl_wait();
    pdl_trigger();
    const u32 S = counters[C_S];
    if (S > cap) return;
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 w = e.y, r = e.x, z = ent_z(w), cs = ent_case(w), own = ent_own(w);
        const bool is_cell = ent_cell(w) && cs != 0u && cs != 255u;
        CellData c;
        if (is_cell) {
            load_cell_values<IMPLICIT>(values, p, r, z, c);
        } else if (own) {
            const u32 x = r / Y, y = r - x * Y;
            const i64 n = (i64) r * Z + z;
            c.v[0] = field_at<IMPLICIT>(values, p, n, 0, x, y, z);
            c.v[1] = (own & 1u) ? field_at<IMPLICIT>(values, p, n, 1, x, y, z + 1) : 0.f;
            c.v[2] = (own & 2u) ? field_at<IMPLICIT>(values, p, n, Z, x, y + 1, z) : 0.f;
            c.v[4] = (own & 4u) ? field_at<IMPLICIT>(values, p, n, p.YZ, x + 1, y, z) : 0.f;
        }
        // the position kernel reads the four values from here (coalesced) instead of gathering them from the field again
        if (own) vals4[s] = make_float4(c.v[0], c.v[1], c.v[2], c.v[4]);
        if (!is_cell) {
            bdelta[s] = owned_buckets(p, r, own, c.v[0], c.v[1], c.v[2], c.v[4]);
            ntri[s] = 0;
            trimask[s] = 0;
            continue;
        }
        const u32 rsY0 = row_start[r + 1], rsY1 = row_start[r + 2], rsX0 = row_start[r + Y], rsX1 = row_start[r + Y + 1],
                  rsXY1 = row_start[r + Y + 2];
        // the three neighbour rows are searched in lockstep: three independent loads per round instead of three
        // binary searches one after the other (the kernel is bound by the latency of exactly this chain)
        u32 lbY = rsY0, hiY = rsY1, lbX = rsX0, hiX = rsX1, lbXY = rsX1, hiXY = rsXY1;
        while (lbY < hiY || lbX < hiX || lbXY < hiXY) {
            const u32 mY = (lbY + hiY) >> 1, mX = (lbX + hiX) >> 1, mXY = (lbXY + hiXY) >> 1;
            const bool aY = lbY < hiY, aX = lbX < hiX, aXY = lbXY < hiXY;
            const u32 zY = aY ? ent_z(entries[mY].y) : 0u, zX = aX ? ent_z(entries[mX].y) : 0u, zXY = aXY ? ent_z(entries[mXY].y) : 0u;
            if (aY) { if (zY < z) lbY = mY + 1; else hiY = mY; }
            if (aX) { if (zX < z) lbX = mX + 1; else hiX = mX; }
            if (aXY) { if (zXY < z) lbXY = mXY + 1; else hiXY = mXY; }
        }
        nb[s] = lbY;                    // structure of arrays (three planes of `cap` entries): coalesced 4-byte accesses
        nb[cap + s] = lbX;
        nb[2 * (size_t) cap + s] = lbXY;
        bdelta[s] = owned_buckets(p, r, own, c.v[0], c.v[1], c.v[2], c.v[4]);

        const u32 status = edge_mask_of_case(cs);
        u32 slot[12];
        cell_edge_slots(entries, s, z, lbY, lbX, lbXY, S, slot);
        const u64 word = tri_word(method, cs);
        const u32 nt = (u32) (word >> 60);
        u32 mask = 0, cnt = 0;
        if (cell_is_plain(c, status, p)) {
            // no crossing is near a cell corner: every LUT triangle is kept and every sign-change edge is used
            // (the LUT rows use exactly the sign-change edges of their case: tools/gen_luts.py checks it)
            mask = (1u << nt) - 1u;
            cnt = nt;
#pragma unroll
            for (int k = 0; k < 12; k++)
                if ((status >> k) & 1u) used[slot[k]] = 1;
        } else {
        load_cell_positions(p, r, z, c);
        float ex[12], ey[12], ez[12];
#pragma unroll
        for (int k = 0; k < 12; k++)
            if ((status >> k) & 1u) cell_edge_point(c, k, p.level, ex[k], ey[k], ez[k]);
        for (u32 k = 0; k < nt; k++) {
            const u32 a = (u32) (word >> (12 * k)) & 15u, b = (u32) (word >> (12 * k + 4)) & 15u,
                      d = (u32) (word >> (12 * k + 8)) & 15u;
            // reference drops a triangle when two corners are bit-equal (src/mc/nagae.cu:71)
            const bool ab = ex[a] != ex[b] || ey[a] != ey[b] || ez[a] != ez[b];
            const bool ad = ex[a] != ex[d] || ey[a] != ey[d] || ez[a] != ez[d];
            const bool bd = ex[b] != ex[d] || ey[b] != ey[d] || ez[b] != ez[d];
            if (ab && ad && bd) {
                mask |= 1u << k;
                cnt++;
                used[slot[a]] = 1;
                used[slot[b]] = 1;
                used[slot[d]] = 1;
            }
        }
        }
        const u32 x = r / Y;
        const bool emit = x >= p.emit_lo && x < p.emit_hi;
        ntri[s] = emit ? (unsigned char) cnt : 0;
        trimask[s] = emit ? (unsigned char) mask : 0;
    }
}

// ---------------------------------------------------------------------------------------------
// K4: exclusive scans over the entries (triangle offsets, candidate ids), persistent + look-back.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scan_entries(u32 cap, u32 *__restrict__ counters, const unsigned char *__restrict__ ntri,
                                                      const unsigned char *__restrict__ used, u32 *__restrict__ tri_off,
                                                      u32 *__restrict__ cand_info, u64 *__restrict__ descT, u64 *__restrict__ descU,
                                                      const uint2 *__restrict__ entries, const u32 *__restrict__ bdelta,
                                                      u32 Y, u32 nsub, u32 *__restrict__ bucket_count, u32 nb, SegHead seg) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preT, s_preU;
    // x-bucket histogram window: the entries of a tile are sorted by x, so almost all of its vertices fall
    // into a handful of consecutive buckets; count those in shared memory, the rest directly in L2
    constexpr u32 WIN = 256;
    __shared__ u32 s_hist[WIN];
    __shared__ u32 s_last, s4[4];
    const u32 S = counters[C_S];
    if (S > cap) return;
    const u32 ntiles = (S + SE_TILE - 1) / SE_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_B], 1u);
        if (threadIdx.x < WIN) s_hist[threadIdx.x] = 0;
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 win0 = (entries[tile * SE_TILE].x / Y) * nsub;   // first bucket of layer x-1 of the tile's first entry
        const u32 s0 = tile * SE_TILE + threadIdx.x * SE_ITEMS;
        u32 nt[SE_ITEMS], um[SE_ITEMS], ex_[SE_ITEMS], bd_[SE_ITEMS], sumT = 0, sumU = 0;
        if (s0 + SE_ITEMS <= S) {
            // full group of 8 entries: ALL loads are issued before anything depends on them (the per-item version
            // serialised 16 dependent L2 round trips per tile: byte loads -> branch -> entry/bucket loads -> atomics)
            static_assert(SE_ITEMS == 8, "vector loads below assume 8 entries per thread");
            const uint2 n8 = *reinterpret_cast<const uint2 *>(ntri + s0);
            const uint2 ua = *reinterpret_cast<const uint2 *>(used + 3 * (size_t) s0);
            const uint2 ub = *reinterpret_cast<const uint2 *>(used + 3 * (size_t) s0 + 8);
            const uint2 uc = *reinterpret_cast<const uint2 *>(used + 3 * (size_t) s0 + 16);
            const uint4 *ep = reinterpret_cast<const uint4 *>(entries + s0);
            const uint4 e0 = ep[0], e1 = ep[1], e2 = ep[2], e3 = ep[3];
            const uint4 b0 = *reinterpret_cast<const uint4 *>(bdelta + s0), b1 = *reinterpret_cast<const uint4 *>(bdelta + s0 + 4);
            const u32 nw[2] = {n8.x, n8.y};
            const u32 uw[6] = {ua.x, ua.y, ub.x, ub.y, uc.x, uc.y};
            ex_[0] = e0.x; ex_[1] = e0.z; ex_[2] = e1.x; ex_[3] = e1.z; ex_[4] = e2.x; ex_[5] = e2.z; ex_[6] = e3.x; ex_[7] = e3.z;
            bd_[0] = b0.x; bd_[1] = b0.y; bd_[2] = b0.z; bd_[3] = b0.w; bd_[4] = b1.x; bd_[5] = b1.y; bd_[6] = b1.z; bd_[7] = b1.w;
#pragma unroll
            for (int j = 0; j < SE_ITEMS; j++) {
                nt[j] = (nw[j >> 2] >> (8 * (j & 3))) & 0xffu;
                u32 m = 0;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const int byte = 3 * j + a;
                    m |= ((uw[byte >> 2] >> (8 * (byte & 3))) & 0xffu) ? (1u << a) : 0u;
                }
                um[j] = m;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SE_ITEMS; j++) {
                const u32 s = s0 + j;
                nt[j] = 0; um[j] = 0; ex_[j] = 0; bd_[j] = 0;
                if (s < S) {
                    nt[j] = ntri[s];
                    um[j] = (used[3 * s] ? 1u : 0u) | (used[3 * s + 1] ? 2u : 0u) | (used[3 * s + 2] ? 4u : 0u);
                    ex_[j] = entries[s].x;
                    bd_[j] = bdelta[s];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < SE_ITEMS; j++) {
            sumT += nt[j];
            sumU += __popc(um[j]);
        }
        u32 totT, totU;
        u32 exT = block_exclusive_scan(sumT, &totT, sw);
        u32 exU = block_exclusive_scan(sumU, &totU, sw);
        const u32 warp = threadIdx.x >> 5;
        // warps 0 / 1 look back for the tile's prefixes; the histogram of the tile (shared-memory atomics) is built by the
        // other warps meanwhile and by these two afterwards, so the block does not idle at the barrier behind the look-back
        if (warp == 0) {
            u32 pre = lookback_exclusive(descT, 1, tile, totT, 1u);
            if (threadIdx.x == 0) s_preT = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descU, 1, tile, totU, 1u);
            if ((threadIdx.x & 31) == 0) s_preU = pre;
        }
#pragma unroll
        for (int j = 0; j < SE_ITEMS; j++) {
            if (um[j]) {   // x-bucket histogram of the vertices this entry owns (segsort.cuh)
                const u32 x = ex_[j] / Y, bd = bd_[j];
#pragma unroll
                for (int a = 0; a < 3; a++)
                    if ((um[j] >> a) & 1u) {
                        const u32 bkt = bucket_of(x, bd >> (8 * a), nsub);
                        if (bkt - win0 < WIN) atomicAdd(&s_hist[bkt - win0], 1u);
                        else atomicAdd(&bucket_count[bkt], 1u);
                    }
            }
        }
        __syncthreads();
        if (threadIdx.x < WIN && s_hist[threadIdx.x]) atomicAdd(&bucket_count[win0 + threadIdx.x], s_hist[threadIdx.x]);
        exT += s_preT;
        exU += s_preU;
        if (s0 + SE_ITEMS <= S) {
            u32 to[SE_ITEMS], ci[SE_ITEMS];
#pragma unroll
            for (int j = 0; j < SE_ITEMS; j++) {
                to[j] = exT;
                ci[j] = exU | (um[j] << 29);
                exT += nt[j];
                exU += __popc(um[j]);
            }
            *reinterpret_cast<uint4 *>(tri_off + s0) = make_uint4(to[0], to[1], to[2], to[3]);
            *reinterpret_cast<uint4 *>(tri_off + s0 + 4) = make_uint4(to[4], to[5], to[6], to[7]);
            *reinterpret_cast<uint4 *>(cand_info + s0) = make_uint4(ci[0], ci[1], ci[2], ci[3]);
            *reinterpret_cast<uint4 *>(cand_info + s0 + 4) = make_uint4(ci[4], ci[5], ci[6], ci[7]);
            if (s0 + SE_ITEMS == S) {
                counters[C_T] = exT;
                counters[C_VC] = exU;
            }
        } else {
#pragma unroll
            for (int j = 0; j < SE_ITEMS; j++) {
                const u32 s = s0 + j;
                if (s < S) {
                    tri_off[s] = exT;
                    cand_info[s] = exU | (um[j] << 29);
                    exT += nt[j];
                    exU += __popc(um[j]);
                    if (s == S - 1) {
                        counters[C_T] = exT;
                        counters[C_VC] = exU;
                    }
                }
            }
        }
        // the block that finishes the last tile turns the bucket histogram into offsets (saves a launch); one fence by the
        // thread that takes the ticket, behind the barrier, orders the whole block's writes before it
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            s_last = (atomicAdd(&counters[C_TICKET_E], 1u) == ntiles - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            seg_scan_block(nb, seg.count, seg.start, seg.cursor, seg.bigoff, counters + C_NBIG, counters + C_MAXB, sw, s4);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K5: positions (as sortable keys) of the used edge slots, computed once by the owning entry.
// ---------------------------------------------------------------------------------------------
// (CandOut / emit_candidate: segsort.cuh)
__global__ void __launch_bounds__(256) k_cand_pos(DenseParams p, const uint2 *__restrict__ entries, const u32 *__restrict__ counters,
                                                  const u32 *__restrict__ cand_info, const u32 *__restrict__ bdelta,
                                                  const float4 *__restrict__ vals4, CandOut out, u32 cand_cap, u32 entry_cap) {
    pdl_wait();
    pdl_trigger();
    const u32 S = counters[C_S];
    if (S > entry_cap || counters[C_VC] > cand_cap) return;   // single-call fast path: the host re-runs with larger buffers
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z, nsub = sort_nsub(p);
    const u32 lane = threadIdx.x & 31;
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < S; base += gridDim.x * blockDim.x) {   // warp-uniform
        const u32 s = base + lane;
        const u32 ci = s < S ? cand_info[s] : 0u;
        const u32 um = ci >> 29;
        if (!__any_sync(0xffffffffu, um != 0u)) continue;
        u32 id = ci & 0x1fffffffu;
        u32 x = 0, bd = 0;
        float px0 = 0.f, py0 = 0.f, pz0 = 0.f;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // values at the point and at its +z / +y / +x neighbours (k_cell_tris)
        u32 y = 0, z = 0, xg = 0;
        if (um) {
            const uint2 e = entries[s];
            v = vals4[s];
            bd = bdelta[s];
            const u32 r = e.x;
            z = ent_z(e.y);
            x = r / Y;
            y = r - x * Y;
            xg = x + (u32) p.g.x_off;
            px0 = axis_pos(xg, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
            py0 = axis_pos(y, Y - 1, p.g.amin[1], p.g.asize[1]);
            pz0 = axis_pos(z, Z - 1, p.g.amin[2], p.g.asize[2]);
        }
        {   // +z edge
            const bool has = um & 1u;
            u32 a = 0, b2 = 0, c = 0;
            if (has) {
                const float t = edge_t(v.x, v.y, p.level);
                const float pz1 = axis_pos(z + 1, Z - 1, p.g.amin[2], p.g.asize[2]);
                a = float_key(lerp_ref(t, px0, px0));
                b2 = float_key(lerp_ref(t, py0, py0));
                c = float_key(lerp_ref(t, pz0, pz1));
            }
            emit_candidate(has, bucket_of(x, bd, nsub), id, a, b2, c, out);
            if (has) id++;
        }
        {   // +y edge
            const bool has = um & 2u;
            u32 a = 0, b2 = 0, c = 0;
            if (has) {
                const float t = edge_t(v.x, v.z, p.level);
                const float py1 = axis_pos(y + 1, Y - 1, p.g.amin[1], p.g.asize[1]);
                a = float_key(lerp_ref(t, px0, px0));
                b2 = float_key(lerp_ref(t, py0, py1));
                c = float_key(lerp_ref(t, pz0, pz0));
            }
            emit_candidate(has, bucket_of(x, bd >> 8, nsub), id, a, b2, c, out);
            if (has) id++;
        }
        {   // +x edge
            const bool has = um & 4u;
            u32 a = 0, b2 = 0, c = 0;
            if (has) {
                const float t = edge_t(v.x, v.w, p.level);
                const float px1 = axis_pos(xg + 1, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
                a = float_key(lerp_ref(t, px0, px1));
                b2 = float_key(lerp_ref(t, py0, py0));
                c = float_key(lerp_ref(t, pz0, pz0));
            }
            emit_candidate(has, bucket_of(x, bd >> 16, nsub), id, a, b2, c, out);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K7: faces.  One thread per emitting cell; triangle k of the LUT row goes to tri_off[s] + (#kept before k).
// (Assembling the triangles of 128 entries in shared memory for coalesced stores was measured slower: 56 vs 48 us at 1024^3.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_emit_faces(DenseParams p, int method, const uint2 *__restrict__ entries,
                                                    const u32 *__restrict__ counters, const u32 *__restrict__ nb,
                                                    const unsigned char *__restrict__ trimask, const u32 *__restrict__ tri_off,
                                                    const u32 *__restrict__ cand_info, const u32 *__restrict__ cand_rank,
                                                    int *__restrict__ F, u32 entry_cap, Gate gate) {
    pdl_wait();
    pdl_trigger();
    const u32 S = counters[C_S];
    if (S > entry_cap || gate_bad(counters, gate)) return;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const u32 mask = trimask[s];
        if (!mask) continue;
        const uint2 e = entries[s];
        const u32 z = ent_z(e.y), cs = ent_case(e.y);
        u32 slot[12];
        cell_edge_slots(entries, s, z, nb[s], nb[entry_cap + s], nb[2 * (size_t) entry_cap + s], S, slot);
        const u64 word = tri_word(method, cs);
        const u32 nt = (u32) (word >> 60);
        size_t o = 3 * (size_t) tri_off[s];
        for (u32 k = 0; k < nt; k++) {
            if (!((mask >> k) & 1u)) continue;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const u32 edge = (u32) (word >> (12 * k + 4 * j)) & 15u;
                const u32 sl = slot[edge];
                const u32 ent = sl / 3, axis = sl - 3 * ent;
                const u32 ci = cand_info[ent];
                const u32 cand = (ci & 0x1fffffffu) + __popc((ci >> 29) & ((1u << axis) - 1u));
                F[o + j] = (int) cand_rank[cand];
            }
            o += 3;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_dense_params(i64 X, i64 Y, i64 Z, i64 x_off, i64 Xg, const float *amin, const float *amax, float level,
                       i64 emit_lo, i64 emit_hi, DenseParams *out) {
    if (X < 1 || Y < 1 || Z < 1) return fail(E_INVALID, "grid shape must be positive");
    if (Z > 65535) return fail(E_INVALID, "Z (points along the last axis) must be <= 65535");
    if (X * Y >= (i64) 1 << 31) return fail(E_INVALID, "X*Y must be < 2^31");
    DenseParams p;
    p.g.X = X; p.g.Y = Y; p.g.Z = Z;
    p.g.x_off = x_off;
    p.g.Xg = Xg;
    for (int a = 0; a < 3; a++) {
        p.g.amin[a] = amin[a];
        p.g.asize[a] = amax[a] - amin[a];   // float subtraction, as the reference does on the host
    }
    p.P = X * Y * Z;
    p.YZ = Y * Z;
    p.CPR = (u32) ((Z + 31) / 32);
    i64 nq = X * Y * (i64) p.CPR;
    if (nq >= (i64) 1 << 31) return fail(E_INVALID, "grid too large for one slab (X*Y*ceil(Z/32) must be < 2^31)");
    p.NQ = (u32) nq;
    p.R = (u32) (X * Y);
    p.level = level;
    p.sdf = nullptr;
    p.emit_lo = (u32) (emit_lo < 0 ? 0 : emit_lo);
    p.emit_hi = (u32) (emit_hi < 0 ? 0 : emit_hi);
    // sort buckets: a layer holds ~ (2..8) * max(Y,Z) vertices and a bucket must stay below SEG_CAP = 4096 for
    // the shared-memory sort; finer buckets do not sort faster (measured: profiles/r1_sort_group_sweep.txt)
    const i64 m = Y > Z ? Y : Z;
    u32 gy = 1;
    while (gy < 16 && (i64) gy * 2 * 1024 <= m) gy *= 2;
    p.gy = gy;
    p.gx = gy > 1 ? gy / 2 : 1;
    {   // tuning hook (tools/tune_sort_groups.py): ISX_SORT_GY / ISX_SORT_GX override the heuristic
        static int egy = -1, egx = -1;
        if (egy < 0) {
            const char *a = getenv("ISX_SORT_GY"), *b = getenv("ISX_SORT_GX");
            egy = a ? atoi(a) : 0;
            egx = b ? atoi(b) : 0;
        }
        if (egy > 0 && egy <= 32) p.gy = (u32) egy;
        if (egx > 0 && egx <= 32) p.gx = (u32) egx;
    }
    p.ystep = (u32) ((Y + p.gy - 1) / p.gy);
    {   // k_cell_tris fast path (dense.cuh: cell_is_plain): eps * cell = 16 ulp of the largest coordinate of the axis
        const i64 res[3] = {Xg, Y, Z};
        for (int a = 0; a < 3; a++) p.eps2[a] = plain_eps2(p.g.amin[a], p.g.asize[a], res[a]);
    }
    *out = p;
    return OK;
}

int device_sms() {
    static thread_local int n_of[64];   // per device (and per host thread: no lock needed)
    int dev = 0;
    cudaGetDevice(&dev);
    int &n = n_of[dev & 63];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}   // namespace isx

using namespace isx;

extern "C" {

size_t isoext_mc_dense_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t cap_entries) {
    DenseParams p;
    float z3[3] = {0, 0, 0}, o3[3] = {1, 1, 1};
    if (make_dense_params(X, Y, Z, 0, X, z3, o3, 0.f, 0, X - 1, &p) != OK) return 0;
    Carver c(nullptr);
    return carve_mc(c, p, (size_t) cap_entries, nullptr);
}

size_t isoext_mc_dense_scratch_bytes(int64_t n_candidates) {
    Carver c(nullptr);
    return carve_mc_scratch(c, (size_t) (n_candidates > 0 ? n_candidates : 1), nullptr);
}

// ---- shared enqueue helpers ---------------------------------------------------------------------
// phase 1 = zero-fill of the per-call state, the volume stream, then the surface-sized analysis
static int enqueue_zero(const McBuffers &b, cudaStream_t stream) {
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, b.zero_bytes, stream));
    return OK;
}
// Halo split (slab-sharded grids): the first halo_lo and the last halo_hi x planes of the slab are still being pulled
// from the neighbours (on another stream; `halo_event` is recorded behind the pull).  The volume stream over the
// owned planes starts at once and only the few halo planes wait for the event, so the pull costs no time.
struct HaloSplit { i64 lo = 0, hi = 0; cudaEvent_t event = nullptr; };
static int enqueue_signbits(const float *values, const DenseParams &p, const McBuffers &b, cudaStream_t stream, const HaloSplit &h = HaloSplit()) {
    const i64 plane = p.YZ;
    if (p.sdf) {                       // implicit field: compute pass with Lipschitz culling, nothing to wait for
        if (h.event) ISX_CUDA(cudaStreamWaitEvent(stream, h.event, 0));
        launch_sdf_bits(p, b.bits, stream);
        return OK;
    }
    unsigned char *sum = span_sum_of(b.bits, p.P);
    if (!h.event) {
        launch_signbits(values, b.bits, sum, p.P, p.level, stream);
        return OK;
    }
    if ((plane & 255) != 0 || h.lo < 0 || h.hi < 0 || h.lo + h.hi >= p.g.X) {   // pieces must be multiples of 256 points
        ISX_CUDA(cudaStreamWaitEvent(stream, h.event, 0));
        launch_signbits(values, b.bits, sum, p.P, p.level, stream);
        return OK;
    }
    const i64 lo = h.lo * plane, hi = h.hi * plane, mid = p.P - lo - hi;
    launch_signbits(values + lo, b.bits + (lo >> 5), sum + (lo >> 7), mid, p.level, stream, hi == 0);
    ISX_CUDA(cudaStreamWaitEvent(stream, h.event, 0));
    if (lo) launch_signbits(values, b.bits, sum, lo, p.level, stream, false);
    if (hi) launch_signbits(values + lo + mid, b.bits + ((lo + mid) >> 5), sum + ((lo + mid) >> 7), hi, p.level, stream, true);
    return OK;
}
static void enqueue_compact(const DenseParams &p, const McBuffers &b, u32 cap, cudaStream_t stream) {
    launch_compact(b.bits, p, b.entries, cap, b.row_start, b.descA, b.counters, b.span_cnt, b.heavy_list, stream);
}
static int enqueue_analysis(const float *values, const DenseParams &p, int method, const McBuffers &b, u32 cap, cudaStream_t stream) {
    const int sms = device_sms();
    const u32 nb = sort_buckets(p);
    const int ct_blocks = sms * (g_tuning[4] > 0 ? g_tuning[4] : 32);
    if (p.sdf)
        ISX_LAUNCH_PDL(k_cell_tris<true>, ct_blocks, 128, 0, stream, values, p, method, b.entries, b.row_start, cap, b.counters, b.nb, b.ntri,
                   b.trimask, b.used, b.bdelta, b.vals4);
    else
        ISX_LAUNCH_PDL(k_cell_tris<false>, ct_blocks, 128, 0, stream, values, p, method, b.entries, b.row_start, cap, b.counters, b.nb, b.ntri,
                   b.trimask, b.used, b.bdelta, b.vals4);
    ISX_LAUNCH_PDL(k_scan_entries, scan_blocks(sms), 256, 0, stream, cap, b.counters, b.ntri, b.used, b.tri_off, b.cand_info, b.descT, b.descU,
               b.entries, b.bdelta, (u32) p.g.Y, sort_nsub(p), b.seg.count, nb, b.seg);
    ISX_CUDA(cudaGetLastError());
    return OK;
}
static int enqueue_phase1(const float *values, const DenseParams &p, int method, const McBuffers &b, u32 cap, cudaStream_t stream,
                          const HaloSplit &halo = HaloSplit()) {
    int rc = enqueue_zero(b, stream);
    if (rc != OK) return rc;
    rc = enqueue_signbits(values, p, b, stream, halo);
    if (rc != OK) return rc;
    enqueue_compact(p, b, cap, stream);
    return enqueue_analysis(values, p, method, b, cap, stream);
}

// host_nc: number of candidates if the host knows it (two-phase path), else 0 with device_counts = true:
// the kernels then read the counts from the counter block and do nothing if a capacity is exceeded.
static int enqueue_phase2(const float *values, const DenseParams &p, int method, const McBuffers &b, const McScratch &s,
                          u32 entry_cap, u32 host_nc, u32 n_big, bool device_counts, u32 cand_cap, u32 tri_cap, u32 big_cap,
                          bool allow_radix, float x_lo_threshold,
                          float x_hi_threshold, float *V, int32_t *F, cudaStream_t stream) {
    const int sms = device_sms();
    const u32 *n_dev = device_counts ? b.counters + C_VC : nullptr;
    const u32 grid_n = device_counts ? cand_cap : host_nc;
    const CandOut co{s.kx, s.ky, s.kz, s.seg.perm0, s.seg.cbucket, b.seg.count, b.seg.start, b.seg.bigoff, b.seg.cursor,
                     s.seg.bkx, s.seg.bky, s.seg.bkz, s.seg.bid, b.seg.xinvmin, b.seg.xmax};
    ISX_LAUNCH_PDL(k_cand_pos, sms * 8, 256, 0, stream, p, b.entries, b.counters, b.cand_info, b.bdelta, b.vals4, co, cand_cap, entry_cap);
    ISX_CUDA(seg_sort_run(s.kx, s.ky, s.kz, host_nc, n_dev, cand_cap, grid_n, sort_buckets(p), n_big,
                          device_counts ? b.counters + C_NBIG : nullptr, big_cap, b.seg, s.seg,
                          SegGeom{p.g.amin[0], p.g.asize[0], p.g.amin[1], p.g.asize[1], p.g.amin[2], p.g.asize[2], (u32) p.g.Xg, (u32) p.g.Y,
                                  (u32) p.g.Z, p.g.x_off, p.gy, p.gx, p.ystep,
                                  p.ystep + 2 <= (u32) SEG_GROUPS && !getenv("ISX_SORT_NO_GROUPS")},
                          allow_radix, b.counters + C_RADIX, stream));
    // single-sync path: the weld and face kernels check the capacities / sort completeness themselves (dense.cuh: Gate)
    Gate gate;
    if (device_counts) {
        gate.entry_cap = entry_cap; gate.cand_cap = cand_cap; gate.tri_cap = tri_cap; gate.big_cap = big_cap;
        gate.allow_radix = allow_radix ? 1 : 0;
        gate.on = 1;
    }
    const u32 klo = host_float_key(x_lo_threshold), khi = host_float_key(x_hi_threshold);
    ISX_LAUNCH(k_unique, scan_blocks(sms), 256, 0, stream, host_nc, s.seg.perm, s.seg.skx, s.seg.sky, s.seg.skz, s.cand_rank, V, b.counters,
               b.descV, klo, khi, n_dev, cand_cap, true, gate);
    ISX_LAUNCH_PDL(k_emit_faces, sms * (g_tuning[5] > 0 ? g_tuning[5] : 16), 128, 0, stream, p, method, b.entries, b.counters, b.nb, b.trimask, b.tri_off, b.cand_info,
               s.cand_rank, F, entry_cap, gate);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

static int read_counters(const McBuffers &b, u32 *h, cudaStream_t stream) {
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, C_COUNT * sizeof(u32), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    return OK;
}

// Phase 1: classify + compact + per-cell triangle analysis + scans.
// counts_out[0..3] = entries S, triangles T, vertex candidates Vc (upper bound of the vertex count), n_big.
int isoext_mc_dense_count(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_x_lo,
                          int64_t emit_x_hi, void *workspace, size_t workspace_bytes, int64_t cap_entries,
                          const void *sdf_program, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (method != 0 && method != 1) return fail(E_METHOD, "Unknown method");
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, emit_x_lo, emit_x_hi, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (!p.sdf && (reinterpret_cast<uintptr_t>(values) & 31u) != 0) return fail(E_INVALID, "values must be 32-byte aligned");
    if (cap_entries < 1 || cap_entries >= ((i64) 1 << 29)) return fail(E_INVALID, "cap_entries out of range");
    Carver c(workspace);
    McBuffers b;
    if (carve_mc(c, p, (size_t) cap_entries, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const u32 cap = (u32) cap_entries;
    rc = enqueue_phase1(values, p, method, b, cap, stream);
    if (rc != OK) return rc;
    u32 h[C_COUNT];
    rc = read_counters(b, h, stream);
    if (rc != OK) return rc;
    counts_out[0] = h[C_S];
    counts_out[1] = h[C_T];
    counts_out[2] = h[C_VC];
    counts_out[3] = h[C_NBIG];   // candidates living in x-buckets too large for the shared-memory sort
    if (h[C_S] > cap) return fail(E_CAPACITY, "entry capacity exceeded; retry with cap_entries >= counts_out[0]");
    return OK;
}

// Phase 2: vertex positions, reference ordering, welding, faces.
//   V: capacity counts[2] x 3 floats; F: counts[1] x 3 int32 (local vertex ids, see counts_out).
//   x_lo_threshold / x_hi_threshold: welded vertices are classified by their x position into
//   [-inf, lo) | [lo, hi) | [hi, +inf) -- the slab ownership rule (pass -inf / +inf on one GPU).
//   counts_out[0..2] = welded vertices V, #with x < lo, #with x < hi.
int isoext_mc_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                         const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_x_lo,
                         int64_t emit_x_hi, void *workspace, size_t workspace_bytes, int64_t cap_entries, void *scratch,
                         size_t scratch_bytes, int64_t n_candidates, int64_t n_big, float x_lo_threshold, float x_hi_threshold,
                         float *V, int32_t *F, const void *sdf_program, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (method != 0 && method != 1) return fail(E_METHOD, "Unknown method");
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, emit_x_lo, emit_x_hi, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    Carver c(workspace);
    McBuffers b;
    if (carve_mc(c, p, (size_t) cap_entries, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    counts_out[0] = counts_out[1] = counts_out[2] = 0;
    if (n_candidates <= 0) return OK;
    if (n_candidates >= ((i64) 1 << 29)) return fail(E_INVALID, "too many vertex candidates for one slab (>= 2^29)");
    Carver cs(scratch);
    McScratch s;
    if (carve_mc_scratch(cs, (size_t) n_candidates, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    rc = enqueue_phase2(values, p, method, b, s, (u32) cap_entries, (u32) n_candidates, (u32) n_big, false, 0xffffffffu, 0xffffffffu, 0, true,
                        x_lo_threshold,
                        x_hi_threshold, V, F, stream);
    if (rc != OK) return rc;
    u32 h[C_COUNT];
    rc = read_counters(b, h, stream);
    if (rc != OK) return rc;
    counts_out[0] = h[C_V];
    counts_out[1] = h[C_NLO];
    counts_out[2] = h[C_NHI];
    return OK;
}

// Single-call fast path: both phases are enqueued back to back and the stream is synchronised ONCE.
// The caller provides output / scratch capacities (typically the sizes of the previous extraction of the
// same grid): V has room for cand_cap rows, F for tri_cap rows, scratch for cand_cap candidates.
// Returns ISOEXT_OK with counts_out = {S, T, Vc, n_big, V, n_lo, n_hi} when everything fitted; returns
// 1 ("not completed") with counts_out[0..3] filled when a capacity was exceeded (entries, candidates,
// triangles, or big_cap = the number of candidates in oversized x-buckets the radix fallback was sized for;
// 0 = do not enqueue the fallback) -- outputs are then undefined and the caller uses count + emit.
int isoext_mc_dense_run(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                        const float *aabb_min, const float *aabb_max, float level, int method, int64_t emit_x_lo,
                        int64_t emit_x_hi, void *workspace, size_t workspace_bytes, int64_t cap_entries, void *scratch,
                        size_t scratch_bytes, int64_t cand_cap, int64_t tri_cap, int64_t big_cap, int radix, float x_lo_threshold,
                        float x_hi_threshold, int64_t halo_planes_lo, int64_t halo_planes_hi, void *halo_event, float *V, int32_t *F,
                        const void *sdf_program, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (method != 0 && method != 1) return fail(E_METHOD, "Unknown method");
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, emit_x_lo, emit_x_hi, &p);
    if (rc != OK) return rc;
    p.sdf = static_cast<const SdfProg *>(sdf_program);
    if (!p.sdf && (reinterpret_cast<uintptr_t>(values) & 31u) != 0) return fail(E_INVALID, "values must be 32-byte aligned");
    if (cap_entries < 1 || cap_entries >= ((i64) 1 << 29)) return fail(E_INVALID, "cap_entries out of range");
    if (cand_cap < 1 || cand_cap >= ((i64) 1 << 29) || tri_cap < 1) return fail(E_INVALID, "capacities out of range");
    Carver c(workspace);
    McBuffers b;
    if (carve_mc(c, p, (size_t) cap_entries, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    Carver cs(scratch);
    McScratch s;
    if (carve_mc_scratch(cs, (size_t) cand_cap, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    const u32 cap = (u32) cap_entries;
    HaloSplit halo;
    halo.lo = halo_planes_lo; halo.hi = halo_planes_hi; halo.event = static_cast<cudaEvent_t>(halo_event);
    rc = enqueue_phase1(values, p, method, b, cap, stream, halo);
    if (rc != OK) return rc;
    if (big_cap < 0 || big_cap > cand_cap) return fail(E_INVALID, "big_cap out of range");
    rc = enqueue_phase2(values, p, method, b, s, cap, 0, 0, true, (u32) cand_cap, (u32) (tri_cap > 0xffffffffLL ? 0xffffffffLL : tri_cap),
                        (u32) big_cap, radix != 0, x_lo_threshold, x_hi_threshold, V, F, stream);
    if (rc != OK) return rc;
    u32 h[C_COUNT];
    rc = read_counters(b, h, stream);
    if (rc != OK) return rc;
    counts_out[0] = h[C_S]; counts_out[1] = h[C_T]; counts_out[2] = h[C_VC]; counts_out[3] = h[C_NBIG];
    counts_out[4] = h[C_V]; counts_out[5] = h[C_NLO]; counts_out[6] = h[C_NHI]; counts_out[7] = h[C_RADIX];
    if (h[C_S] > cap || h[C_VC] > (u32) cand_cap || (i64) h[C_T] > tri_cap || h[C_NBIG] > (u32) big_cap) return 1;
    if (h[C_RADIX] > 0 && !radix) return 1;   // the radix last resort was needed but not enqueued
    return OK;
}

int isoext_grid_points_dense(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                             const float *aabb_max, float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    Geom g;
    g.X = X; g.Y = Y; g.Z = Z; g.x_off = x_offset; g.Xg = X_global;
    for (int a = 0; a < 3; a++) { g.amin[a] = aabb_min[a]; g.asize[a] = aabb_max[a] - aabb_min[a]; }
    i64 P = X * Y * Z;
    if (P <= 0) return OK;
    i64 want = (P + 255) / 256;
    int blocks = (int) (want > (i64) device_sms() * 16 ? (i64) device_sms() * 16 : want);
    ISX_LAUNCH(k_grid_points, blocks, 256, 0, stream, g, out);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// Map local (extended-slab) vertex ids in F to global ids after the count all-gather:
//   id <  n_lo         -> base_mine - (n_lo - id)      (owned by the previous rank: tail of its range)
//   n_lo <= id < n_hi  -> base_mine + (id - n_lo)
//   id >= n_hi         -> base_next + (id - n_hi)      (owned by the next rank: head of its range)
__global__ void k_relabel_faces(int *F, i64 n, i64 n_lo, i64 n_hi, i64 base_mine, i64 base_next) {
    for (i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64) gridDim.x * blockDim.x) {
        i64 id = F[i];
        i64 g = id < n_lo ? base_mine - (n_lo - id) : (id < n_hi ? base_mine + (id - n_lo) : base_next + (id - n_hi));
        F[i] = (int) g;
    }
}
int isoext_relabel_faces(int32_t *F, int64_t n_ids, int64_t n_lo, int64_t n_hi, int64_t base_mine, int64_t base_next,
                         void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_ids <= 0) return OK;
    i64 want = (n_ids + 255) / 256;
    int blocks = (int) (want > 148 * 16 ? 148 * 16 : want);
    ISX_LAUNCH(k_relabel_faces, blocks, 256, 0, stream, F, n_ids, n_lo, n_hi, base_mine, base_next);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

}   // extern "C"
