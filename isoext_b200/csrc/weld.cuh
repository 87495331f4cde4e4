// weld.cuh -- positional welding of sorted vertex candidates (shared by MC and DC).
//
// Replaces thrust::unique + lower_bound of the reference's vertex_welding (src/utils.cu:49-55):
// after radix_sort96 the candidates are in lexicographic (x,y,z) order; equal neighbours collapse
// into one output vertex and every candidate learns its final rank.
#pragma once
#include "dense.cuh"

#include <cstring>

namespace isx {

constexpr int UQ_ITEMS = 8;   // consecutive sorted candidates per thread
constexpr int UQ_TILE = 256 * UQ_ITEMS;

// unique over the sorted candidates (look-back scan of "differs from predecessor"); also counts how
// many welded vertices lie below the slab thresholds (multi-GPU ownership).
static __global__ void __launch_bounds__(256) k_unique(u32 n, const u32 *__restrict__ perm, const u32 *__restrict__ kx,
                                                const u32 *__restrict__ ky, const u32 *__restrict__ kz,
                                                u32 *__restrict__ cand_rank, float *__restrict__ V, u32 *__restrict__ counters,
                                                u64 *__restrict__ desc, u32 key_lo, u32 key_hi,
                                                const u32 *__restrict__ n_dev = nullptr, u32 n_cap = 0xffffffffu,
                                                bool keys_sorted = false, Gate gate = Gate()) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_pre;
    // the tile's new vertices have consecutive ranks: they are staged here and copied out with coalesced stores
    // (three scalar stores per vertex at a 12-byte stride cost 22 of the kernel's 47 us at 1024^3)
    __shared__ float sv[3 * UQ_TILE];
    if (n_dev) n = *n_dev;       // single-call fast path: the count lives on the device
    if (n > n_cap || gate_bad(counters, gate)) return;   // `perm` is incomplete, the host re-runs
    const u32 ntiles = (n + UQ_TILE - 1) / UQ_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_C], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 i0 = tile * UQ_TILE + threadIdx.x * UQ_ITEMS;
        u32 c[UQ_ITEMS], x[UQ_ITEMS], y[UQ_ITEMS], z[UQ_ITEMS], isnew = 0, cnt = 0, nlo = 0, nhi = 0;
        // predecessor of the thread's first item
        u32 px = 0, py = 0, pz = 0;
        const bool full = i0 + UQ_ITEMS <= n;
        // keys_sorted: kx/ky/kz are already in sorted order (indexed by position, not by candidate id)
        if (full && keys_sorted) {
            // all loads up front, 128 bits each (the per-item version serialised its L2 round trips)
            static_assert(UQ_ITEMS == 8, "vector loads below assume 8 candidates per thread");
            const uint4 c0 = *reinterpret_cast<const uint4 *>(perm + i0), c1 = *reinterpret_cast<const uint4 *>(perm + i0 + 4);
            const uint4 x0 = *reinterpret_cast<const uint4 *>(kx + i0), x1 = *reinterpret_cast<const uint4 *>(kx + i0 + 4);
            const uint4 y0 = *reinterpret_cast<const uint4 *>(ky + i0), y1 = *reinterpret_cast<const uint4 *>(ky + i0 + 4);
            const uint4 z0 = *reinterpret_cast<const uint4 *>(kz + i0), z1 = *reinterpret_cast<const uint4 *>(kz + i0 + 4);
            if (i0 > 0) { px = kx[i0 - 1]; py = ky[i0 - 1]; pz = kz[i0 - 1]; }
            c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
            x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
            y[0] = y0.x; y[1] = y0.y; y[2] = y0.z; y[3] = y0.w; y[4] = y1.x; y[5] = y1.y; y[6] = y1.z; y[7] = y1.w;
            z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
        } else if (full) {
            const uint4 c0 = *reinterpret_cast<const uint4 *>(perm + i0), c1 = *reinterpret_cast<const uint4 *>(perm + i0 + 4);
            c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
            const u32 pc = i0 > 0 ? perm[i0 - 1] : 0u;
#pragma unroll
            for (int j = 0; j < UQ_ITEMS; j++) { x[j] = kx[c[j]]; y[j] = ky[c[j]]; z[j] = kz[c[j]]; }
            if (i0 > 0) { px = kx[pc]; py = ky[pc]; pz = kz[pc]; }
        } else {
            if (i0 > 0 && i0 < n) {
                const u32 pc = keys_sorted ? i0 - 1 : perm[i0 - 1];
                px = kx[pc]; py = ky[pc]; pz = kz[pc];
            }
#pragma unroll
            for (int j = 0; j < UQ_ITEMS; j++) {
                const u32 i = i0 + j;
                c[j] = 0; x[j] = 0; y[j] = 0; z[j] = 0;
                if (i < n) {
                    c[j] = perm[i];
                    const u32 kidx = keys_sorted ? i : c[j];
                    x[j] = kx[kidx]; y[j] = ky[kidx]; z[j] = kz[kidx];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < UQ_ITEMS; j++) {
            const u32 i = i0 + j;
            if (i < n) {
                const bool nw = (i == 0) || x[j] != px || y[j] != py || z[j] != pz;
                px = x[j]; py = y[j]; pz = z[j];
                if (nw) {
                    isnew |= 1u << j;
                    cnt++;
                    nlo += x[j] < key_lo;
                    nhi += x[j] < key_hi;
                }
            }
        }
        u32 tot;
        const u32 ex = block_exclusive_scan(cnt, &tot, sw);
        if (threadIdx.x < 32) {
            u32 pre = lookback_exclusive(desc, 1, tile, tot, 1u);
            if (threadIdx.x == 0) s_pre = pre;
        }
        for (int o = 16; o > 0; o >>= 1) {
            nlo += __shfl_xor_sync(0xffffffffu, nlo, o);
            nhi += __shfl_xor_sync(0xffffffffu, nhi, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (nlo) atomicAdd(&counters[C_NLO], nlo);
            if (nhi) atomicAdd(&counters[C_NHI], nhi);
        }
        u32 li = ex;             // local index of the thread's next new vertex inside the tile
#pragma unroll
        for (int j = 0; j < UQ_ITEMS; j++)
            if ((isnew >> j) & 1u) {
                sv[3 * li + 0] = key_float(x[j]);
                sv[3 * li + 1] = key_float(y[j]);
                sv[3 * li + 2] = key_float(z[j]);
                li++;
            }
        __syncthreads();
        const u32 pre = s_pre;
        float *dst = V + 3 * (size_t) pre;
        for (u32 i = threadIdx.x; i < 3 * tot; i += blockDim.x) dst[i] = sv[i];
        u32 rank = pre + ex;   // rank of the next new vertex
#pragma unroll
        for (int j = 0; j < UQ_ITEMS; j++) {
            const u32 i = i0 + j;
            if (i < n) {
                if ((isnew >> j) & 1u) rank++;
                if (c[j] < n) cand_rank[c[j]] = rank - 1;   // (never out of range with a complete permutation)
                if (i == n - 1) counters[C_V] = rank;
            }
        }
    }
}


// x < thr  <=>  key(x) < key(thr) for non-NaN values (host-side twin of float_key)
static inline u32 host_float_key(float f) {
    u32 b;
    memcpy(&b, &f, 4);
    if ((b << 1) == 0u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

}   // namespace isx
