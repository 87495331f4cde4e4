// segsort.cuh -- segmented sort of vertex candidates by 96-bit position key, exploiting grid structure.
//
// The reference's vertex order is lexicographic (x, y, z) on float positions.  Every candidate knows
// the x-plane interval [px[b], px[b+1]) that contains its x coordinate *exactly* (a float compare
// against the two neighbouring plane positions), and b is a monotone function of x, so
//     global order = concatenation over b of (bucket b sorted by full key).
// Buckets are small for curved surfaces (~V/X elements), so one thread block sorts a bucket entirely
// in shared memory (bitonic network on 96-bit keys).  Buckets larger than SEG_CAP (an axis-aligned
// face lying inside one slab) are sorted together by the global radix sort (radix.cuh); their
// elements are already grouped by bucket in that sort's output because x decides the bucket.
//
//   k_seg_count    histogram of bucket ids
//   k_seg_scan     exclusive scan of the histogram (one block), big-bucket bookkeeping
//   k_seg_scatter  group candidate ids by bucket (order inside a bucket is irrelevant: it is sorted next)
//   k_seg_sort     one block per bucket: load keys, bitonic sort in smem, write sorted ids
//   (fallback)     k_seg_big_gather -> radix_sort96 -> k_seg_big_scatter
// 4 launches instead of 14 in the common case.
#pragma once
#include "common.cuh"
#include "radix.cuh"

namespace isx {

constexpr int SEG_CAP = 4096;        // largest bucket sorted in shared memory
constexpr int SEG_THREADS = 512;

// bucket bookkeeping (lives in the phase-1 workspace: the histogram is filled while scanning the entries)
struct SegHead {
    u32 *count;      // nb + 1
    u32 *start;      // nb + 1
    u32 *cursor;     // nb + 1
    u32 *bigoff;     // nb + 1   (offset of a big bucket inside the compacted big list)
    static size_t words(size_t nb) { return 4 * (nb + 1); }
    static void carve(Carver &c, size_t nb, SegHead *out) {
        SegHead h;
        h.count = c.take<u32>(words(nb));
        h.start = h.count + (nb + 1);
        h.cursor = h.start + (nb + 1);
        h.bigoff = h.cursor + (nb + 1);
        if (out) *out = h;
    }
};
// per-candidate buffers (phase-2 scratch)
struct SegScratch {
    u32 *cbucket;    // n   bucket id of every candidate
    u32 *perm0;      // n   ids grouped by bucket
    u32 *perm;       // n   ids in final sorted order
    u32 *bkx, *bky, *bkz, *bid;   // fallback: compacted keys / ids of the big buckets
    RadixBuffers radix;
    static void carve(Carver &c, size_t n, SegScratch *out) {
        SegScratch b;
        b.cbucket = c.take<u32>(n);
        b.perm0 = c.take<u32>(n);
        b.perm = c.take<u32>(n);
        b.bkx = c.take<u32>(n);
        b.bky = c.take<u32>(n);
        b.bkz = c.take<u32>(n);
        b.bid = c.take<u32>(n);
        RadixBuffers::carve(c, n, &b.radix);
        if (out) *out = b;
    }
};

static __global__ void __launch_bounds__(256) k_seg_count(const u32 *__restrict__ cbucket, u32 n, u32 *__restrict__ count) {
    const u32 lane = threadIdx.x & 31;
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const u32 i = base + lane;
        const bool valid = i < n;
        const u32 active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            // consecutive candidates mostly share a bucket: aggregate per warp before touching L2
            const u32 b = cbucket[i];
            const u32 peers = __match_any_sync(active, b);
            if (lane == (u32) (__ffs(peers) - 1)) atomicAdd(&count[b], (u32) __popc(peers));
        }
    }
}

// one block: exclusive scan over nb buckets; big buckets (> SEG_CAP) get offsets in the compacted big list
static __global__ void __launch_bounds__(1024) k_seg_scan(u32 nb, const u32 *__restrict__ count, u32 *__restrict__ start,
                                                           u32 *__restrict__ cursor, u32 *__restrict__ bigoff, u32 *__restrict__ info_nbig,
                                                           u32 *__restrict__ info_max) {
    __shared__ u32 sw[33];
    __shared__ u32 s_carry, s_carry_big, s_max, s_nbig;
    if (threadIdx.x == 0) { s_carry = 0; s_carry_big = 0; s_max = 0; s_nbig = 0; }
    __syncthreads();
    for (u32 base = 0; base < nb; base += 1024) {
        const u32 b = base + threadIdx.x;
        const u32 c = b < nb ? count[b] : 0u;
        const u32 cb = c > (u32) SEG_CAP ? c : 0u;
        u32 tot, totb;
        const u32 ex = block_exclusive_scan(c, &tot, sw);
        const u32 exb = block_exclusive_scan(cb, &totb, sw);
        if (b < nb) {
            start[b] = s_carry + ex;
            cursor[b] = 0;
            bigoff[b] = s_carry_big + exb;
            if (c) atomicMax(&s_max, c);
            if (cb) atomicAdd(&s_nbig, 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry += tot; s_carry_big += totb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        start[nb] = s_carry;
        *info_nbig = s_carry_big;   // number of candidates living in big buckets
        *info_max = s_max;          // largest bucket
    }
}

static __global__ void __launch_bounds__(256) k_seg_scatter(const u32 *__restrict__ cbucket, u32 n, const u32 *__restrict__ start,
                                                            u32 *__restrict__ cursor, u32 *__restrict__ perm0,
                                                            const u32 *__restrict__ n_dev, u32 n_cap) {
    if (n_dev) n = *n_dev;
    if (n > n_cap) return;
    const u32 lane = threadIdx.x & 31;
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const u32 i = base + lane;
        const bool valid = i < n;
        const u32 active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const u32 b = cbucket[i];
            const u32 peers = __match_any_sync(active, b);
            const u32 leader = __ffs(peers) - 1;
            u32 off = 0;
            if (lane == leader) off = atomicAdd(&cursor[b], (u32) __popc(peers));
            off = __shfl_sync(peers, off, leader);
            perm0[start[b] + off + __popc(peers & ((1u << lane) - 1u))] = i;
        }
    }
}

__device__ __forceinline__ bool key_less(u32 ax, u32 ay, u32 az, u32 bx, u32 by, u32 bz) {
    return ax < bx || (ax == bx && (ay < by || (ay == by && az < bz)));
}

struct SKey { u32 x, y, z, id; };
// (x,y,z,id) as one 128-bit big-endian integer: a < b  <=>  the subtraction a - b borrows.
// 5 instructions instead of a branchy lexicographic compare; ids are distinct, so the order is total.
__device__ __forceinline__ u32 skey_less(const SKey &a, const SKey &b) {
    u32 r;
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 t, %1, %5;\n\t"
        "subc.cc.u32 t, %2, %6;\n\t"
        "subc.cc.u32 t, %3, %7;\n\t"
        "subc.cc.u32 t, %4, %8;\n\t"
        "subc.u32 %0, 0, 0;\n\t}"
        : "=r"(r)
        : "r"(a.id), "r"(a.z), "r"(a.y), "r"(a.x), "r"(b.id), "r"(b.z), "r"(b.y), "r"(b.x));
    return r;   // 0xffffffff if a < b, else 0
}

// compare-exchange of `mine` with `other` (the element at index e ^ j): keep the smaller one if
// keep_min, else the larger; equal elements (only the padding) keep their own copy.
__device__ __forceinline__ void cmpx(SKey &mine, const SKey &other, bool keep_min) {
    const u32 take = keep_min ? skey_less(other, mine) : skey_less(mine, other);
    if (take) mine = other;
}

// Bitonic sort of m = SEG_THREADS * E elements, element e = tid * E + r held in registers.
// Stages with j < E are thread-local, E <= j < 32 E use warp shuffles, wider ones go through smem.
template <int E>
__device__ __forceinline__ void bitonic_block(SKey (&v)[E], u32 m, u32 *sx, u32 *sy, u32 *sz, u32 *si) {
    const u32 tid = threadIdx.x;
    for (u32 k = 2; k <= m; k <<= 1) {
        for (u32 j = k >> 1; j > 0; j >>= 1) {
            if (j < (u32) E) {
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const int q = r ^ (int) j;
                    if (q > r) {
                        const u32 e = tid * E + r;
                        const bool up = (e & k) == 0;
                        SKey a = v[r], b = v[q];
                        cmpx(v[r], b, up);
                        cmpx(v[q], a, !up);
                    }
                }
            } else if (j < 32u * E) {
                const u32 lj = j / E;   // lane distance
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const u32 e = tid * E + r;
                    SKey o;
                    o.x = __shfl_xor_sync(0xffffffffu, v[r].x, lj);
                    o.y = __shfl_xor_sync(0xffffffffu, v[r].y, lj);
                    o.z = __shfl_xor_sync(0xffffffffu, v[r].z, lj);
                    o.id = __shfl_xor_sync(0xffffffffu, v[r].id, lj);
                    const bool lower = (e & j) == 0, up = (e & k) == 0;
                    cmpx(v[r], o, lower == up);
                }
            } else {
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const u32 e = tid * E + r;
                    sx[e] = v[r].x; sy[e] = v[r].y; sz[e] = v[r].z; si[e] = v[r].id;
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < E; r++) {
                    const u32 e = tid * E + r, q = e ^ j;
                    SKey o = {sx[q], sy[q], sz[q], si[q]};
                    const bool lower = (e & j) == 0, up = (e & k) == 0;
                    cmpx(v[r], o, lower == up);
                }
            }
        }
    }
}

template <int E>
__device__ __forceinline__ void seg_sort_bucket(const u32 *__restrict__ kx, const u32 *__restrict__ ky, const u32 *__restrict__ kz,
                                                const u32 *__restrict__ perm0, u32 *__restrict__ perm, u32 s0, u32 n, u32 *smem) {
    u32 *sx = smem, *sy = smem + SEG_CAP, *sz = smem + 2 * SEG_CAP, *si = smem + 3 * SEG_CAP;
    SKey v[E];
#pragma unroll
    for (int r = 0; r < E; r++) {
        const u32 e = threadIdx.x * E + r;
        if (e < n) {
            const u32 id = perm0[s0 + e];
            v[r].x = kx[id]; v[r].y = ky[id]; v[r].z = kz[id]; v[r].id = id;
        } else {
            v[r].x = 0xffffffffu; v[r].y = 0xffffffffu; v[r].z = 0xffffffffu; v[r].id = 0xffffffffu;
        }
    }
    bitonic_block<E>(v, SEG_THREADS * E, sx, sy, sz, si);
#pragma unroll
    for (int r = 0; r < E; r++) {
        const u32 e = threadIdx.x * E + r;
        if (e < n) perm[s0 + e] = v[r].id;
    }
}

// one block per bucket; dynamic smem = 4 arrays of SEG_CAP u32
static __global__ void __launch_bounds__(SEG_THREADS) k_seg_sort(const u32 *__restrict__ kx, const u32 *__restrict__ ky,
                                                                 const u32 *__restrict__ kz, const u32 *__restrict__ count,
                                                                 const u32 *__restrict__ start, const u32 *__restrict__ perm0,
                                                                 u32 *__restrict__ perm, const u32 *__restrict__ n_dev, u32 n_cap) {
    extern __shared__ u32 smem[];
    if (n_dev && *n_dev > n_cap) return;
    const u32 b = blockIdx.x;
    const u32 n = count[b];
    if (n == 0 || n > (u32) SEG_CAP) return;
    const u32 s0 = start[b];
    if (n == 1) {
        if (threadIdx.x == 0) perm[s0] = perm0[s0];
        return;
    }
    if (n <= SEG_THREADS) seg_sort_bucket<1>(kx, ky, kz, perm0, perm, s0, n, smem);
    else if (n <= 2 * SEG_THREADS) seg_sort_bucket<2>(kx, ky, kz, perm0, perm, s0, n, smem);
    else if (n <= 4 * SEG_THREADS) seg_sort_bucket<4>(kx, ky, kz, perm0, perm, s0, n, smem);
    else seg_sort_bucket<8>(kx, ky, kz, perm0, perm, s0, n, smem);
}

// ---- fallback for big buckets -------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_seg_big_gather(u32 nb, const u32 *__restrict__ count, const u32 *__restrict__ start,
                                                               const u32 *__restrict__ bigoff, const u32 *__restrict__ perm0,
                                                               const u32 *__restrict__ kx, const u32 *__restrict__ ky,
                                                               const u32 *__restrict__ kz, u32 *__restrict__ bkx, u32 *__restrict__ bky,
                                                               u32 *__restrict__ bkz, u32 *__restrict__ bid) {
    for (u32 b = blockIdx.y; b < nb; b += gridDim.y) {
        const u32 n = count[b];
        if (n <= (u32) SEG_CAP) continue;
        const u32 s0 = start[b], o0 = bigoff[b];
        for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const u32 id = perm0[s0 + i];
            bkx[o0 + i] = kx[id]; bky[o0 + i] = ky[id]; bkz[o0 + i] = kz[id]; bid[o0 + i] = id;
        }
    }
}
// sorted rank r of the big list -> final position: the big list is ordered by bucket (x decides the bucket)
static __global__ void __launch_bounds__(256) k_seg_big_scatter(u32 n_big, const u32 *__restrict__ sorted, const u32 *__restrict__ bid,
                                                                const u32 *__restrict__ cbucket, const u32 *__restrict__ start,
                                                                const u32 *__restrict__ bigoff, u32 *__restrict__ perm) {
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_big; r += gridDim.x * blockDim.x) {
        const u32 id = bid[sorted[r]];
        const u32 b = cbucket[id];
        perm[start[b] + (r - bigoff[b])] = id;
    }
}

// Phase-2 part of the segmented sort: group by bucket, sort the small buckets in shared memory, and --
// if the host learned at the phase-1 sync that big buckets exist (n_big > 0) -- run the radix fallback.
// The histogram (h.count) and its scan (h.start / h.bigoff) were produced in phase 1.
// n_dev != nullptr: the item count is read on the device (must be <= n_cap); grid_n sizes the launches.
static inline cudaError_t seg_sort_run(const u32 *kx, const u32 *ky, const u32 *kz, u32 n, const u32 *n_dev, u32 n_cap, u32 grid_n,
                                       u32 nb, u32 n_big, const SegHead &h, const SegScratch &b, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_seg_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * SEG_CAP * (int) sizeof(u32));
        attr_set = true;
    }
    const int blocks = (int) ((grid_n + 255) / 256 > 148 * 8 ? 148 * 8 : (grid_n + 255) / 256);
    ISX_LAUNCH(k_seg_scatter, blocks < 1 ? 1 : blocks, 256, 0, stream, b.cbucket, n, h.start, h.cursor, b.perm0, n_dev, n_cap);
    ISX_LAUNCH(k_seg_sort, nb, SEG_THREADS, 4 * SEG_CAP * sizeof(u32), stream, kx, ky, kz, h.count, h.start, b.perm0, b.perm, n_dev,
               n_cap);
    if (n_big > 0) {
        dim3 grid(32, nb < 1024 ? nb : 1024);
        ISX_LAUNCH(k_seg_big_gather, grid, 256, 0, stream, nb, h.count, h.start, h.bigoff, b.perm0, kx, ky, kz, b.bkx, b.bky, b.bkz, b.bid);
        cudaError_t e = radix_sort96(b.bkx, b.bky, b.bkz, n_big, b.radix, stream);
        if (e != cudaSuccess) return e;
        const int bb = (int) ((n_big + 255) / 256 > 148 * 8 ? 148 * 8 : (n_big + 255) / 256);
        ISX_LAUNCH(k_seg_big_scatter, bb, 256, 0, stream, n_big, b.radix.perm[0], b.bid, b.cbucket, h.start, h.bigoff, b.perm);
    }
    return cudaGetLastError();
}

}   // namespace isx
