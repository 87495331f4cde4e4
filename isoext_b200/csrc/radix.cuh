// radix.cuh -- hand-written single-sweep LSD radix sort of n items by a 96-bit key (kx, ky, kz).
//
// Why it exists: the reference defines vertex ids by a lexicographic (x, y, z) sort of the welded
// positions (src/utils.cu:49-55 + include/math.cuh:112-126 of the reference), so bit-exact face
// connectivity needs the same order.  The sort is surface-sized (n = #vertices) and lives in L2.
//
// Structure (12 digit places of 8 bits: z bytes, then y bytes, then x bytes):
//   k_radix_hist    one read of the keys -> 12 global 256-bin histograms
//   k_radix_prefix  exclusive scan of each histogram
//   k_radix_pass    x12: each block takes a tile by ticket, ranks its items stably with warp
//                   match_any + per-warp counters, resolves the tile's base per bin with a
//                   per-bin decoupled look-back, and scatters (key, perm) pairs.
// The descriptor array is reused across passes through an epoch tag (no re-zeroing).
#pragma once
#include "common.cuh"

namespace isx {

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_ITEMS = 8;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;
constexpr int RADIX_PASSES = 12;

struct RadixBuffers {
    u32 *key[2];    // n each
    u32 *perm[2];   // n each
    u32 *hist;      // 12 * 256  (becomes exclusive prefixes)
    u32 *ticket;    // 12
    u64 *desc;      // ntiles * 256
    static size_t carve(Carver &c, size_t n, RadixBuffers *out) {
        size_t ntiles = (n + RADIX_TILE - 1) / RADIX_TILE + 1;
        RadixBuffers b;
        b.key[0] = c.take<u32>(n);
        b.key[1] = c.take<u32>(n);
        b.perm[0] = c.take<u32>(n);
        b.perm[1] = c.take<u32>(n);
        b.hist = c.take<u32>(RADIX_PASSES * 256 + 16);   // hist followed by tickets: one memset
        b.ticket = b.hist + RADIX_PASSES * 256;
        b.desc = c.take<u64>(ntiles * 256);
        if (out) *out = b;
        return c.bytes();
    }
};

static __global__ void __launch_bounds__(256) k_radix_hist(const u32 *__restrict__ kx, const u32 *__restrict__ ky,
                                                    const u32 *__restrict__ kz, u32 n, u32 *__restrict__ hist,
                                                    const u32 *__restrict__ n_dev) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sh[RADIX_PASSES * 256];
    if (n_dev) n = *n_dev;
    for (int i = threadIdx.x; i < RADIX_PASSES * 256; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u32 k[3] = {kz[i], ky[i], kx[i]};
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int b = 0; b < 4; b++) atomicAdd(&sh[(c * 4 + b) * 256 + ((k[c] >> (8 * b)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX_PASSES * 256; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// one warp per digit place: exclusive scan of its 256-bin histogram (8 bins per lane)
static __global__ void __launch_bounds__(32 * RADIX_PASSES) k_radix_prefix(u32 *__restrict__ hist) {
    pdl_wait();
    pdl_trigger();
    const u32 lane = threadIdx.x & 31, p = threadIdx.x >> 5;
    u32 v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        v[k] = hist[p * 256 + lane * 8 + k];
        sum += v[k];
    }
    u32 incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32) o) incl += t;
    }
    u32 run = incl - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        hist[p * 256 + lane * 8 + k] = run;
        run += v[k];
    }
}

// Per-bin look-back by a single thread.  desc index = tile * 256 + bin.
__device__ __forceinline__ u32 lookback_bin(u64 *desc, u32 tile, u32 bin, u32 aggregate, u32 epoch) {
    u64 *mine = desc + (size_t) tile * 256 + bin;
    if (tile == 0) {
        st_relaxed_u64(mine, desc_pack(2, epoch, aggregate));
        return 0;
    }
    st_relaxed_u64(mine, desc_pack(1, epoch, aggregate));
    u32 excl = 0;
    // Walk back with 4 independent loads in flight (the walk is latency-bound, not bandwidth-bound).
    for (int t = (int) tile - 1; t >= 0; t -= 4) {
        u64 d[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            d[k] = (t - k >= 0) ? ld_relaxed_u64(desc + (size_t) (t - k) * 256 + bin) : desc_pack(2, epoch, 0);
        bool done = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (done) break;
            u64 dk = d[k];
            u32 state = (((u32) (dk >> 32)) & 0x3fffffffu) == (epoch & 0x3fffffffu) ? (u32) (dk >> 62) : 0u;
            while (state == 0) {
                dk = ld_relaxed_u64(desc + (size_t) (t - k) * 256 + bin);
                state = (((u32) (dk >> 32)) & 0x3fffffffu) == (epoch & 0x3fffffffu) ? (u32) (dk >> 62) : 0u;
            }
            excl += (u32) dk;
            if (state == 2) done = true;
        }
        if (done) break;
    }
    st_relaxed_u64(mine, desc_pack(2, epoch, excl + aggregate));
    return excl;
}

// One digit pass.  GATHER: the key of item i is src_keys[perm] (start of a new coordinate);
// IDENTITY: perm_in is the identity (very first pass).
template <bool GATHER, bool IDENTITY>
static __global__ void __launch_bounds__(RADIX_THREADS)
k_radix_pass(const u32 *__restrict__ key_in, const u32 *__restrict__ perm_in, u32 *__restrict__ key_out,
             u32 *__restrict__ perm_out, const u32 *__restrict__ src_keys, u32 shift,
             const u32 *__restrict__ gprefix, u64 *__restrict__ desc, u32 *__restrict__ ticket, u32 epoch, u32 n,
             const u32 *__restrict__ n_dev) {
    __shared__ u32 cnt[RADIX_THREADS / 32][256];
    pdl_wait();                  // launched with programmatic serialization behind the previous pass (common.cuh)
    pdl_trigger();
    if (n_dev) n = *n_dev;       // device-side count: blocks beyond the last tile leave right after their ticket
    __shared__ u32 gbase[256];
    __shared__ u32 s_tile;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < (RADIX_THREADS / 32) * 256; i += RADIX_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    if (tile * RADIX_TILE >= n) return;
    const u32 base = tile * RADIX_TILE + warp * (32 * RADIX_ITEMS);

    u32 key[RADIX_ITEMS], perm[RADIX_ITEMS], rank[RADIX_ITEMS];
#pragma unroll
    for (int j = 0; j < RADIX_ITEMS; j++) {
        u32 idx = base + j * 32 + lane;
        bool valid = idx < n;
        if (valid) {
            perm[j] = IDENTITY ? idx : perm_in[idx];
            key[j] = GATHER ? src_keys[perm[j]] : key_in[idx];
        } else {
            perm[j] = 0;
            key[j] = 0xffffffffu;
        }
    }
#pragma unroll
    for (int j = 0; j < RADIX_ITEMS; j++) {
        u32 idx = base + j * 32 + lane;
        bool valid = idx < n;
        u32 d = (key[j] >> shift) & 255u;
        u32 active = __ballot_sync(0xffffffffu, valid);
        rank[j] = 0;
        if (valid) {
            u32 peers = __match_any_sync(active, d);
            u32 leader = __ffs(peers) - 1;
            u32 old = 0;
            if (lane == leader) {
                old = cnt[warp][d];
                cnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(peers, old, leader);
            rank[j] = old + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
    }
    __syncthreads();
    {
        // thread `bin`: exclusive scan over the warps, then the tile's global base for this bin
        const u32 bin = threadIdx.x;
        u32 sum = 0;
#pragma unroll
        for (int w = 0; w < RADIX_THREADS / 32; w++) {
            u32 t = cnt[w][bin];
            cnt[w][bin] = sum;
            sum += t;
        }
        u32 excl = lookback_bin(desc, tile, bin, sum, epoch);
        gbase[bin] = gprefix[bin] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RADIX_ITEMS; j++) {
        u32 idx = base + j * 32 + lane;
        if (idx < n) {
            u32 d = (key[j] >> shift) & 255u;
            u32 pos = gbase[d] + cnt[warp][d] + rank[j];
            key_out[pos] = key[j];
            perm_out[pos] = perm[j];
        }
    }
}

// Enqueue the whole sort on `stream`.  Result: b.perm[0] holds the sorted order (item ids).
// kx/ky/kz: order-preserving u32 keys (float_key).  n may be 0.  If n_dev != nullptr the item count is read
// on the device (it must be <= n, which then only sizes the launches and the descriptor reset).
// first_pass (0, 4 or 8): skip the digit places of kz (and ky): 4 = 64-bit keys (kx, ky), 8 = 32-bit keys (kx).
static inline cudaError_t radix_sort96(const u32 *kx, const u32 *ky, const u32 *kz, u32 n, const RadixBuffers &b,
                                       cudaStream_t stream, const u32 *n_dev = nullptr, int first_pass = 0) {
    if (n == 0) return cudaSuccess;
    const u32 ntiles = (n + RADIX_TILE - 1) / RADIX_TILE;
    cudaError_t e = cudaMemsetAsync(b.hist, 0, (RADIX_PASSES * 256 + 16) * sizeof(u32), stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.desc, 0, (size_t) ntiles * 256 * sizeof(u64), stream);
    if (e != cudaSuccess) return e;
    u32 hist_blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (hist_blocks > 148 * 4) hist_blocks = 148 * 4;
    ISX_LAUNCH(k_radix_hist, hist_blocks, 256, 0, stream, kx, ky, kz, n, b.hist, n_dev);
    ISX_LAUNCH_PDL(k_radix_prefix, 1, 32 * RADIX_PASSES, 0, stream, b.hist);
    const u32 *src[3] = {kz, ky, kx};
    for (int p = first_pass; p < RADIX_PASSES; p++) {
        const int in = p & 1, out = in ^ 1;
        const u32 shift = 8 * (p & 3);
        const u32 *coord = src[p >> 2];
        if (p == first_pass)
            ISX_LAUNCH_PDL((k_radix_pass<true, true>), ntiles, RADIX_THREADS, 0, stream, (const u32 *) b.key[in], (const u32 *) b.perm[in], b.key[out], b.perm[out], coord,
                           shift, (const u32 *) (b.hist + p * 256), b.desc, b.ticket + p, (u32) (p + 1), n, n_dev);
        else if ((p & 3) == 0)
            ISX_LAUNCH_PDL((k_radix_pass<true, false>), ntiles, RADIX_THREADS, 0, stream, (const u32 *) b.key[in], (const u32 *) b.perm[in], b.key[out], b.perm[out], coord,
                           shift, (const u32 *) (b.hist + p * 256), b.desc, b.ticket + p, (u32) (p + 1), n, n_dev);
        else
            ISX_LAUNCH_PDL((k_radix_pass<false, false>), ntiles, RADIX_THREADS, 0, stream, (const u32 *) b.key[in], (const u32 *) b.perm[in], b.key[out], b.perm[out], coord,
                           shift, (const u32 *) (b.hist + p * 256), b.desc, b.ticket + p, (u32) (p + 1), n, n_dev);
    }
    return cudaGetLastError();
}

}   // namespace isx
