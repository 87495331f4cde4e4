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
//   (histogram of bucket ids: built by the caller while it scans its entries, mc_dense.cu k_scan_entries)
//   k_seg_scan     exclusive scan of the histogram (one block), big-bucket bookkeeping
//   k_seg_scatter  group candidate ids by bucket (order inside a bucket is irrelevant: it is sorted next);
//                  elements of oversized buckets go straight to a compacted list of keys
//   k_seg_sort     one block per bucket: load keys, bitonic sort in smem, write sorted ids
//   (oversized)    k_big_plan -> k_big_sub -> k_seg_scan -> k_seg_scatter -> k_seg_sort:
//                  a second level of buckets inside every oversized bucket (an axis-aligned face: nearly all
//                  vertices share one x, so cut that value by y-plane and the rest of the x range uniformly)
//   (last resort)  radix_sort96 over the oversized buckets if a second-level bucket is still too large
// 4 launches instead of 14 in the common case.
#pragma once
#include "common.cuh"
#include "dense.cuh"   // Counter
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
    u32 *xinvmin;    // nb + 1   max of ~xkey over a big bucket (zero-initialised, so a max)
    u32 *xmax;       // nb + 1   max of xkey over a big bucket
    u32 *sub_base;   // nb + 1   first second-level bucket of a big bucket
    u32 *sub_par;    // nb + 1   how finely a big bucket is cut (k_big_plan)
    u32 *sub_xm;     // nb + 1   the dominant x key of a big bucket
    static size_t words(size_t nb) { return 9 * (nb + 1); }
    static void carve(Carver &c, size_t nb, SegHead *out) {
        SegHead h;
        h.count = c.take<u32>(words(nb));
        h.start = h.count + (nb + 1);
        h.cursor = h.start + (nb + 1);
        h.bigoff = h.cursor + (nb + 1);
        h.xinvmin = h.bigoff + (nb + 1);
        h.xmax = h.xinvmin + (nb + 1);
        h.sub_base = h.xmax + (nb + 1);
        h.sub_par = h.sub_base + (nb + 1);
        h.sub_xm = h.sub_par + (nb + 1);
        if (out) *out = h;
    }
};
// per-candidate buffers (phase-2 scratch)
struct SegScratch {
    u32 *cbucket;    // n   bucket id of every candidate
    u32 *perm0;      // n   ids grouped by bucket
    u32 *perm;       // n   ids in final sorted order
    u32 *skx, *sky, *skz;         // n   keys in final sorted order (so that the weld pass reads sequentially)
    u32 *bkx, *bky, *bkz, *bid;   // compacted keys / ids of the big buckets
    u32 *csub, *perm2;            // second level: bucket of every big-list element, big-list ids grouped by it
    u32 *info2;                   // [8]: 0 second-level buckets, 1 elements in oversized ones, 2 largest, 3 overflow, 4 radix count, 5 no level 2, 6 abort
    u32 *count2;                  // 4 * (subcap + 1): count / start / cursor / bigoff of the second-level buckets
    u32 subcap;
    RadixBuffers radix;
    static u32 sub_capacity(size_t n) { return (u32) (n / 64 + 64); }   // k_big_plan: nsub <= 3 * count / 512 per big bucket
    static void carve(Carver &c, size_t n, SegScratch *out) {
        SegScratch b;
        b.cbucket = c.take<u32>(n);
        b.perm0 = c.take<u32>(n);
        b.perm = c.take<u32>(n);
        b.skx = c.take<u32>(n);
        b.sky = c.take<u32>(n);
        b.skz = c.take<u32>(n);
        b.bkx = c.take<u32>(n);
        b.bky = c.take<u32>(n);
        b.bkz = c.take<u32>(n);
        b.bid = c.take<u32>(n);
        b.csub = c.take<u32>(n);
        b.perm2 = c.take<u32>(n);
        b.subcap = sub_capacity(n);
        b.info2 = c.take<u32>(8 + 4 * ((size_t) b.subcap + 1));   // info2 and count2 are cleared by one memset
        b.count2 = b.info2 + 8;
        RadixBuffers::carve(c, n, &b.radix);
        if (out) *out = b;
    }
};

// Executed by ONE block (any blockDim that is a multiple of 32, <= 1024): exclusive scan over nb buckets
// (SS_ITEMS consecutive buckets per thread); big buckets (> SEG_CAP) also get offsets in the compacted big
// list.  `count` is read with L2 loads (it was produced by atomics of other blocks).
constexpr int SS_ITEMS = 16;
__device__ __forceinline__ void seg_scan_block(u32 nb, const u32 *__restrict__ count, u32 *__restrict__ start, u32 *__restrict__ cursor,
                                               u32 *__restrict__ bigoff, u32 *__restrict__ info_nbig, u32 *__restrict__ info_max,
                                               u32 *sw /* >= 33 */, u32 *s4 /* >= 4 */) {
    if (threadIdx.x == 0) { s4[0] = 0; s4[1] = 0; s4[2] = 0; }
    __syncthreads();
    for (u32 base = 0; base < nb; base += blockDim.x * SS_ITEMS) {
        const u32 b0 = base + threadIdx.x * SS_ITEMS;
        u32 c[SS_ITEMS], sum = 0, sumb = 0, mx = 0;
#pragma unroll
        for (int j = 0; j < SS_ITEMS; j++) {
            c[j] = b0 + j < nb ? __ldcg(count + b0 + j) : 0u;
            sum += c[j];
            sumb += c[j] > (u32) SEG_CAP ? c[j] : 0u;
            mx = c[j] > mx ? c[j] : mx;
        }
        u32 tot, totb;
        u32 ex = block_exclusive_scan(sum, &tot, sw) + s4[0];
        u32 exb = block_exclusive_scan(sumb, &totb, sw) + s4[1];
#pragma unroll
        for (int j = 0; j < SS_ITEMS; j++)
            if (b0 + j < nb) {
                start[b0 + j] = ex;
                cursor[b0 + j] = 0;
                bigoff[b0 + j] = exb;
                ex += c[j];
                exb += c[j] > (u32) SEG_CAP ? c[j] : 0u;
            }
        if (mx) atomicMax(&s4[2], mx);
        __syncthreads();
        if (threadIdx.x == 0) { s4[0] += tot; s4[1] += totb; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        start[nb] = s4[0];
        *info_nbig = s4[1];   // number of candidates living in big buckets
        *info_max = s4[2];    // largest bucket
    }
}

static __global__ void __launch_bounds__(1024) k_seg_scan(u32 nb, const u32 *__restrict__ count, u32 *__restrict__ start,
                                                           u32 *__restrict__ cursor, u32 *__restrict__ bigoff, u32 *__restrict__ info_nbig,
                                                           u32 *__restrict__ info_max, const u32 *__restrict__ nb_dev = nullptr) {
    __shared__ u32 sw[33];
    __shared__ u32 s4[4];
    if (nb_dev && *nb_dev < nb) nb = *nb_dev;   // buckets in use (the rest of the table is empty)
    seg_scan_block(nb, count, start, cursor, bigoff, info_nbig, info_max, sw, s4);
}

// Group candidate ids by bucket.  WITH_BIG (first level): the elements of oversized buckets go straight to the
// compacted big list (keys + ids, at bigoff[b] + position inside the bucket) and the x range of every oversized
// bucket is accumulated for k_big_plan -- the order inside a bucket is irrelevant, it is sorted next.
struct SegBigOut {
    const u32 *count, *bigoff, *kx, *ky, *kz;
    u32 *bkx, *bky, *bkz, *bid, *xinvmin, *xmax;
};
template <bool WITH_BIG>
static __global__ void __launch_bounds__(256) k_seg_scatter(const u32 *__restrict__ cbucket, u32 n, const u32 *__restrict__ start,
                                                            u32 *__restrict__ cursor, u32 *__restrict__ perm0,
                                                            const u32 *__restrict__ n_dev, u32 n_cap,
                                                            const u32 *__restrict__ skip, SegBigOut big) {
    if (skip && *skip) return;
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
            off = __shfl_sync(peers, off, leader) + __popc(peers & ((1u << lane) - 1u));
            if (WITH_BIG && big.count[b] > (u32) SEG_CAP) {
                const u32 o = big.bigoff[b] + off, x = big.kx[i];
                big.bkx[o] = x; big.bky[o] = big.ky[i]; big.bkz[o] = big.kz[i]; big.bid[o] = i;
                const u32 mx = __reduce_max_sync(peers, x), mn = __reduce_max_sync(peers, ~x);
                if (lane == leader) {
                    atomicMax(&big.xmax[b], mx);
                    atomicMax(&big.xinvmin[b], mn);
                }
            } else {
                perm0[start[b] + off] = i;
            }
        }
    }
}

// Grid facts the sort uses: plane positions along x and y and how the first-level buckets were cut (mc_dense.cu
// sub_bucket): bucket id = layer * (gy + gx) + sub, layer = local x plane + 1 (x_off = global index of local plane 0).
struct SegGeom {
    float amin_x, asize_x, amin_y, asize_y, amin_z, asize_z;
    u32 Xg, Y, Z;        // global points along x, points along y and z
    i64 x_off;
    u32 gy, gx, ystep;
    bool grouped;        // the group stage of k_seg_sort is usable (ystep + 2 <= SEG_GROUPS)
};
// index of the last plane at or below v along an axis (0 if v is below all planes): exact float compares, monotone in v
__device__ __forceinline__ u32 plane_index(float v, float amin, float asize, u32 res) {
    float f = __fdiv_rn(__fsub_rn(v, amin), asize) * (float) res;
    int j = f > 0.f ? (f < (float) res ? (int) f : (int) res) : 0;
    while (j > 0 && axis_pos((u32) j, res, amin, asize) > v) j--;
    while (j < (int) res && axis_pos((u32) j + 1, res, amin, asize) <= v) j++;
    return (u32) j;
}
__device__ __forceinline__ u32 yplane_of(float y, const SegGeom &g) { return plane_index(y, g.amin_y, g.asize_y, g.Y - 1); }

// One block per bucket (block-stride loop): LSD radix sort (8-bit digits) entirely in shared memory.
//   keys stay in place (sk[3][CAP]); only 16-bit local indices move (ord ping-pong);
//   ranking is stable: each warp owns a contiguous run of positions, ranks 32 of them at a time with
//   match_any against its private digit counters, then a (digit, warp) exclusive scan gives the bases;
//   digit places on which all keys of the bucket agree (typically the high bytes of x) are skipped.
// Two instances by bucket size: <SEG_SMALL, 128> (25 KB, 8 blocks/SM) and <SEG_CAP, 512> (96 KB, 2 blocks/SM).
// (A third <2048, 256> instance was measured slower overall: 209 vs 181 us at 1024^3.)
// Group stage (first level only), tried before the radix passes: inside a bucket all keys share x exactly (vertices
// lying on an x plane) or are spread in x (vertices inside a layer), so a fine monotone index -- the y-plane index
// resp. 1/1024 of the layer thickness -- cuts the bucket into groups of a handful of elements: ONE counting pass in
// shared memory, then every element finds its rank inside its group by comparing full keys (O(group size) each).
// A group larger than SEG_GROUP_MAX is a face row inside the bucket; if all its keys share (x, y) exactly -- the
// axis-aligned case -- the <SEG_CAP> instance cuts it once more by z-plane index (a second counting pass) and ranks
// inside those pieces.  Anything else (or more than SEG_LARGE_MAX such groups) falls back to the radix passes.
constexpr int SEG_SMALL = 1024;
constexpr int SEG_GROUPS = 1026;     // group ids 0 .. SEG_GROUPS-1
constexpr int SEG_GROUP_MAX = 192;
constexpr int SEG_LARGE_MAX = 8;      // large groups handled per bucket by the nested z stage
template <int CAP, int THREADS>
struct SegCfg {
    static constexpr int WARPS = THREADS / 32;
    static constexpr int CHUNKS = CAP / THREADS;          // 32-element chunks per warp at full capacity
    static constexpr int CNT_WORDS = WARPS * 256 > SEG_GROUPS + 2 ? WARPS * 256 : SEG_GROUPS + 2;   // digit counters / group table
    static constexpr size_t SMEM = (size_t) CAP * 4 * 4       // sk[3] + ids
                                   + (size_t) CAP * 2 * 2     // ord[2] (u16)
                                   + (size_t) CNT_WORDS * 4   // cnt
                                   + (256 + 64) * 4;          // digit bases + vary + scan scratch
};

// second level (k_seg_sort<.., true>): the "candidates" are positions of the big list (keys bk*, ids bid) and the
// sorted run of a bucket goes to its place inside the first-level bucket it belongs to
struct SegLevel2 {
    const u32 *bid, *cbucket, *start1, *bigoff1, *skip, *nb_dev;
};

template <int CAP, int THREADS, int MIN_N, bool LEVEL2 = false>
static __global__ void __launch_bounds__(THREADS, (CAP <= 1024 ? 8 : 2)) k_seg_sort(const u32 *__restrict__ kx, const u32 *__restrict__ ky,
                                                             const u32 *__restrict__ kz, const u32 *__restrict__ count,
                                                             const u32 *__restrict__ start, const u32 *__restrict__ perm0,
                                                             u32 *__restrict__ perm, u32 *__restrict__ skx, u32 *__restrict__ sky,
                                                             u32 *__restrict__ skz, const u32 *__restrict__ n_dev, u32 n_cap, u32 nb,
                                                             SegGeom geom,
                                                             SegLevel2 l2 = SegLevel2{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}) {
    constexpr int WARPS = SegCfg<CAP, THREADS>::WARPS, CHUNKS = SegCfg<CAP, THREADS>::CHUNKS;
    extern __shared__ u32 smem[];
    if (n_dev && *n_dev > n_cap) return;
    if (LEVEL2) {
        if (*l2.skip) return;
        if (*l2.nb_dev < nb) nb = *l2.nb_dev;   // second-level buckets in use
    }
    u32 *sk = smem;                                   // [3][CAP]: z, y, x keys (LSD order)
    u32 *ids = smem + 3 * CAP;                        // [CAP]
    unsigned short *ord = reinterpret_cast<unsigned short *>(smem + 4 * CAP);   // [2][CAP]
    u32 *cnt = smem + 5 * CAP;                        // [WARPS][256] digit counters; group table of the group stage
    u32 *dbase = cnt + SegCfg<CAP, THREADS>::CNT_WORDS;   // [256]
    u32 *vary = dbase + 256;                          // [3] OR of (key ^ key[0]) per coordinate
    u32 *sw = vary + 4;                               // [33] scan scratch
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 b = blockIdx.x; b < nb; b += gridDim.x) {
        const u32 n = count[b];
        if (n < (u32) MIN_N || n > (u32) CAP) continue;
        const u32 s0 = start[b];
        u32 o0 = s0;                                   // where the sorted run goes
        if (LEVEL2) {
            const u32 b1 = l2.cbucket[l2.bid[perm0[s0]]];
            o0 = l2.start1[b1] + (s0 - l2.bigoff1[b1]);
        }
        if (n == 1) {
            if (tid == 0) {
                const u32 id = perm0[s0];
                const u32 src = LEVEL2 ? id : s0;      // first level: keys are stored grouped, next to their ids
                perm[o0] = LEVEL2 ? l2.bid[id] : id;
                skx[o0] = kx[src]; sky[o0] = ky[src]; skz[o0] = kz[src];
            }
            continue;
        }
        __syncthreads();   // shared memory of the previous bucket is free
        if (tid < 3) vary[tid] = 0;
        __syncthreads();
        {
            u32 vz = 0, vy = 0, vx = 0;
            const u32 id0 = LEVEL2 ? perm0[s0] : s0;
            const u32 z0 = kz[id0], y0 = ky[id0], x0 = kx[id0];
            for (u32 i = tid; i < n; i += THREADS) {
                const u32 id = perm0[s0 + i];
                const u32 src = LEVEL2 ? id : s0 + i;
                const u32 z = kz[src], y = ky[src], x = kx[src];
                sk[i] = z; sk[CAP + i] = y; sk[2 * CAP + i] = x;
                ids[i] = id;
                ord[i] = (unsigned short) i;
                vz |= z ^ z0; vy |= y ^ y0; vx |= x ^ x0;
            }
            for (int o = 16; o > 0; o >>= 1) {
                vz |= __shfl_xor_sync(0xffffffffu, vz, o);
                vy |= __shfl_xor_sync(0xffffffffu, vy, o);
                vx |= __shfl_xor_sync(0xffffffffu, vx, o);
            }
            if (lane == 0) { atomicOr(&vary[0], vz); atomicOr(&vary[1], vy); atomicOr(&vary[2], vx); }
        }
        __syncthreads();
        u32 cur = 0;
        bool grouped = false;
        // Group stage: first-level buckets of a grid layer, and second-level pieces whose keys all share x (the cut of an
        // axis-aligned face by y plane: k_big_sub) -- there the y-plane index relative to the piece's lowest y is the index.
        bool piece_const_x = false;
        int j0_piece = 0;
        // nested z stage available?  (needs room for one counter per z plane behind the group table)
        constexpr u32 ZCAP = (u32) SegCfg<CAP, THREADS>::CNT_WORDS > (u32) SEG_GROUPS + 2 + 64
                                 ? (u32) SegCfg<CAP, THREADS>::CNT_WORDS - ((u32) SEG_GROUPS + 2) : 0u;
        const bool nested_ok = ZCAP > 0 && geom.Z + 1 <= ZCAP;
        // A bucket whose keys all share (x, y) -- one row of an axis-aligned face, the typical second-level piece -- is ONE
        // group for the y stage: skip that stage and sort by z plane at once (half the barrier-separated phases).
        const bool xy_const = geom.grouped && nested_ok && vary[2] == 0u && vary[1] == 0u && n > (u32) SEG_GROUP_MAX &&
                              (LEVEL2 || b >= geom.gy + geom.gx);                // block-uniform
        if (LEVEL2 && geom.grouped && vary[2] == 0u && !xy_const) {          // block-uniform
            // lowest / highest y key of the piece -> its first y plane; NaN keys have no plane: radix passes
            if (tid == 0) { dbase[0] = 0xffffffffu; dbase[1] = 0u; }
            __syncthreads();
            u32 mn = 0xffffffffu, mx = 0u;
            for (u32 i = tid; i < n; i += THREADS) { mn = min(mn, sk[CAP + i]); mx = max(mx, sk[CAP + i]); }
            mn = __reduce_min_sync(0xffffffffu, mn);
            mx = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0) { atomicMin(&dbase[0], mn); atomicMax(&dbase[1], mx); }
            __syncthreads();
            mn = dbase[0]; mx = dbase[1];
            __syncthreads();
            piece_const_x = mn >= 0x007fffffu && mx <= 0xff800000u;   // float_key(-inf) .. float_key(+inf)
            if (piece_const_x) j0_piece = (int) yplane_of(key_float(mn), geom) - 1;
        }
        if (xy_const || (geom.grouped && (LEVEL2 ? piece_const_x : b >= geom.gy + geom.gx))) {   // (layer 0 = below the first plane: radix passes)
            const u32 nsub = geom.gy + geom.gx;
            const u32 L = b / nsub, sub = LEVEL2 ? 0u : b - L * nsub;
            const u32 xb = (u32) ((i64) L - 1 + geom.x_off);    // global index of the layer's lower plane (first level only)
            u32 *gtab = cnt;                                    // [SEG_GROUPS + 1]: counts -> starts -> ends
            unsigned short *gid = ord + CAP;                    // group of every element (second half of ord as temp)
            constexpr u32 ngroups = (u32) SEG_GROUPS;
            if (xy_const) {               // the whole bucket is group 0 (ord is still the identity of the load)
                if (tid == 0) { gtab[0] = n; vary[3] = 1; }
                __syncthreads();
            } else {
            for (u32 i = tid; i < (u32) SEG_GROUPS + 2; i += THREADS) gtab[i] = 0;
            if (tid == 0) vary[3] = 0;
            __syncthreads();
            if (sub < geom.gy) {          // all keys share x: groups by y plane
                const int j0 = LEVEL2 ? j0_piece : (int) (sub * geom.ystep) - 1;
                for (u32 i = tid; i < n; i += THREADS) {
                    int g = (int) yplane_of(key_float(sk[CAP + i]), geom) - j0;
                    g = g < 0 ? 0 : (g > SEG_GROUPS - 1 ? SEG_GROUPS - 1 : g);
                    gid[i] = (unsigned short) g;
                    atomicAdd(&gtab[g], 1u);
                }
            } else {                      // keys spread in x inside the layer: groups uniform in x
                const float pb = axis_pos(xb, geom.Xg - 1, geom.amin_x, geom.asize_x);
                const float pb1 = axis_pos(xb + 1, geom.Xg - 1, geom.amin_x, geom.asize_x);
                const float scale = __fdiv_rn((float) (SEG_GROUPS - 1), __fsub_rn(pb1, pb));
                for (u32 i = tid; i < n; i += THREADS) {
                    const float f = __fmul_rn(__fsub_rn(key_float(sk[2 * CAP + i]), pb), scale);
                    u32 g = f > 0.f ? (u32) f : 0u;
                    g = g > (u32) SEG_GROUPS - 1 ? (u32) SEG_GROUPS - 1 : g;
                    gid[i] = (unsigned short) g;
                    atomicAdd(&gtab[g], 1u);
                }
            }
            __syncthreads();
            {   // exclusive scan of the group counts (consecutive groups per thread), largest group
                constexpr int GPT = (SEG_GROUPS + THREADS - 1) / THREADS;
                u32 c[GPT], sum = 0, mx = 0;
#pragma unroll
                for (int q = 0; q < GPT; q++) {
                    const u32 g = tid * GPT + q;
                    c[q] = g < ngroups ? gtab[g] : 0u;
                    sum += c[q];
                    mx = c[q] > mx ? c[q] : mx;
                }
                if (mx > (u32) SEG_GROUP_MAX) vary[3] = 1;      // benign race: every writer stores 1
                u32 total;
                u32 ex = block_exclusive_scan(sum, &total, sw);
#pragma unroll
                for (int q = 0; q < GPT; q++) {
                    const u32 g = tid * GPT + q;
                    if (g < ngroups) gtab[g] = ex;
                    ex += c[q];
                }
            }
            __syncthreads();
            }   // !xy_const
            grouped = vary[3] == 0 || nested_ok;
            if (grouped) {
                u32 *zc = cnt + SEG_GROUPS + 2;                 // [Z + 1] z-plane counters of one large group
                u32 *s_nl = dbase, *s_lg = dbase + 1, *s_bad = dbase + 1 + SEG_LARGE_MAX;   // dbase is free outside the radix passes
                if (tid == 0) { *s_nl = xy_const ? 1u : 0u; *s_bad = 0; s_lg[0] = 0; }
                if (!xy_const)
                    for (u32 i = tid; i < n; i += THREADS) ord[atomicAdd(&gtab[gid[i]], 1u)] = (unsigned short) i;   // gtab: starts -> ends
                __syncthreads();
                // rank of every element inside its group on the full key (ties cannot occur between different elements
                // other than exact duplicates, which are ordered by their slot): O(group size) per element, all threads busy
                // Threads take POSITIONS of the group-ordered list, not elements: neighbouring lanes then rank members of the
                // same (or the next) group -- equal loop bounds instead of the largest group among 32 unrelated elements,
                // and the loads of the loop are broadcasts.
                u32 dst[CHUNKS], del[CHUNKS];
#pragma unroll
                for (int q = 0; q < CHUNKS; q++) {
                    const u32 jp = tid + (u32) q * THREADS;
                    dst[q] = 0xffffffffu;
                    del[q] = 0;
                    if (jp < n && !xy_const) {
                        const u32 i = ord[jp];
                        const u32 g = gid[i];
                        const u32 s1 = g ? gtab[g - 1] : 0u, e1 = gtab[g];
                        if (e1 - s1 > (u32) SEG_GROUP_MAX) continue;   // member of a large group: nested stage below
                        const u32 x = sk[2 * CAP + i], y = sk[CAP + i], z = sk[i];
                        auto before = [&](u32 lj) -> u32 {
                            const u32 xj = sk[2 * CAP + lj], yj = sk[CAP + lj], zj = sk[lj];
                            return (xj < x || (xj == x && (yj < y || (yj == y && (zj < z || (zj == z && lj < i)))))) ? 1u : 0u;
                        };
                        u32 r = 0, j = s1;
                        for (; j + 4 <= e1; j += 4) {
                            const u32 l0 = ord[j], l1 = ord[j + 1], l2 = ord[j + 2], l3 = ord[j + 3];
                            r += before(l0) + before(l1) + before(l2) + before(l3);
                        }
                        for (; j < e1; j++) r += before(ord[j]);
                        dst[q] = s1 + r;
                        del[q] = i;
                    }
                }
                if (vary[3] && !xy_const) {   // list the large groups
                    for (u32 g = tid; g < ngroups; g += THREADS) {
                        const u32 s1 = g ? gtab[g - 1] : 0u, e1 = gtab[g];
                        if (e1 - s1 > (u32) SEG_GROUP_MAX) {
                            const u32 k = atomicAdd(s_nl, 1u);
                            if (k < (u32) SEG_LARGE_MAX) s_lg[k] = g;
                        }
                    }
                }
                __syncthreads();          // everybody has read ord (group order) and gid: the second half of ord is free
                const u32 nl = *s_nl;
                if (nl > (u32) SEG_LARGE_MAX) grouped = false;
                for (u32 t = 0; grouped && t < nl; t++) {
                    const u32 g = s_lg[t];
                    const u32 s1 = g ? gtab[g - 1] : 0u, e1 = gtab[g];
                    for (u32 i = tid; i < geom.Z + 1; i += THREADS) zc[i] = 0;
                    __syncthreads();
                    // (x, y) constant over the group?  z-plane index of every member
                    const u32 l0 = ord[s1];
                    const u32 x0 = sk[2 * CAP + l0], y0 = sk[CAP + l0];
                    u32 lq[CHUNKS], hq[CHUNKS];
#pragma unroll
                    for (int q = 0; q < CHUNKS; q++) {
                        const u32 j = s1 + tid + (u32) q * THREADS;
                        lq[q] = 0xffffffffu;
                        hq[q] = 0;
                        if (j < e1) {
                            const u32 li = ord[j];
                            if (sk[2 * CAP + li] != x0 || sk[CAP + li] != y0) *s_bad = 1;
                            const float zf = key_float(sk[li]);
                            // planes 0..Z-1 -> 1..Z; below the first plane (or NaN) -> 0
                            u32 h = plane_index(zf, geom.amin_z, geom.asize_z, geom.Z - 1) + 1;
                            if (!(axis_pos(0u, geom.Z - 1, geom.amin_z, geom.asize_z) <= zf)) h = 0;
                            lq[q] = li;
                            hq[q] = h;
                            atomicAdd(&zc[h], 1u);
                        }
                    }
                    __syncthreads();
                    if (*s_bad) { grouped = false; break; }
                    {   // exclusive scan of the z-plane counters
                        constexpr int ZPT = (int) ((ZCAP + THREADS - 1) / THREADS) > 0 ? (int) ((ZCAP + THREADS - 1) / THREADS) : 1;
                        u32 c[ZPT], sum = 0;
#pragma unroll
                        for (int q = 0; q < ZPT; q++) {
                            const u32 k = tid * ZPT + q;
                            c[q] = k < geom.Z + 1 ? zc[k] : 0u;
                            sum += c[q];
                        }
                        u32 total;
                        u32 ex = block_exclusive_scan(sum, &total, sw);
#pragma unroll
                        for (int q = 0; q < ZPT; q++) {
                            const u32 k = tid * ZPT + q;
                            if (k < geom.Z + 1) zc[k] = ex;
                            ex += c[q];
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int q = 0; q < CHUNKS; q++)
                        if (lq[q] != 0xffffffffu) {
                            const u32 pos = s1 + atomicAdd(&zc[hq[q]], 1u);      // zc: starts -> ends
                            ord[pos] = (unsigned short) lq[q];
                            gid[pos] = (unsigned short) hq[q];
                        }
                    __syncthreads();
                    u32 pq[CHUNKS];
#pragma unroll
                    for (int q = 0; q < CHUNKS; q++) {
                        const u32 j = s1 + tid + (u32) q * THREADS;
                        pq[q] = 0xffffffffu;
                        if (j < e1) {
                            const u32 li = ord[j], h = gid[j];
                            const u32 a0 = s1 + (h ? zc[h - 1] : 0u), a1 = s1 + zc[h];
                            const u32 z = sk[li];
                            u32 r = 0;
                            for (u32 k = a0; k < a1; k++) {
                                const u32 lk = ord[k];
                                const u32 zk = sk[lk];
                                r += (zk < z || (zk == z && lk < li)) ? 1u : 0u;     // x and y are equal throughout
                            }
                            pq[q] = a0 + r;
                            lq[q] = li;
                        }
                    }
                    __syncthreads();
#pragma unroll
                    for (int q = 0; q < CHUNKS; q++)
                        if (pq[q] != 0xffffffffu) gid[pq[q]] = (unsigned short) lq[q];   // result, second half of ord
                    __syncthreads();
                }
                if (grouped) {
#pragma unroll
                    for (int q = 0; q < CHUNKS; q++)
                        if (dst[q] != 0xffffffffu) gid[dst[q]] = (unsigned short) del[q];   // second half of ord = result
                    cur = 1;
                }
            }
            __syncthreads();
        }
        // contiguous run of positions per warp, a multiple of 32
        const u32 run = ((n + WARPS - 1) / WARPS + 31) & ~31u;
        const u32 nchunks = run >> 5;                     // <= CHUNKS
        const u32 wbase = warp * run;
        for (int pass = 0; pass < (grouped ? 0 : 12); pass++) {
            const u32 c = pass >> 2, shift = 8 * (pass & 3);
            if (((vary[c] >> shift) & 255u) == 0) continue;   // all keys agree on this digit
            const u32 *key = sk + c * CAP;
            const unsigned short *oin = ord + cur * CAP;
            unsigned short *oout = ord + (cur ^ 1) * CAP;
            for (u32 i = tid; i < (u32) WARPS * 256; i += THREADS) cnt[i] = 0;
            __syncthreads();
            u32 packed[CHUNKS];                           // local index | digit << 16
            u32 rnk[CHUNKS];
#pragma unroll
            for (int j = 0; j < CHUNKS; j++) {
                packed[j] = 0xffffffffu;
                rnk[j] = 0;
                if ((u32) j < nchunks) {
                    const u32 i = wbase + j * 32 + lane;
                    const bool valid = i < n;
                    const u32 active = __ballot_sync(0xffffffffu, valid);
                    if (valid) {
                        const u32 li = oin[i];
                        const u32 d = (key[li] >> shift) & 255u;
                        const u32 peers = __match_any_sync(active, d);
                        const u32 leader = __ffs(peers) - 1;
                        u32 old = 0;
                        if (lane == leader) {
                            old = cnt[warp * 256 + d];
                            cnt[warp * 256 + d] = old + __popc(peers);
                        }
                        old = __shfl_sync(peers, old, leader);
                        rnk[j] = old + __popc(peers & ((1u << lane) - 1u));
                        packed[j] = li | (d << 16);
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            // digit totals over the warps (exclusive per warp), then exclusive scan over the 256 digits
            u32 ex;
            if (THREADS >= 256) {
                u32 tot = 0;
                if (tid < 256) {
#pragma unroll
                    for (int w = 0; w < WARPS; w++) {
                        const u32 t = cnt[w * 256 + tid];
                        cnt[w * 256 + tid] = tot;
                        tot += t;
                    }
                }
                u32 total;
                ex = block_exclusive_scan(tot, &total, sw);
                if (tid < 256) dbase[tid] = ex;
            } else {
                // fewer threads than digits: each thread owns 256/THREADS consecutive digits
                constexpr int DPT = 256 / THREADS > 0 ? 256 / THREADS : 1;
                u32 tots[DPT], sum = 0;
#pragma unroll
                for (int q = 0; q < DPT; q++) {
                    const u32 d = tid * DPT + q;
                    u32 tot = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) {
                        const u32 t = cnt[w * 256 + d];
                        cnt[w * 256 + d] = tot;
                        tot += t;
                    }
                    tots[q] = tot;
                    sum += tot;
                }
                u32 total;
                ex = block_exclusive_scan(sum, &total, sw);
#pragma unroll
                for (int q = 0; q < DPT; q++) {
                    dbase[tid * DPT + q] = ex;
                    ex += tots[q];
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < CHUNKS; j++) {
                if (packed[j] != 0xffffffffu) {
                    const u32 d = (packed[j] >> 16) & 255u;
                    oout[dbase[d] + cnt[warp * 256 + d] + rnk[j]] = (unsigned short) (packed[j] & 0xffffu);
                }
            }
            __syncthreads();
            cur ^= 1;
        }
        const unsigned short *ofin = ord + cur * CAP;
        for (u32 i = tid; i < n; i += THREADS) {
            const u32 li = ofin[i];
            perm[o0 + i] = LEVEL2 ? l2.bid[ids[li]] : ids[li];
            skz[o0 + i] = sk[li]; sky[o0 + i] = sk[CAP + li]; skx[o0 + i] = sk[2 * CAP + li];
        }
    }
}

// ---- oversized buckets ----------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 pow2ceil_u32(u32 v) {
    u32 p = 1;
    while (p < v) p <<= 1;
    return p;
}

// One block: decide how every oversized bucket is cut and where its second-level buckets start.
// An oversized bucket is almost always an axis-aligned face: nearly all its vertices share ONE x value xm (found
// by majority vote over 32 samples), the rest is rounding noise and bystanders of the same layer.  Cut:
//     x <  xm : w pieces, uniform in the key range [xmin, xm)
//     x == xm : gy2 pieces by y-plane index (x is constant here, so y decides)
//     x >  xm : w pieces, uniform in (xm, xmax]
// with w = count / 1024 rounded up to a power of two.  Monotone in the lexicographic key order, so sorting the
// pieces sorts the bucket; a slanted sheet (no dominant x) is spread by the two uniform ranges.
static __global__ void __launch_bounds__(1024) k_big_plan(u32 nb, const u32 *__restrict__ count, const u32 *__restrict__ bigoff,
                                                          const u32 *__restrict__ bkx, u32 *__restrict__ sub_base,
                                                          u32 *__restrict__ sub_par, u32 *__restrict__ sub_xm, u32 Y, u32 subcap,
                                                          u32 *__restrict__ info2, const u32 *__restrict__ n_dev, u32 n_cap) {
    __shared__ u32 sw[33];
    __shared__ u32 s_run;
    const bool abort = n_dev && *n_dev > n_cap;   // the candidate buffers were too small: nothing valid to sort
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (u32 base = 0; base < nb; base += blockDim.x) {
        const u32 b = base + threadIdx.x;
        u32 nsub = 0, w = 0, xm = 0;
        if (b < nb && !abort) {
            const u32 c = __ldcg(count + b);
            if (c > (u32) SEG_CAP) {
                w = pow2ceil_u32((c + 1023u) / 1024u);
                if (w > 4096u) w = 4096u;
                const u32 o = bigoff[b], step = c / 32u;
                u32 xs[32], best = 0;
#pragma unroll
                for (int i = 0; i < 32; i++) xs[i] = bkx[o + (u32) i * step];
#pragma unroll
                for (int i = 0; i < 32; i++) {        // majority vote over 32 samples
                    u32 votes = 0;
#pragma unroll
                    for (int k = 0; k < 32; k++) votes += xs[k] == xs[i] ? 1u : 0u;
                    if (votes > best) { best = votes; xm = xs[i]; }
                }
                const u32 gy2 = w > pow2ceil_u32(Y) ? pow2ceil_u32(Y) : w;
                nsub = 2 * w + gy2;
            }
        }
        u32 tot;
        const u32 ex = block_exclusive_scan(nsub, &tot, sw);
        if (b < nb) {
            sub_base[b] = s_run + ex;
            sub_par[b] = w;
            sub_xm[b] = xm;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_run += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        info2[0] = s_run;
        info2[3] = s_run > subcap ? 1u : 0u;          // cannot be represented: the radix sort takes over
        info2[5] = (abort || s_run > subcap) ? 1u : 0u;   // no second level
        info2[6] = abort ? 1u : 0u;
    }
}

static __global__ void __launch_bounds__(256) k_big_sub(u32 n_big, const u32 *__restrict__ nbig_dev, const u32 *__restrict__ bkx,
                                                        const u32 *__restrict__ bky, const u32 *__restrict__ bid,
                                                        const u32 *__restrict__ cbucket, const u32 *__restrict__ xinvmin,
                                                        const u32 *__restrict__ xmax, const u32 *__restrict__ sub_base,
                                                        const u32 *__restrict__ sub_par, const u32 *__restrict__ sub_xm, SegGeom g,
                                                        u32 *__restrict__ csub, u32 *__restrict__ count2, const u32 *__restrict__ info2) {
    if (nbig_dev) n_big = *nbig_dev;
    if (info2[5]) return;
    const u32 lane = threadIdx.x & 31;
    const u32 ypow = pow2ceil_u32(g.Y);
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_big; base += gridDim.x * blockDim.x) {
        const u32 j = base + lane;
        const bool valid = j < n_big;
        const u32 active = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const u32 b = cbucket[bid[j]];
            const u32 w = sub_par[b], xm = sub_xm[b], x = bkx[j];
            const u32 gy2 = w > ypow ? ypow : w;
            u32 sub;
            if (x < xm) {
                const u32 xmin = ~xinvmin[b];
                sub = (u32) (((u64) (x - xmin) * w) / (u64) (xm - xmin));
            } else if (x == xm) {
                sub = w + (u32) (((u64) yplane_of(key_float(bky[j]), g) * gy2) / g.Y);
            } else {
                sub = w + gy2 + (u32) (((u64) (x - xm - 1u) * w) / (u64) (xmax[b] - xm));
            }
            const u32 cs = sub_base[b] + sub;
            csub[j] = cs;
            const u32 peers = __match_any_sync(active, cs);
            if (lane == (u32) (__ffs(peers) - 1)) atomicAdd(&count2[cs], (u32) __popc(peers));
        }
    }
}

// after the second-level scan: does the radix sort have to take over?  info2[4] = number of elements it sorts
static __global__ void k_big_decide(u32 n_big, const u32 *__restrict__ nbig_dev, u32 *__restrict__ info2, u32 *__restrict__ radix_needed) {
    if (nbig_dev) n_big = *nbig_dev;
    info2[4] = info2[6] ? 0u : ((info2[3] || info2[2] > (u32) SEG_CAP) ? n_big : 0u);
    if (radix_needed) *radix_needed = info2[4];
}
// sorted rank r of the big list -> final position: the big list is ordered by bucket (x decides the bucket)
static __global__ void __launch_bounds__(256) k_seg_big_scatter(u32 n_big, const u32 *__restrict__ nbig_dev, const u32 *__restrict__ sorted,
                                                                const u32 *__restrict__ bid,
                                                                const u32 *__restrict__ cbucket, const u32 *__restrict__ start,
                                                                const u32 *__restrict__ bigoff, u32 *__restrict__ perm,
                                                                const u32 *__restrict__ bkx, const u32 *__restrict__ bky,
                                                                const u32 *__restrict__ bkz, u32 *__restrict__ skx, u32 *__restrict__ sky,
                                                                u32 *__restrict__ skz) {
    if (nbig_dev) n_big = *nbig_dev;
    for (u32 r = blockIdx.x * blockDim.x + threadIdx.x; r < n_big; r += gridDim.x * blockDim.x) {
        const u32 j = sorted[r];
        const u32 id = bid[j];
        const u32 b = cbucket[id];
        const u32 pos = start[b] + (r - bigoff[b]);
        perm[pos] = id;
        skx[pos] = bkx[j]; sky[pos] = bky[j]; skz[pos] = bkz[j];
    }
}

// Where the candidates go: straight to their place in the bucket-grouped arrays of the segmented sort (the bucket
// offsets were scanned in phase 1), so no separate scatter pass and the sort reads its keys contiguously.  Elements of
// oversized buckets go to the compacted big list instead and feed the x range of their bucket (segsort.cuh).
struct CandOut {
    u32 *gkx, *gky, *gkz, *gid;        // grouped by bucket: keys + candidate id
    u32 *cbucket;                      // bucket of every candidate (candidate order)
    const u32 *count, *start, *bigoff;
    u32 *cursor;
    u32 *bkx, *bky, *bkz, *bid, *xinvmin, *xmax;
};
// executed by all 32 lanes; `has` = this lane emits a candidate
__device__ __forceinline__ void emit_candidate(bool has, u32 b, u32 id, u32 kxv, u32 kyv, u32 kzv, const CandOut &o) {
    const u32 lane = threadIdx.x & 31;
    const u32 act = __ballot_sync(0xffffffffu, has);
    if (!has) return;
    const u32 peers = __match_any_sync(act, b);
    const u32 leader = __ffs(peers) - 1;
    u32 off = 0;
    if (lane == leader) off = atomicAdd(&o.cursor[b], (u32) __popc(peers));
    off = __shfl_sync(peers, off, leader) + __popc(peers & ((1u << lane) - 1u));
    o.cbucket[id] = b;
    if (o.count[b] > (u32) SEG_CAP) {
        const u32 q = o.bigoff[b] + off;
        o.bkx[q] = kxv; o.bky[q] = kyv; o.bkz[q] = kzv; o.bid[q] = id;
        const u32 mx = __reduce_max_sync(peers, kxv), mn = __reduce_max_sync(peers, ~kxv);
        if (lane == leader) {
            atomicMax(&o.xmax[b], mx);
            atomicMax(&o.xinvmin[b], mn);
        }
    } else {
        const u32 q = o.start[b] + off;
        o.gkx[q] = kxv; o.gky[q] = kyv; o.gkz[q] = kzv; o.gid[q] = id;
    }
}


// Phase-2 part of the segmented sort: group by bucket, sort the small buckets in shared memory, and run the
// radix fallback for the oversized buckets.  The histogram (h.count) and its scan (h.start / h.bigoff) were
// produced in phase 1.
//   two-phase path : n, n_big known on the host (n_dev = nbig_dev = nullptr); the fallback runs iff n_big > 0
//   single-sync path: counts live on the device; n_cap sizes the launches; the second level is enqueued iff
//                    big_cap > 0 (the caller's guess) and does nothing when the device count is 0
static inline cudaError_t seg_sort_run(const u32 *kx, const u32 *ky, const u32 *kz, u32 n, const u32 *n_dev, u32 n_cap, u32 grid_n,
                                       u32 nb, u32 n_big, const u32 *nbig_dev, u32 big_cap, const SegHead &h, const SegScratch &b,
                                       const SegGeom &geom, bool allow_radix, u32 *radix_needed, cudaStream_t stream) {
    constexpr size_t smem_big = SegCfg<SEG_CAP, SEG_THREADS>::SMEM, smem_small = SegCfg<SEG_SMALL, 128>::SMEM;
    // per device: the attribute belongs to the device's context; per host thread: the auxiliary stream / event pair
    // below must not be shared by concurrent callers (thread_local, so no lock is needed)
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    static thread_local bool attr_set_of[64];
    bool &attr_set = attr_set_of[dev];
    if (!attr_set) {
        cudaFuncSetAttribute(k_seg_sort<SEG_CAP, SEG_THREADS, SEG_SMALL + 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_big);
        cudaFuncSetAttribute(k_seg_sort<SEG_SMALL, 128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_small);
        cudaFuncSetAttribute(k_seg_sort<SEG_CAP, SEG_THREADS, SEG_SMALL + 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_big);
        cudaFuncSetAttribute(k_seg_sort<SEG_SMALL, 128, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_small);
        attr_set = true;
    }
    const int blocks = (int) ((grid_n + 255) / 256 > 148 * 8 ? 148 * 8 : (grid_n + 255) / 256);
    // first level: the producer has already placed keys (kx/ky/kz) and ids (b.perm0) grouped by bucket, and the members
    // of oversized buckets in the compacted big list (mc_dense.cu emit_candidate)
    (void) blocks;
    // The two instances work on disjoint buckets and each is dominated by its slowest bucket (ncu: 13 % / 32 % SM
    // throughput), so the small-bucket instance runs on an auxiliary stream next to the large one and the second level.
    static thread_local cudaStream_t aux_of[64];   // one auxiliary stream + event pair per device and host thread, created on first use
    static thread_local cudaEvent_t fork_of[64], join_of[64];
    if (!aux_of[dev]) {
        cudaError_t e0 = cudaStreamCreateWithFlags(&aux_of[dev], cudaStreamNonBlocking);
        if (e0 == cudaSuccess) e0 = cudaEventCreateWithFlags(&fork_of[dev], cudaEventDisableTiming);
        if (e0 == cudaSuccess) e0 = cudaEventCreateWithFlags(&join_of[dev], cudaEventDisableTiming);
        if (e0 != cudaSuccess) { aux_of[dev] = nullptr; return e0; }
    }
    const cudaStream_t aux = aux_of[dev];
    const cudaEvent_t ev_fork = fork_of[dev], ev_join = join_of[dev];
    cudaError_t ef = cudaEventRecord(ev_fork, stream);
    if (ef == cudaSuccess) ef = cudaStreamWaitEvent(aux, ev_fork, 0);
    if (ef != cudaSuccess) return ef;
    const u32 grid_small = nb < 148u * 32u ? nb : 148u * 32u, grid_large = nb < 148u * 8u ? nb : 148u * 8u;
    ISX_LAUNCH((k_seg_sort<SEG_SMALL, 128, 1>), grid_small, 128, smem_small, aux, kx, ky, kz, h.count, h.start, b.perm0, b.perm,
               b.skx, b.sky, b.skz, n_dev, n_cap, nb, geom);
    if ((ef = cudaEventRecord(ev_join, aux)) != cudaSuccess) return ef;
    ISX_LAUNCH((k_seg_sort<SEG_CAP, SEG_THREADS, SEG_SMALL + 1>), grid_large, SEG_THREADS, smem_big, stream, kx, ky, kz, h.count,
               h.start, b.perm0, b.perm, b.skx, b.sky, b.skz, n_dev, n_cap, nb, geom);
    const u32 fb_n = nbig_dev ? big_cap : n_big;
    if (fb_n > 0) {
        // second level inside the oversized buckets
        const u32 nb2 = b.subcap;
        u32 *count2 = b.count2, *start2 = count2 + (nb2 + 1), *cursor2 = start2 + (nb2 + 1), *bigoff2 = cursor2 + (nb2 + 1);
        cudaError_t e = cudaMemsetAsync(b.info2, 0, (8 + (size_t) nb2 + 1) * sizeof(u32), stream);
        if (e != cudaSuccess) return e;
        ISX_LAUNCH(k_big_plan, 1, 1024, 0, stream, nb, h.count, h.bigoff, b.bkx, h.sub_base, h.sub_par, h.sub_xm, geom.Y, nb2, b.info2,
                   n_dev, n_cap);
        const int bb = (int) ((fb_n + 255) / 256 > 148 * 8 ? 148 * 8 : (fb_n + 255) / 256);
        ISX_LAUNCH(k_big_sub, bb, 256, 0, stream, fb_n, nbig_dev, b.bkx, b.bky, b.bid, b.cbucket, h.xinvmin, h.xmax, h.sub_base, h.sub_par,
                   h.sub_xm, geom, b.csub, count2, b.info2);
        ISX_LAUNCH(k_seg_scan, 1, 1024, 0, stream, nb2, count2, start2, cursor2, bigoff2, b.info2 + 1, b.info2 + 2, b.info2);
        ISX_LAUNCH(k_big_decide, 1, 1, 0, stream, fb_n, nbig_dev, b.info2, radix_needed);
        ISX_LAUNCH(k_seg_scatter<false>, bb, 256, 0, stream, b.csub, fb_n, start2, cursor2, b.perm2, nbig_dev, 0xffffffffu, b.info2 + 5,
                   SegBigOut{});
        const SegLevel2 l2{b.bid, b.cbucket, h.start, h.bigoff, b.info2 + 5, b.info2};
        const u32 g2s = nb2 < 148u * 32u ? nb2 : 148u * 32u, g2l = nb2 < 148u * 8u ? nb2 : 148u * 8u;
        ISX_LAUNCH((k_seg_sort<SEG_SMALL, 128, 1, true>), g2s, 128, smem_small, stream, b.bkx, b.bky, b.bkz, count2, start2, b.perm2, b.perm,
                   b.skx, b.sky, b.skz, nullptr, 0u, nb2, geom, l2);
        ISX_LAUNCH((k_seg_sort<SEG_CAP, SEG_THREADS, SEG_SMALL + 1, true>), g2l, SEG_THREADS, smem_big, stream, b.bkx, b.bky, b.bkz, count2,
                   start2, b.perm2, b.perm, b.skx, b.sky, b.skz, nullptr, 0u, nb2, geom, l2);
        // last resort: a second-level bucket is still oversized (info2[4] != 0): global radix sort of the big list.
        // 14 launches that do nothing in the common case, so the single-sync path only enqueues them when the
        // previous extraction of the grid needed them (*radix_needed tells the host)
        if (!allow_radix) {
            if ((ef = cudaStreamWaitEvent(stream, ev_join, 0)) != cudaSuccess) return ef;
            return cudaGetLastError();
        }
        e = radix_sort96(b.bkx, b.bky, b.bkz, fb_n, b.radix, stream, b.info2 + 4);
        if (e != cudaSuccess) return e;
        ISX_LAUNCH(k_seg_big_scatter, bb, 256, 0, stream, fb_n, b.info2 + 4, b.radix.perm[0], b.bid, b.cbucket, h.start, h.bigoff, b.perm, b.bkx,
                   b.bky, b.bkz, b.skx, b.sky, b.skz);
    }
    if ((ef = cudaStreamWaitEvent(stream, ev_join, 0)) != cudaSuccess) return ef;
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------
// Generic front end: segmented sort of ANY set of 96-bit position keys that lie on / inside the grid (dual
// vertices of dual contouring, the per-(cell, edge) vertices of sparse marching cubes).  The marching-cubes path
// derives the bucket of a vertex from the edge that owns it while it scans its entries; here the bucket comes from
// the key itself:
//     layer  L = (index of the last x plane at or below x) - x_off + 1      (exact float compares, monotone in x)
//     sub      = y-plane index / ystep           if x == px[L-1] exactly     (gy pieces, monotone in y)
//              = gy + floor((x - pb) * gx / (pb1 - pb))  otherwise           (gx pieces, monotone in x)
// which is what the group stage of k_seg_sort assumes about a first-level bucket.  Three extra passes over the
// candidates (bucket ids + histogram, one-block scan, grouping) replace the 14-launch global radix sort.
// NaN keys have no place in this order: they raise *radix_needed and the caller falls back to radix_sort96.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_gen_bucket(const u32 *__restrict__ kx, const u32 *__restrict__ ky, u32 n, SegGeom g,
                                                           u32 x_planes /* local point planes */, u32 *__restrict__ cbucket,
                                                           u32 *__restrict__ count, u32 *__restrict__ nan_count) {
    const u32 lane = threadIdx.x & 31, nsub = g.gy + g.gx;
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {
        const u32 i = base + lane;
        const bool valid = i < n;
        const u32 active = __ballot_sync(0xffffffffu, valid);
        if (!valid) continue;
        const float x = key_float(kx[i]), y = key_float(ky[i]);
        u32 b;
        if (x != x || y != y) {
            atomicAdd(nan_count, 1u);
            b = 0;
        } else {
            const u32 resx = g.Xg - 1;
            const i64 j = (i64) plane_index(x, g.amin_x, g.asize_x, resx);        // global plane index (0 if below all planes)
            const bool below = !(axis_pos(0u, resx, g.amin_x, g.asize_x) <= x);
            i64 L = below ? 0 : j - g.x_off + 1;
            L = L < 0 ? 0 : (L > (i64) x_planes ? (i64) x_planes : L);
            u32 sub = 0;
            if (L >= 1) {
                const u32 xb = (u32) (L - 1 + g.x_off);
                const float pb = axis_pos(xb, resx, g.amin_x, g.asize_x);
                if (x == pb) {
                    sub = yplane_of(y, g) / g.ystep;
                    sub = sub > g.gy - 1 ? g.gy - 1 : sub;
                } else {
                    u32 k = 0;
                    if (g.gx > 1 && xb < resx) {
                        const float pb1 = axis_pos(xb + 1, resx, g.amin_x, g.asize_x);
                        const float f = __fmul_rn(__fsub_rn(x, pb), __fdiv_rn((float) g.gx, __fsub_rn(pb1, pb)));
                        k = f > 0.f ? (u32) f : 0u;
                        k = k > g.gx - 1 ? g.gx - 1 : k;
                    }
                    sub = g.gy + k;
                }
            }
            b = (u32) L * nsub + sub;
        }
        cbucket[i] = b;
        const u32 peers = __match_any_sync(active, b);
        if (lane == (u32) (__ffs(peers) - 1)) atomicAdd(&count[b], (u32) __popc(peers));
    }
}

static __global__ void __launch_bounds__(256) k_gen_group(const u32 *__restrict__ kx, const u32 *__restrict__ ky, const u32 *__restrict__ kz,
                                                          u32 n, const u32 *__restrict__ cbucket_in, CandOut out) {
    const u32 lane = threadIdx.x & 31;
    for (u32 base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += gridDim.x * blockDim.x) {   // warp-uniform
        const u32 i = base + lane;
        const bool has = i < n;
        emit_candidate(has, has ? cbucket_in[i] : 0u, i, has ? kx[i] : 0u, has ? ky[i] : 0u, has ? kz[i] : 0u, out);
    }
}

// scratch of the generic sort for n candidates on a grid with x_planes local point planes
struct GenSort {
    SegHead head;
    SegScratch seg;
    u32 *gkx, *gky, *gkz;    // keys grouped by bucket
    u32 *cb;                 // bucket of every candidate (k_gen_bucket -> k_gen_group; emit_candidate rewrites seg.cbucket)
    u32 nb;
    static u32 pick_groups(u32 n, u32 x_planes, u32 Y) {
        // pieces per layer so that a bucket holds ~1-2 k elements, and few enough y planes per piece for the group stage
        if (g_tuning[3] > 0 && g_tuning[3] < 100) return (u32) g_tuning[3];
        u32 g = 1;
        const u64 per_layer = (u64) n / (x_planes ? x_planes : 1u);
        while (g < 16 && (u64) g * 1536u < per_layer) g *= 2;
        while (g < 32 && (Y + g - 1) / g + 2 > (u32) SEG_GROUPS) g *= 2;
        return g;
    }
    static size_t carve(Carver &c, size_t n, u32 x_planes, u32 Y, GenSort *out) {
        GenSort s;
        const u32 g = pick_groups((u32) n, x_planes, Y);
        s.nb = (x_planes + 2) * (2 * g);
        SegHead::carve(c, s.nb, &s.head);
        SegScratch::carve(c, n, &s.seg);
        s.gkx = c.take<u32>(n);
        s.gky = c.take<u32>(n);
        s.gkz = c.take<u32>(n);
        s.cb = c.take<u32>(n);
        if (out) *out = s;
        return c.bytes();
    }
};

// Sorts n candidates (keys kx/ky/kz in candidate order).  Result: s.seg.perm (candidate ids in sorted order) and
// s.seg.skx/sky/skz (their keys), ready for k_unique(..., keys_sorted = true).  counters: the caller's counter block
// (C_NBIG, C_MAXB, C_RADIX, C_ABORT are used); counters[C_ABORT] or counters[C_RADIX] != 0 afterwards means the result is incomplete (NaN keys,
// or a piece that the two bucket levels could not cut below SEG_CAP) and the caller must use radix_sort96 instead.
static inline cudaError_t gen_sort_run(const u32 *kx, const u32 *ky, const u32 *kz, u32 n, const Geom &gm, const GenSort &s,
                                       u32 *counters, cudaStream_t stream) {
    const u32 X = (u32) gm.X, Y = (u32) gm.Y;
    const u32 g = GenSort::pick_groups(n, X, Y);
    SegGeom geom{gm.amin[0], gm.asize[0], gm.amin[1], gm.asize[1], gm.amin[2], gm.asize[2], (u32) gm.Xg, Y, (u32) gm.Z, gm.x_off, g, g,
                 (Y + g - 1) / g, false};
    geom.grouped = geom.ystep + 2 <= (u32) SEG_GROUPS;
    cudaError_t e = cudaMemsetAsync(s.head.count, 0, SegHead::words(s.nb) * sizeof(u32), stream);
    if (e != cudaSuccess) return e;
    const int blocks = (int) ((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    ISX_LAUNCH(k_gen_bucket, blocks, 256, 0, stream, kx, ky, n, geom, X, s.cb, s.head.count, counters + C_ABORT);
    ISX_LAUNCH(k_seg_scan, 1, 1024, 0, stream, s.nb, s.head.count, s.head.start, s.head.cursor, s.head.bigoff, counters + C_NBIG,
               counters + C_MAXB);
    const CandOut co{s.gkx, s.gky, s.gkz, s.seg.perm0, s.seg.cbucket, s.head.count, s.head.start, s.head.bigoff, s.head.cursor,
                     s.seg.bkx, s.seg.bky, s.seg.bkz, s.seg.bid, s.head.xinvmin, s.head.xmax};
    ISX_LAUNCH(k_gen_group, blocks, 256, 0, stream, kx, ky, kz, n, s.cb, co);
    e = seg_sort_run(s.gkx, s.gky, s.gkz, n, nullptr, n, n, s.nb, 0, counters + C_NBIG, n, s.head, s.seg, geom, false,
                     counters + C_RADIX, stream);
    return e;
}
// gate of the weld pass behind gen_sort_run: no capacities to check, but NaN keys (C_ABORT) or a still-oversized
// second-level bucket (C_RADIX) mean the permutation is incomplete
static inline Gate gen_sort_gate() {
    Gate g;
    g.allow_radix = 0;
    g.on = 1;
    return g;
}
// The caller's second attempt after counters[C_ABORT] / counters[C_RADIX] != 0: clears what the first attempt left in the counter block.
static inline cudaError_t gen_sort_reset_for_radix(u32 *counters, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(counters + C_TICKET_C, 0, 4 * sizeof(u32), stream);   // C_TICKET_C, C_V, C_NLO, C_NHI
    if (e == cudaSuccess) e = cudaMemsetAsync(counters + C_RADIX, 0, 2 * sizeof(u32), stream);   // C_RADIX, C_ABORT
    return e;
}

}   // namespace isx
