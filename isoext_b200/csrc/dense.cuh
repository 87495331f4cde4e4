// dense.cuh -- shared front end of the dense-grid pipelines (marching cubes, intersections, DC).
//
// Replaces, without materialising points (12 B/pt) or cells (32 B/cell), the reference's
//   get_values/get_points/get_cells + get_case_num_op + remove_if compaction
//   (src/mc/mc.cu:19-42, src/its.cu:94-126, include/utils.cuh:32-102 of the reference).
//
// Data layout in HBM
//   values    (X,Y,Z) f32, z fastest (index = x*Y*Z + y*Z + z)          caller-owned, read ONCE by k_signbits
//   bits      1 bit per point, flat (bit n <-> point n): [values[n] - level < 0]      P/8 bytes
//   entries   ordered list of "active points": a point is active if the cell whose corner 0 it is
//             crosses the surface, or if it owns (+z,+y,+x edge from it) a sign-change edge.
//             uint2 { row = x*Y + y,  z | case<<16 | own<<24 | cellvalid<<27 }             8 B each
//   row_start first entry of every point row (x,y), X*Y+1 u32: the only neighbour-lookup structure
//             (an edge shared by 4 cells is owned by exactly one entry; no hash pass, no dense map)
#pragma once
#include "common.cuh"
#include "sdfprog.cuh"

namespace isx {

// counters (u32) shared by the pipelines
enum Counter : int {
    C_TICKET_A = 0,   // k_compact tiles
    C_S = 1,          // number of entries
    C_TICKET_B = 2,   // entry scan tiles
    C_T = 3,          // triangles
    C_VC = 4,         // vertex candidates (edge-unique, before positional welding)
    C_TICKET_C = 5,   // unique scan tiles
    C_V = 6,          // welded vertices
    C_NLO = 7,        // welded vertices with x <  slab lower threshold
    C_NHI = 8,        // welded vertices with x <  slab upper threshold
    C_I = 9,          // intersections (DC)
    C_Q = 10,         // quads (DC)
    C_TICKET_D = 11,
    C_TICKET_E = 12,
    C_NBIG = 13,      // candidates in x-buckets larger than SEG_CAP (radix fallback needed if > 0)
    C_MAXB = 14,      // largest x-bucket
    C_NHEAVY = 15,    // rows handed to k_rowfill_heavy
    C_RADIX = 16,     // candidates that need the radix last resort of the segmented sort (0 = none)
    C_ABORT = 17,     // single-call path: the sort was not completed (a capacity or the radix opt-in was missed)
    C_COUNT = 20
};

// Single-sync paths: the kernels behind the sort decide for themselves whether its output is complete (they do nothing
// otherwise and the host falls back): a capacity was exceeded, there are more candidates in oversized buckets than the
// second bucket level was sized for, or the radix last resort was needed but not enqueued.
struct Gate {
    u32 entry_cap = 0xffffffffu, cand_cap = 0xffffffffu, tri_cap = 0xffffffffu, big_cap = 0xffffffffu;
    int allow_radix = 1;
    int on = 0;
};
__device__ __forceinline__ bool gate_bad(const u32 *__restrict__ counters, const Gate &g) {
    if (counters[C_ABORT]) return true;
    if (!g.on) return false;
    return counters[C_S] > g.entry_cap || counters[C_VC] > g.cand_cap || counters[C_T] > g.tri_cap || counters[C_NBIG] > g.big_cap ||
           (counters[C_RADIX] > 0u && !g.allow_radix);
}

struct DenseParams {
    Geom g;
    i64 P;       // X*Y*Z
    i64 YZ;      // Y*Z
    u32 CPR;     // 32-point chunks per row
    u32 NQ;      // total chunks = X*Y*CPR
    u32 R;       // rows = X*Y
    float level;
    u32 emit_lo, emit_hi;   // local cell-x range [lo,hi) whose faces / quads are emitted
    // sort buckets (segsort.cuh): every x-layer [px[b], px[b+1]) is cut into gy groups of vertices lying exactly
    // on plane b (split by y) and gx groups of vertices inside the layer (split by x)
    u32 gy, gx, ystep;
    // k_cell_tris fast path: an edge whose crossing parameter t lies in [eps2[a], 1 - eps2[a]] (a = axis of the edge:
    // 0 x, 1 y, 2 z) is provably more than an ulp away from both end points, see cell_is_plain()
    float eps2[3];
    // implicit field (sdfprog.cuh): when set, no `values` array exists and every kernel that would load a value
    // evaluates the program at the point's position instead
    const SdfProg *sdf;
};

// value of the field at local point (x, y, z): a load, or an evaluation of the analytic program at the position
// get_points() reports for that point
// IMPLICIT is a template parameter of the kernels (not a run-time branch), so the explicit-field instances carry no
// trace of the interpreter (no call, no stack frame, no spills).
template <bool IMPLICIT>
__device__ __forceinline__ float field_value(const float *__restrict__ values, const DenseParams &p, u32 x, u32 y, u32 z) {
    if (IMPLICIT)
        return sdf_eval(p.sdf, axis_pos(x + (u32) p.g.x_off, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]),
                        axis_pos(y, (u32) p.g.Y - 1, p.g.amin[1], p.g.asize[1]), axis_pos(z, (u32) p.g.Z - 1, p.g.amin[2], p.g.asize[2]));
    return __ldg(values + ((i64) x * (u32) p.g.Y + y) * (u32) p.g.Z + z);
}
__host__ __device__ __forceinline__ u32 sort_nsub(const DenseParams &p) { return p.gy + p.gx; }
__host__ __device__ __forceinline__ u32 sort_buckets(const DenseParams &p) { return ((u32) p.g.X + 2) * (p.gy + p.gx); }

constexpr u32 ENT_Z_MASK = 0xffffu;
__device__ __forceinline__ u32 ent_z(u32 w) { return w & ENT_Z_MASK; }
__device__ __forceinline__ u32 ent_case(u32 w) { return (w >> 16) & 0xffu; }
__device__ __forceinline__ u32 ent_own(u32 w) { return (w >> 24) & 7u; }   // bit0 +z, bit1 +y, bit2 +x
__device__ __forceinline__ u32 ent_cell(u32 w) { return (w >> 27) & 1u; }

// corner pairs of the 12 cell edges (cell convention of the reference, include/shared_luts.cuh:3-38);
// constexpr so that unrolled loops over e fold them to immediates
__host__ __device__ constexpr int edge_c0(int e) {
    constexpr int t[12] = {0, 1, 2, 0, 4, 5, 6, 4, 0, 1, 3, 2};
    return t[e];
}
__host__ __device__ constexpr int edge_c1(int e) {
    constexpr int t[12] = {1, 3, 3, 2, 5, 7, 7, 6, 4, 5, 7, 6};
    return t[e];
}

// sign-change mask of the 12 edges for a case (== the reference's edge_status_table, see tools/gen_luts.py)
__device__ __forceinline__ u32 edge_mask_of_case(u32 c) {
    u32 m = 0;
#pragma unroll
    for (int e = 0; e < 12; e++) {
        m |= (((c >> edge_c0(e)) ^ (c >> edge_c1(e))) & 1u) << e;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// K1: values -> sign bits.  The one volume-sized HBM stream of the whole pipeline (4 B/voxel read,
// 1/8 B/voxel written).  Each lane loads one float4 (a warp covers 128 consecutive z, 512 B,
// fully coalesced), builds a nibble, and 8-lane OR-reductions assemble 32-bit words.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ u32 nibble_of(float4 a, float level) {
    return (u32) (__fsub_rn(a.x, level) < 0.0f) | ((u32) (__fsub_rn(a.y, level) < 0.0f) << 1) |
           ((u32) (__fsub_rn(a.z, level) < 0.0f) << 2) | ((u32) (__fsub_rn(a.w, level) < 0.0f) << 3);
}
__device__ __forceinline__ u32 gather_word(u32 nib, u32 lane) {
    u32 w = nib << (4 * (lane & 7));
    w |= __shfl_xor_sync(0xffffffffu, w, 1);
    w |= __shfl_xor_sync(0xffffffffu, w, 2);
    w |= __shfl_xor_sync(0xffffffffu, w, 4);
    return w;
}

// 256-bit streaming load (sm_100a): 8 consecutive floats per lane, no L1 allocation, L2 evict-first
__device__ __forceinline__ void ld_stream_f8(const float *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// Per-span summary (one byte per 128-point span of the flat point array, written beside the sign bits): 0 = all 128
// bits clear, 1 = all set, 2 = mixed.  The compaction decides "this part of the volume holds no entry" from the
// summary alone (1/128 B per voxel instead of the 1/8 B per voxel of the bits, which no longer fit the L2 at 2048^3).
constexpr u32 SUM_MIXED = 2u;
__device__ __forceinline__ u32 span_summary(u32 zero_lanes, u32 ones_lanes, u32 mask) {
    return (zero_lanes & mask) == mask ? 0u : ((ones_lanes & mask) == mask ? 1u : SUM_MIXED);
}

// tail group (possibly empty) + one zero group of padding so that readers may over-fetch
__device__ __forceinline__ void signbits_tail(const float *__restrict__ v, u32 *__restrict__ bits, unsigned char *__restrict__ sum, i64 P,
                                              float level, i64 first_tail_group, i64 ngroups_total, u32 lane, bool pad) {
    if (!pad && first_tail_group >= ngroups_total && (P & 127) == 0) return;   // sub-range of a larger volume: nothing to finish
    for (i64 g = first_tail_group; g <= ngroups_total; g++) {   // remaining full groups + the partial one
        i64 base = g << 7;
        u32 nib = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            i64 i = base + lane * 4 + k;
            if (i < P) nib |= (u32) (__fsub_rn(v[i], level) < 0.0f) << k;
        }
        u32 w = gather_word(nib, lane);
        if ((lane & 7) == 0) bits[(g << 2) + (lane >> 3)] = w;
        const u32 bz = __ballot_sync(0xffffffffu, w == 0u), bo = __ballot_sync(0xffffffffu, w == 0xffffffffu);
        if (lane == 0) sum[g] = (unsigned char) (g < ngroups_total ? span_summary(bz, bo, 0xffffffffu) : SUM_MIXED);   // partial group: never trusted
    }
    if (pad && (lane & 7) == 0) bits[((ngroups_total + 1) << 2) + (lane >> 3)] = 0u;
    if (pad && lane == 0) sum[ngroups_total + 1] = (unsigned char) SUM_MIXED;
}

// VEC8 = false: one float4 per lane per step (warp = 128 points); true: one 256-bit load (warp = 256 points).
template <bool VEC8, int UNROLL>
static __global__ void __launch_bounds__(256, 6) k_signbits_t(const float *__restrict__ v, u32 *__restrict__ bits,
                                                           unsigned char *__restrict__ sum, i64 P, float level, bool pad) {
    pdl_trigger();   // the count pass may be scheduled while the stream drains (it waits for completion: pdl_wait)
    const u32 lane = threadIdx.x & 31;
    const i64 ngroups = P >> 7;   // full groups of 128 points
    const i64 nwarps = ((i64) gridDim.x * blockDim.x) >> 5;
    const i64 wg = ((i64) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (!VEC8) {
        for (i64 g = wg * UNROLL; g < ngroups; g += nwarps * UNROLL) {
            float4 a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (g + u < ngroups) a[u] = ld_stream_f4(v + ((g + u) << 7) + lane * 4);
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (g + u < ngroups) {
                    u32 w = gather_word(nibble_of(a[u], level), lane);
                    if ((lane & 7) == 0) bits[((g + u) << 2) + (lane >> 3)] = w;
                    const u32 bz = __ballot_sync(0xffffffffu, w == 0u), bo = __ballot_sync(0xffffffffu, w == 0xffffffffu);
                    if (lane == 0) sum[g + u] = (unsigned char) span_summary(bz, bo, 0xffffffffu);
                }
        }
        if (wg == 0) signbits_tail(v, bits, sum, P, level, ngroups, ngroups, lane, pad);
    } else {
        const i64 npairs = ngroups >> 1;   // units of 256 points
        unsigned short *sum2 = reinterpret_cast<unsigned short *>(sum);   // two spans per warp step (the array is 2-byte aligned)
        for (i64 g = wg * UNROLL; g < npairs; g += nwarps * UNROLL) {
            float4 a[UNROLL], b[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (g + u < npairs) ld_stream_f8(v + ((g + u) << 8) + lane * 8, a[u], b[u]);
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
                if (g + u < npairs) {
                    // lane holds 8 consecutive bits; 4 lanes make one 32-bit word, the warp makes 8 words
                    u32 byte = nibble_of(a[u], level) | (nibble_of(b[u], level) << 4);
                    const u32 bz = __ballot_sync(0xffffffffu, byte == 0u), bo = __ballot_sync(0xffffffffu, byte == 0xffu);
                    u32 w = byte << (8 * (lane & 3));
                    w |= __shfl_xor_sync(0xffffffffu, w, 1);
                    w |= __shfl_xor_sync(0xffffffffu, w, 2);
                    if ((lane & 3) == 0) bits[((g + u) << 3) + (lane >> 2)] = w;
                    if (lane == 0)
                        sum2[g + u] = (unsigned short) (span_summary(bz, bo, 0x0000ffffu) | (span_summary(bz, bo, 0xffff0000u) << 8));
                }
        }
        if (wg == 0) signbits_tail(v, bits, sum, P, level, npairs << 1, ngroups, lane, pad);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 for implicit fields: sign bits straight from the analytic program, no field in HBM.  One thread per 32-bit
// word of the flat bit array.  If the word's 32 points lie in one row, the program is evaluated ONCE at the middle of
// the word: every supported operator is 1-Lipschitz, so when |f(mid) - level| exceeds (half the word's length) x L
// plus a float-error margin all 32 points have the sign of the middle and the word is written without evaluating them
// (> 95 % of the words of a typical volume).  Otherwise (and for words straddling two rows) the points are evaluated.
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_sdf_bits(DenseParams p, u32 *__restrict__ bits, unsigned char *__restrict__ sum) {
    pdl_wait();
    pdl_trigger();
    // Two phases per block of 256 consecutive words.  Phase 1, thread per word: one evaluation at the middle of the
    // word decides it if the surface is provably farther than half a word (1-Lipschitz bound + float margin);
    // undecided words go to a shared list.  Phase 2, warp per listed word: the 32 lanes evaluate the 32 points at
    // once and a ballot assembles the word (a thread-per-word loop over 32 evaluations left 31 lanes of most warps
    // idle: 2.2 ms instead of 0.x ms at 1024^3).
    __shared__ u32 s_n;
    __shared__ u32 s_list[256];
    const i64 P = p.P, nwords = (P + 31) >> 5;
    const i64 nall = (i64) (((P >> 7) + 2) << 2);             // = signbits_bit_words(P): the array is zero-padded to here
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    const u32 resx = (u32) p.g.Xg - 1, resy = Y - 1, resz = Z - 1;
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // rows made of whole words (Z % 32 == 0, nwords < 2^32): 32-bit index arithmetic, one division per word
    const bool aligned = (Z & 31u) == 0u && nwords < ((i64) 1 << 32);
    const u32 wpr = Z >> 5;
    for (i64 base = (i64) blockIdx.x * 256; base < nall; base += (i64) gridDim.x * 256) {
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        const i64 w = base + threadIdx.x;
        if (w < nall) {
            if (w >= nwords) {
                bits[w] = 0u;                                  // padding words (readers may over-fetch, see signbits_tail)
            } else {
                u32 x, y, z0;
                bool inrow;
                if (aligned) {
                    const u32 r = (u32) w / wpr;
                    z0 = ((u32) w - r * wpr) << 5;
                    x = r / Y; y = r - x * Y;
                    inrow = true;
                } else {
                    const i64 n0 = w << 5, r = n0 / Z;
                    z0 = (u32) (n0 - r * Z);
                    y = (u32) (r % Y); x = (u32) (r / Y);
                    inrow = z0 + 32u <= Z;
                }
                bool decided = false;
                if (inrow) {                                   // the word lies in one row
                    const float fx = axis_pos(x + (u32) p.g.x_off, resx, p.g.amin[0], p.g.asize[0]), fy = axis_pos(y, resy, p.g.amin[1], p.g.asize[1]);
                    const float za = axis_pos(z0, resz, p.g.amin[2], p.g.asize[2]), zb = axis_pos(z0 + 31u, resz, p.g.amin[2], p.g.asize[2]);
                    const float mid = 0.5f * (za + zb), half = 0.5f * fabsf(zb - za);
                    const float fm = sdf_eval(p.sdf, fx, fy, mid);
                    // margin: evaluation error of f (a few ulp of its operands' magnitude) and of the positions
                    const float margin = 1e-4f * (1.0f + fabsf(fm) + fabsf(p.level) + fabsf(fx) + fabsf(fy) + fabsf(mid));
                    if (fabsf(fm - p.level) > p.sdf->lipschitz * half + margin) {
                        bits[w] = (__fsub_rn(fm, p.level) < 0.0f) ? 0xffffffffu : 0u;
                        decided = true;
                    }
                }
                if (!decided) s_list[atomicAdd(&s_n, 1u)] = threadIdx.x;
            }
        }
        __syncthreads();
        const u32 nlist = s_n;
        for (u32 i = warp; i < nlist; i += 8) {
            const i64 ww = base + s_list[i];
            bool neg = false;
            if (aligned) {
                const u32 r = (u32) ww / wpr;
                const u32 z0 = ((u32) ww - r * wpr) << 5, x = r / Y, y = r - x * Y;
                neg = __fsub_rn(field_value<true>(nullptr, p, x, y, z0 + lane), p.level) < 0.0f;
            } else {
                const i64 n = (ww << 5) + lane;
                if (n < P) {
                    const i64 rr = n / Z;
                    neg = __fsub_rn(field_value<true>(nullptr, p, (u32) (rr / Y), (u32) (rr % Y), (u32) (n - rr * Z)), p.level) < 0.0f;
                }
            }
            const u32 word = __ballot_sync(0xffffffffu, neg);
            if (lane == 0) bits[ww] = word;
        }
        __syncthreads();
        if (threadIdx.x < 64) {                                // per-span summary of the block's 64 spans (span_summary)
            const i64 sp = (base >> 2) + threadIdx.x, full = P >> 7;
            if (sp <= full + 1) {
                u32 sv = SUM_MIXED;
                if (sp < full) {
                    const u32 w0 = bits[4 * sp], w1 = bits[4 * sp + 1], w2 = bits[4 * sp + 2], w3 = bits[4 * sp + 3];
                    const u32 any = w0 | w1 | w2 | w3, all = w0 & w1 & w2 & w3;
                    sv = any == 0u ? 0u : (all == 0xffffffffu ? 1u : SUM_MIXED);
                }
                sum[sp] = (unsigned char) sv;
            }
        }
        __syncthreads();
    }
}

// materialise the field of an implicit grid (the two-step path; parity reference of the fused path)
static __global__ void __launch_bounds__(256) k_sdf_fill(DenseParams p, float *__restrict__ out) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (i64 n = (i64) blockIdx.x * blockDim.x + threadIdx.x; n < p.P; n += (i64) gridDim.x * blockDim.x) {
        const u32 z = (u32) (n % Z);
        const i64 r = n / Z;
        out[n] = field_value<true>(nullptr, p, (u32) (r / Y), (u32) (r % Y), z);
    }
}

int device_sms();
extern int g_signbits_variant;   // tuning knob (api.cu); 0 = default
// Launch the volume-streaming kernel (timed by the bench hooks).
// pad = false: the range is a piece (a multiple of 256 points) of a larger volume whose last piece writes the padding
// sum: the span summary of the range (span_sum_of(bits base, P total) + first span of the range)
static inline void launch_signbits(const float *values, u32 *bits, unsigned char *sum, i64 P, float level, cudaStream_t stream, bool pad = true) {
    const int sms = device_sms();
    const int var = g_signbits_variant;
    const int per_sm = (var >> 8) ? (var >> 8) : 32;          // blocks per SM to launch (tools/tune_signbits.py)
    i64 groups = P >> 7;
    i64 want = (groups + 31) / 32;
    int blocks = (int) (want < 1 ? 1 : (want > (i64) sms * per_sm ? (i64) sms * per_sm : want));
    stream_timer_mark(stream);
    switch (var & 0xff) {
        case 1: ISX_LAUNCH((k_signbits_t<false, 8>), blocks, 256, 0, stream, values, bits, sum, P, level, pad); break;
        case 3: ISX_LAUNCH((k_signbits_t<true, 4>), blocks, 256, 0, stream, values, bits, sum, P, level, pad); break;
        case 4: ISX_LAUNCH((k_signbits_t<false, 4>), blocks, 256, 0, stream, values, bits, sum, P, level, pad); break;
        case 5: ISX_LAUNCH((k_signbits_t<true, 1>), blocks, 256, 0, stream, values, bits, sum, P, level, pad); break;
        // default: 256-bit loads, 2 in flight per lane -- 5.9 TB/s at 512^3, 6.5 TB/s at 1024^3 on B200.  Measured and
        // dropped (profiles/r2_signbits_variants.txt): software-pipelined loads (6.3 TB/s) and per-warp TMA rings of
        // cp.async.bulk + mbarrier stages (5.7-6.4 TB/s): the plain loop already sits at the copy peak of the part
        default: ISX_LAUNCH((k_signbits_t<true, 2>), blocks, 256, 0, stream, values, bits, sum, P, level, pad); break;
    }
    stream_timer_mark(stream);
}
// u32 words of the sign-bit array (incl. one zero group of padding) followed by the per-span summary bytes
static inline size_t signbits_bit_words(i64 P) { return (size_t) (((P >> 7) + 2) << 2); }
static inline size_t signbits_words(i64 P) { return signbits_bit_words(P) + (size_t) (((P >> 7) + 2 + 3) >> 2) + 4; }
static inline unsigned char *span_sum_of(u32 *bits, i64 P) { return reinterpret_cast<unsigned char *>(bits + signbits_bit_words(P)); }
static inline const unsigned char *span_sum_of(const u32 *bits, i64 P) { return reinterpret_cast<const unsigned char *>(bits + signbits_bit_words(P)); }
// volume pass of an implicit field (timed like the streaming kernel)
static inline void launch_sdf_bits(const DenseParams &p, u32 *bits, cudaStream_t stream) {
    const i64 nwords = (i64) signbits_bit_words(p.P);
    i64 want = (nwords + 255) / 256;
    const i64 cap = (i64) device_sms() * 64;
    stream_timer_mark(stream);
    ISX_LAUNCH(k_sdf_bits, (int) (want > cap ? cap : want), 256, 0, stream, p, bits, span_sum_of(bits, p.P));
    stream_timer_mark(stream);
}

// bits n..n+31 -> b0 ; bits n+1..n+32 -> b1
__device__ __forceinline__ void fetch33(const u32 *__restrict__ bits, i64 n, u32 &b0, u32 &b1) {
    const i64 w = n >> 5;
    const u32 sh = (u32) n & 31u;
    u32 w0 = __ldg(bits + w), w1 = __ldg(bits + w + 1), w2 = __ldg(bits + w + 2);
    b0 = __funnelshift_r(w0, w1, sh);
    u32 hi = __funnelshift_r(w1, w2, sh);
    b1 = (b0 >> 1) | (hi << 31);
}

// ---------------------------------------------------------------------------------------------
// K2: sign bits -> ordered entry list + row_start, in ONE pass (decoupled look-back scan).
// One thread per 32-point chunk of a row; warp-free bit tricks give the active-cell mask and the
// owned sign-change edges of 32 points at once.
// ---------------------------------------------------------------------------------------------
constexpr int CP_ITEMS = 8;                 // consecutive 32-point chunks per thread
constexpr int CP_TILE = 256 * CP_ITEMS;     // chunks per block

struct ChunkBits {
    u32 a0, a1, b0, b1, c0, c1, d0, d1;   // sign bits of rows (x,y),(x,y+1),(x+1,y),(x+1,y+1) at z / z+1
    u32 ez, ey, ex, cellact;              // owned sign-change edges, active cells
    u32 r, z0;
    bool cellrow;
};

__device__ __forceinline__ u32 chunk_classify(const u32 *__restrict__ bits, const DenseParams &p, u32 q, ChunkBits &k) {
    const u32 X = (u32) p.g.X, Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    k.a0 = k.a1 = k.b0 = k.b1 = k.c0 = k.c1 = k.d0 = k.d1 = 0;
    k.ez = k.ey = k.ex = k.cellact = 0;
    k.r = q / p.CPR;
    const u32 c = q - k.r * p.CPR;
    const u32 x = k.r / Y, y = k.r - x * Y;
    k.z0 = c * 32u;
    const i64 n0 = (i64) k.r * Z + k.z0;
    const u32 nvalid = min(32u, Z - k.z0);            // points of this chunk inside the row
    const u32 nv1 = min(32u, Z - k.z0 - 1u);          // ... that also have a z+1 neighbour
    const u32 mz = nvalid == 32u ? 0xffffffffu : ((1u << nvalid) - 1u);
    const u32 mz1 = nv1 == 32u ? 0xffffffffu : ((1u << nv1) - 1u);
    const bool hasY = y + 1u < Y, hasX = x + 1u < X;
    k.cellrow = hasX && hasY;
    fetch33(bits, n0, k.a0, k.a1);
    k.ez = (k.a0 ^ k.a1) & mz1;
    if (hasY) {
        fetch33(bits, n0 + Z, k.b0, k.b1);
        k.ey = (k.a0 ^ k.b0) & mz;
    }
    if (hasX) {
        fetch33(bits, n0 + p.YZ, k.c0, k.c1);
        k.ex = (k.a0 ^ k.c0) & mz;
    }
    if (hasX && hasY) {
        fetch33(bits, n0 + p.YZ + Z, k.d0, k.d1);
        u32 any = k.a0 | k.a1 | k.b0 | k.b1 | k.c0 | k.c1 | k.d0 | k.d1;
        u32 all = k.a0 & k.a1 & k.b0 & k.b1 & k.c0 & k.c1 & k.d0 & k.d1;
        k.cellact = any & ~all & mz1;
    }
    return k.cellact | k.ez | k.ey | k.ex;
}

static __global__ void __launch_bounds__(256) k_compact(const u32 *__restrict__ bits, DenseParams p, uint2 *__restrict__ entries,
                                                 u32 cap, u32 *__restrict__ row_start, u64 *__restrict__ desc,
                                                 u32 *__restrict__ counters) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_A], 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u32 q0 = tile * CP_TILE + threadIdx.x * CP_ITEMS;
    const u32 Z = (u32) p.g.Z;
    // pass 1: count (the masks are recomputed in pass 2; the words stay in L1)
    u32 cnt = 0;
#pragma unroll
    for (int j = 0; j < CP_ITEMS; j++) {
        const u32 q = q0 + j;
        if (q < p.NQ) {
            ChunkBits k;
            cnt += __popc(chunk_classify(bits, p, q, k));
        }
    }
    u32 total;
    const u32 excl = block_exclusive_scan(cnt, &total, sw);
    if (threadIdx.x < 32) {
        u32 pre = lookback_exclusive(desc, 1, tile, total, 1u);
        if (threadIdx.x == 0) s_prefix = pre;
    }
    __syncthreads();
    u32 off = s_prefix + excl;
#pragma unroll 1
    for (int j = 0; j < CP_ITEMS; j++) {
        const u32 q = q0 + j;
        if (q >= p.NQ) break;
        ChunkBits k;
        u32 m = chunk_classify(bits, p, q, k);
        if (k.z0 == 0) row_start[k.r] = off;
        while (m) {
            const u32 b = __ffs(m) - 1;
            m &= m - 1;
            if (off < cap) {
                u32 cs = ((k.a0 >> b) & 1u) | (((k.a1 >> b) & 1u) << 1) | (((k.b0 >> b) & 1u) << 2) | (((k.b1 >> b) & 1u) << 3) |
                         (((k.c0 >> b) & 1u) << 4) | (((k.c1 >> b) & 1u) << 5) | (((k.d0 >> b) & 1u) << 6) | (((k.d1 >> b) & 1u) << 7);
                u32 own = ((k.ez >> b) & 1u) | (((k.ey >> b) & 1u) << 1) | (((k.ex >> b) & 1u) << 2);
                u32 cv = (k.cellrow && (k.z0 + b + 1u < Z)) ? 1u : 0u;
                entries[off] = make_uint2(k.r, (k.z0 + b) | (cs << 16) | (own << 24) | (cv << 27));
            }
            off++;
        }
        if (q == p.NQ - 1) {
            row_start[p.R] = off;
            counters[C_S] = off;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2 fast path (Z % 128 == 0): the unit of work is a 128-point span of a row.  The four rows of the
// span's cells arrive as four aligned uint4 loads (+1 word for the z+1 neighbour of the last point);
// a span whose 4x(128+1) sign bits are all equal -- >98% of a typical volume -- owns no entry, and is
// rejected from the one-byte span summary without touching the bits at all.
// ---------------------------------------------------------------------------------------------
struct SpanBits {
    u32 a[5], b[5], c[5], d[5];   // words 0..3 of the span + the following word (only bit 0 used)
    u32 r, z0;
    bool hasX, hasY, last;        // last: this span ends the row (no z+1 neighbour for its last point)
};

// masks of word j of the span: m = entries, plus per-bit pieces
__device__ __forceinline__ u32 span_word_masks(const SpanBits &k, int j, u32 &a0, u32 &a1, u32 &b0, u32 &b1, u32 &c0, u32 &c1,
                                               u32 &d0, u32 &d1, u32 &ez, u32 &ey, u32 &ex, u32 &cellact) {
    a0 = k.a[j]; a1 = (k.a[j] >> 1) | (k.a[j + 1] << 31);
    b0 = k.b[j]; b1 = (k.b[j] >> 1) | (k.b[j + 1] << 31);
    c0 = k.c[j]; c1 = (k.c[j] >> 1) | (k.c[j + 1] << 31);
    d0 = k.d[j]; d1 = (k.d[j] >> 1) | (k.d[j + 1] << 31);
    const u32 mz1 = (k.last && j == 3) ? 0x7fffffffu : 0xffffffffu;   // points with a z+1 neighbour
    ez = (a0 ^ a1) & mz1;
    ey = k.hasY ? (a0 ^ b0) : 0u;
    ex = k.hasX ? (a0 ^ c0) : 0u;
    cellact = 0;
    if (k.hasX && k.hasY) {
        u32 any = a0 | a1 | b0 | b1 | c0 | c1 | d0 | d1;
        u32 all = a0 & a1 & b0 & b1 & c0 & c1 & d0 & d1;
        cellact = any & ~all & mz1;
    }
    return cellact | ez | ey | ex;
}

// ---- span fast path, three stages (Z % 128 == 0) -----------------------------------------------
//   k_rowcount_blk  block = 256 rows: span summary (written by k_signbits) -> candidate spans, classified from the
//                   sign bits with one thread per span -> entry count of every row + per-span counts
//                   (k_rowcount_sum: thread-per-row variant for rows of more than 32 spans)
//   k_scan_rows     decoupled look-back exclusive scan over the X*Y row counts (in place) -> row_start, S
//   k_rowfill_cnt   thread = row that owns entries: re-classifies its non-empty spans and writes the entries
//                   (k_rowfill128: the variant driven by the summary alone, rows of more than 32 spans)
// (a single-pass chained-scan version of this stage spent its time at barriers waiting for the
//  look-back warp and ran 3x slower at 1024^3; fusing only the row scan into the count kernel: +10 us; see profiles/)
struct SpanRows {
    const uint4 *a, *c;   // row (x,y) in plane x and in plane x+1 (aliases plane x when x+1 does not exist)
    u32 dy;               // uint4 offset of row y+1 (0 when it does not exist)
    bool hasX, hasY, last;
};

// global-memory variant of span_load (rows addressed through SpanRows)
__device__ __forceinline__ void span_load_g(const SpanRows &q, u32 r, u32 c4, SpanBits &k) {
    k.r = r;
    k.z0 = c4 * 128u;
    k.hasX = q.hasX; k.hasY = q.hasY; k.last = q.last;
    const uint4 A = __ldg(q.a + c4), B = __ldg(q.a + q.dy + c4), Cc = __ldg(q.c + c4), D = __ldg(q.c + q.dy + c4);
    k.a[0] = A.x; k.a[1] = A.y; k.a[2] = A.z; k.a[3] = A.w;
    k.b[0] = B.x; k.b[1] = B.y; k.b[2] = B.z; k.b[3] = B.w;
    k.c[0] = Cc.x; k.c[1] = Cc.y; k.c[2] = Cc.z; k.c[3] = Cc.w;
    k.d[0] = D.x; k.d[1] = D.y; k.d[2] = D.z; k.d[3] = D.w;
    if (q.last) {
        k.a[4] = k.b[4] = k.c[4] = k.d[4] = 0u;
    } else {
        k.a[4] = __ldg(reinterpret_cast<const u32 *>(q.a + c4 + 1));
        k.b[4] = __ldg(reinterpret_cast<const u32 *>(q.a + q.dy + c4 + 1));
        k.c[4] = __ldg(reinterpret_cast<const u32 *>(q.c + c4 + 1));
        k.d[4] = __ldg(reinterpret_cast<const u32 *>(q.c + q.dy + c4 + 1));
    }
}

// number of entries of a span; if entries != nullptr also writes them starting at off
static __device__ __noinline__ u32 span_detail(const SpanRows q, u32 r, u32 c4, u32 Z, uint2 *__restrict__ entries, u32 cap, u32 off) {
    SpanBits k;
    span_load_g(q, r, c4, k);
    const bool cellrow = k.hasX && k.hasY;
    u32 n = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u32 a0, a1, b0, b1, cc0, c1, d0, d1, ez, ey, ex, ca;
        u32 m = span_word_masks(k, j, a0, a1, b0, b1, cc0, c1, d0, d1, ez, ey, ex, ca);
        n += __popc(m);
        if (entries) {
            const u32 zb = k.z0 + 32u * j;
            while (m) {
                const u32 b = __ffs(m) - 1;
                m &= m - 1;
                if (off < cap) {
                    u32 cs = ((a0 >> b) & 1u) | (((a1 >> b) & 1u) << 1) | (((b0 >> b) & 1u) << 2) | (((b1 >> b) & 1u) << 3) |
                             (((cc0 >> b) & 1u) << 4) | (((c1 >> b) & 1u) << 5) | (((d0 >> b) & 1u) << 6) | (((d1 >> b) & 1u) << 7);
                    if (!cellrow) cs &= 0x03u;   // rows y+1 / x+1 do not exist: their bits are copies
                    u32 own = ((ez >> b) & 1u) | (((ey >> b) & 1u) << 1) | (((ex >> b) & 1u) << 2);
                    u32 cv = (cellrow && (zb + b + 1u < Z)) ? 1u : 0u;
                    entries[off] = make_uint2(r, (zb + b) | (cs << 16) | (own << 24) | (cv << 27));
                }
                off++;
            }
        }
    }
    return n;
}

// count-only twin of span_detail (the count pass is bound by the latency of these loads: fewer registers = more
// spans in flight per SM)
__device__ __forceinline__ u32 span_count(const SpanRows &q, u32 c4) {
    SpanBits k;
    span_load_g(q, 0, c4, k);
    u32 n = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u32 a0, a1, b0, b1, cc0, c1, d0, d1, ez, ey, ex, ca;
        n += __popc(span_word_masks(k, j, a0, a1, b0, b1, cc0, c1, d0, d1, ez, ey, ex, ca));
    }
    return n;
}

__device__ __forceinline__ SpanRows span_rows(const u32 *__restrict__ bits, const DenseParams &p, u32 x, u32 y, u32 spr) {
    SpanRows q;
    q.hasY = y + 1u < (u32) p.g.Y;
    q.hasX = x + 1u < (u32) p.g.X;
    q.last = false;
    const u64 r = (u64) x * (u32) p.g.Y + y;
    q.a = reinterpret_cast<const uint4 *>(bits) + r * spr;
    q.c = q.a + (q.hasX ? (u64) (u32) p.g.Y * spr : 0ull);
    q.dy = q.hasY ? spr : 0u;
    return q;
}

// Summary rows round the cells of point row (x, y): (x,y), (x,y+1), (x+1,y), (x+1,y+1); a missing row aliases an
// existing one, exactly like the bit rows of span_rows().
struct SumRows { const unsigned char *a, *b, *c, *d; };
__device__ __forceinline__ SumRows sum_rows(const unsigned char *__restrict__ sum, const DenseParams &p, u32 x, u32 y, u32 spr) {
    SumRows q;
    q.a = sum + ((size_t) x * (u32) p.g.Y + y) * spr;
    q.b = q.a + (y + 1u < (u32) p.g.Y ? spr : 0u);
    const size_t dx = x + 1u < (u32) p.g.X ? (size_t) (u32) p.g.Y * spr : 0;
    q.c = q.a + dx;
    q.d = q.b + dx;
    return q;
}
// True if span c4 of the row provably owns no entry: its 4 x 128 sign bits and the z+1 neighbours of its last points
// are all equal.  Conservative (a mixed neighbour span may still start with the right bit): spans that fail the test
// are classified from the bits (span_detail), so the result is exact either way.
__device__ __forceinline__ bool span_provably_empty(const SumRows &q, u32 c4, bool last) {
    const u32 v = __ldg(q.a + c4);
    if (v >= SUM_MIXED || __ldg(q.b + c4) != v || __ldg(q.c + c4) != v || __ldg(q.d + c4) != v) return false;
    if (last) return true;
    return __ldg(q.a + c4 + 1) == v && __ldg(q.b + c4 + 1) == v && __ldg(q.c + c4 + 1) == v && __ldg(q.d + c4 + 1) == v;
}
template <int G> struct SumWord;
template <> struct SumWord<8> { using T = unsigned long long; static constexpr T ONES = 0x0101010101010101ull; };
template <> struct SumWord<4> { using T = u32; static constexpr T ONES = 0x01010101u; };
template <> struct SumWord<1> { using T = unsigned char; static constexpr T ONES = 1u; };

// Entry count of every point row, from the span summary: one thread per row; G spans are tested per load (G divides
// the spans per row) and >98 % of a typical volume is rejected with four loads per G*128 points; only spans round the
// surface touch the sign bits.  Writes row_count[r] for every row (no atomics, nothing to zero).
template <int G>
static __global__ void __launch_bounds__(256) k_rowcount_sum(const u32 *__restrict__ bits, const unsigned char *__restrict__ sum, DenseParams p,
                                                             u32 *__restrict__ row_count) {
    pdl_wait();
    pdl_trigger();
    using T = typename SumWord<G>::T;
    const u32 r = blockIdx.x * 256u + threadIdx.x;
    if (r >= p.R) return;
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z, spr = Z >> 7;
    const u32 x = r / Y, y = r - x * Y;
    const SumRows q = sum_rows(sum, p, x, y, spr);
    u32 n = 0;
    for (u32 g0 = 0; g0 < spr; g0 += G) {
        const T A = __ldg(reinterpret_cast<const T *>(q.a + g0)), B = __ldg(reinterpret_cast<const T *>(q.b + g0));
        const T C = __ldg(reinterpret_cast<const T *>(q.c + g0)), D = __ldg(reinterpret_cast<const T *>(q.d + g0));
        bool empty = A == B && A == C && A == D && (A == (T) 0 || A == SumWord<G>::ONES);
        if (empty && g0 + G < spr) {
            const u32 v = (u32) (A & (T) 0xff);
            empty = __ldg(q.a + g0 + G) == v && __ldg(q.b + g0 + G) == v && __ldg(q.c + g0 + G) == v && __ldg(q.d + g0 + G) == v;
        }
        if (empty) continue;
#pragma unroll 1
        for (u32 c4 = g0; c4 < g0 + G; c4++) {
            const bool last = c4 + 1u == spr;
            if (span_provably_empty(q, c4, last)) continue;
            SpanRows qb = span_rows(bits, p, x, y, spr);
            qb.last = last;
            n += span_detail(qb, r, c4, Z, nullptr, 0, 0);
        }
    }
    row_count[r] = n;
}

// Bit c4 set <=> span c4 of the row may own entries (rows of at most 32 spans).  G spans are tested per load.
template <int G>
__device__ __forceinline__ u32 row_span_mask(const SumRows &q, u32 spr) {
    using T = typename SumWord<G>::T;
    u32 m = 0;
    for (u32 g0 = 0; g0 < spr; g0 += G) {
        const T A = __ldg(reinterpret_cast<const T *>(q.a + g0)), B = __ldg(reinterpret_cast<const T *>(q.b + g0));
        const T C = __ldg(reinterpret_cast<const T *>(q.c + g0)), D = __ldg(reinterpret_cast<const T *>(q.d + g0));
        bool empty = A == B && A == C && A == D && (A == (T) 0 || A == SumWord<G>::ONES);
        if (empty && g0 + G < spr) {
            const u32 v = (u32) (A & (T) 0xff);
            empty = __ldg(q.a + g0 + G) == v && __ldg(q.b + g0 + G) == v && __ldg(q.c + g0 + G) == v && __ldg(q.d + g0 + G) == v;
        }
        if (empty) continue;
#pragma unroll 1
        for (u32 c4 = g0; c4 < g0 + G; c4++)
            if (!span_provably_empty(q, c4, c4 + 1u == spr)) m |= 1u << c4;
    }
    return m;
}

// Count, rows of <= 32 spans: a block takes 256 consecutive rows.  Stage 1, thread per row: span mask from the summary.
// Stage 2: the candidate spans of the block are listed in shared memory and classified with one thread per SPAN (the
// per-row loop left most lanes of a warp idle while one row worked through its spans).
// Stage 3, thread per row: row_count[r] = sum over the row's spans (every row is written: no atomics, nothing to zero);
// rows with candidate spans also write their per-span counts for the fill pass (span_cnt needs no zeroing either: the
// fill pass only reads the rows that own entries).
constexpr u32 RC_ROWS = 256;
template <int G>
static __global__ void __launch_bounds__(RC_ROWS, 6) k_rowcount_blk(const u32 *__restrict__ bits, const unsigned char *__restrict__ sum, DenseParams p,
                                                                 u32 *__restrict__ row_count, unsigned char *__restrict__ span_cnt) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sw[33];
    __shared__ unsigned short s_item[RC_ROWS * 32];
    __shared__ unsigned char s_n[RC_ROWS * 32];
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z, spr = Z >> 7;
    const u32 r0 = blockIdx.x * RC_ROWS, r = r0 + threadIdx.x;
    u32 mask = 0;
    if (r < p.R) {
        const u32 x = r / Y, y = r - x * Y;
        mask = row_span_mask<G>(sum_rows(sum, p, x, y, spr), spr);
    }
    const u32 cnt = __popc(mask);
    u32 total;
    const u32 ex0 = block_exclusive_scan(cnt, &total, sw);
    {
        u32 m = mask, e = ex0;
        while (m) {
            const u32 c4 = __ffs(m) - 1;
            m &= m - 1;
            s_item[e++] = (unsigned short) (threadIdx.x | (c4 << 8));
        }
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < total; i += RC_ROWS) {
        const u32 it = s_item[i], r2 = r0 + (it & 255u), c4 = it >> 8;
        const u32 x = r2 / Y, y = r2 - x * Y;
        SpanRows qb = span_rows(bits, p, x, y, spr);
        qb.last = c4 + 1u == spr;
        const u32 n = span_count(qb, c4);
        s_n[i] = (unsigned char) n;                       // <= 128
    }
    __syncthreads();
    if (r < p.R) {
        u32 n = 0;
        if (mask) {                                       // rows without candidate spans own no entry: the fill pass never reads their counts
            unsigned char *rc = span_cnt + (size_t) r * spr;
            u32 i = ex0;
            for (u32 c4 = 0; c4 < spr; c4++) {
                const u32 k = (mask >> c4) & 1u ? s_n[i++] : 0u;
                rc[c4] = (unsigned char) k;
                n += k;
            }
        }
        row_count[r] = n;
    }
}

// Fill, rows of <= 32 spans: one thread per row that owns entries (no grid-stride loop: a stride that is a multiple of Y
// would hand every full-face row of a CSG box to the same threads); only the spans with a non-zero count are
// re-classified.  (Listing the spans in shared memory and filling with one thread per span, like the count pass,
// was measured slower: 54 vs 40 us at 1024^3, 362 vs 120 us at 2048^3.)  HEAVY rows -- a row lying in an axis-aligned
// face owns up to Z entries -- are appended to a list and filled by k_rowfill_heavy with one warp per row.
constexpr u32 FILL_HEAVY = 160;   // entries per row above which the row goes to the heavy list
static __global__ void __launch_bounds__(128) k_rowfill_cnt(const u32 *__restrict__ bits, DenseParams p, const u32 *__restrict__ row_start,
                                                            const unsigned char *__restrict__ span_cnt, uint2 *__restrict__ entries, u32 cap,
                                                            u32 *__restrict__ heavy_list, u32 heavy_cap, u32 *__restrict__ n_heavy) {
    pdl_wait();
    pdl_trigger();
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z, spr = Z >> 7;
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    u32 off = row_start[r];
    const u32 end = row_start[r + 1];
    if (off == end) return;
    if (end - off > FILL_HEAVY) {
        const u32 slot = atomicAdd(n_heavy, 1u);
        if (slot < heavy_cap) heavy_list[slot] = r;   // heavy_cap >= cap / FILL_HEAVY + 1 always suffices when S <= cap
        return;
    }
    const u32 x = r / Y, y = r - x * Y;
    SpanRows q = span_rows(bits, p, x, y, spr);
    const unsigned char *rc = span_cnt + (size_t) r * spr;
    for (u32 c4 = 0; c4 < spr && off < end; c4++) {
        const u32 n = rc[c4];
        if (!n) continue;
        q.last = c4 + 1u == spr;
        span_detail(q, r, c4, Z, entries, cap, off);
        off += n;
    }
}

// in-place exclusive scan of n u32 (8 per thread, look-back across blocks); total -> counters[total_idx], data[n]
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = 256 * RS_ITEMS;
static __global__ void __launch_bounds__(256) k_scan_rows(u32 *__restrict__ data, u32 n, u64 *__restrict__ desc, u32 *__restrict__ counters,
                                                          int ticket_idx, int total_idx) {
    pdl_wait();
    pdl_trigger();
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_pre;
    const u32 ntiles = (n + RS_TILE - 1) / RS_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[ticket_idx], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 i0 = tile * RS_TILE + threadIdx.x * RS_ITEMS;
        u32 v[RS_ITEMS], sum = 0;
        if (i0 + RS_ITEMS <= n) {
            const uint4 lo = *reinterpret_cast<const uint4 *>(data + i0), hi = *reinterpret_cast<const uint4 *>(data + i0 + 4);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        } else {
#pragma unroll
            for (int j = 0; j < RS_ITEMS; j++) v[j] = i0 + j < n ? data[i0 + j] : 0u;
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) sum += v[j];
        u32 tot;
        u32 ex = block_exclusive_scan(sum, &tot, sw);
        if (threadIdx.x < 32) {
            u32 pre = lookback_exclusive(desc, 1, tile, tot, 1u);
            if (threadIdx.x == 0) s_pre = pre;
        }
        __syncthreads();
        ex += s_pre;
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            if (i0 + j < n) {
                data[i0 + j] = ex;
                ex += v[j];
                if (i0 + j == n - 1) {
                    data[n] = ex;
                    counters[total_idx] = ex;
                }
            }
        }
    }
}

// Fill, ordinary rows: one thread per row (no grid-stride loop: a stride that is a multiple of Y would hand
// every full-face row of a CSG box to the same threads); only the mixed spans (count byte != 0) are
// re-classified.  HEAVY rows -- a row lying in an axis-aligned face owns up to Z entries -- are appended to a
// list and filled by k_rowfill_heavy with one warp per row.
static __global__ void __launch_bounds__(128) k_rowfill128(const u32 *__restrict__ bits, DenseParams p, const u32 *__restrict__ row_start,
                                                           const unsigned char *__restrict__ sum, uint2 *__restrict__ entries, u32 cap,
                                                           u32 *__restrict__ heavy_list, u32 heavy_cap, u32 *__restrict__ n_heavy) {
    pdl_wait();
    pdl_trigger();
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z, spr = Z >> 7;
    const u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    u32 off = row_start[r];
    const u32 end = row_start[r + 1];
    if (off == end) return;
    if (end - off > FILL_HEAVY) {
        const u32 slot = atomicAdd(n_heavy, 1u);
        if (slot < heavy_cap) heavy_list[slot] = r;   // heavy_cap >= cap / FILL_HEAVY + 1 always suffices when S <= cap
        return;
    }
    const u32 x = r / Y, y = r - x * Y;
    SpanRows q = span_rows(bits, p, x, y, spr);
    const SumRows qs = sum_rows(sum, p, x, y, spr);
    for (u32 c4 = 0; c4 < spr && off < end; c4++) {
        const bool last = c4 + 1u == spr;
        if (span_provably_empty(qs, c4, last)) continue;
        q.last = last;
        off += span_detail(q, r, c4, Z, entries, cap, off);
    }
}

// Fill, heavy rows: warp per row, lane per 32-point word; a shuffle scan of the per-word entry counts gives
// the offsets, every lane emits at most 32 entries.
static __global__ void __launch_bounds__(256) k_rowfill_heavy(const u32 *__restrict__ bits, DenseParams p, const u32 *__restrict__ row_start,
                                                              uint2 *__restrict__ entries, u32 cap, const u32 *__restrict__ heavy_list,
                                                              u32 heavy_cap, const u32 *__restrict__ n_heavy) {
    pdl_wait();
    pdl_trigger();
    const u32 X = (u32) p.g.X, Y = (u32) p.g.Y, Z = (u32) p.g.Z, wpr = Z >> 5;   // words per row
    const u32 lane = threadIdx.x & 31;
    const u32 nwarps = (gridDim.x * blockDim.x) >> 5;
    u32 nh = *n_heavy;
    if (nh > heavy_cap) nh = heavy_cap;
    for (u32 h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < nh; h += nwarps) {
        const u32 r = heavy_list[h];
        const u32 x = r / Y, y = r - x * Y;
        const bool hasX = x + 1u < X, hasY = y + 1u < Y, cellrow = hasX && hasY;
        const u32 *ra = bits + (size_t) r * wpr;
        const u32 *rb = ra + (hasY ? wpr : 0u);
        const u32 *rcn = ra + (hasX ? (size_t) Y * wpr : 0);
        const u32 *rd = rcn + (hasY ? wpr : 0u);
        u32 off = row_start[r];
        for (u32 wb = 0; wb < wpr; wb += 32) {
            const u32 wi = wb + lane;
            u32 m = 0, a0 = 0, a1 = 0, b0 = 0, b1 = 0, c0 = 0, c1 = 0, d0 = 0, d1 = 0, ez = 0, ey = 0, ex = 0;
            if (wi < wpr) {
                const bool lastw = wi + 1u == wpr;
                const u32 an = lastw ? 0u : __ldg(ra + wi + 1), bn = lastw ? 0u : __ldg(rb + wi + 1);
                const u32 cn = lastw ? 0u : __ldg(rcn + wi + 1), dn = lastw ? 0u : __ldg(rd + wi + 1);
                a0 = __ldg(ra + wi); b0 = __ldg(rb + wi); c0 = __ldg(rcn + wi); d0 = __ldg(rd + wi);
                a1 = (a0 >> 1) | (an << 31); b1 = (b0 >> 1) | (bn << 31); c1 = (c0 >> 1) | (cn << 31); d1 = (d0 >> 1) | (dn << 31);
                const u32 mz1 = lastw ? 0x7fffffffu : 0xffffffffu;
                ez = (a0 ^ a1) & mz1;
                ey = hasY ? (a0 ^ b0) : 0u;
                ex = hasX ? (a0 ^ c0) : 0u;
                u32 cellact = 0;
                if (cellrow) {
                    const u32 any = a0 | a1 | b0 | b1 | c0 | c1 | d0 | d1, all = a0 & a1 & b0 & b1 & c0 & c1 & d0 & d1;
                    cellact = any & ~all & mz1;
                }
                m = cellact | ez | ey | ex;
            }
            const u32 n = __popc(m);
            u32 incl = n;
            for (u32 o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            u32 o = off + incl - n;
            const u32 zb = wi * 32u;
            while (m) {
                const u32 b = __ffs(m) - 1;
                m &= m - 1;
                if (o < cap) {
                    u32 cs = ((a0 >> b) & 1u) | (((a1 >> b) & 1u) << 1) | (((b0 >> b) & 1u) << 2) | (((b1 >> b) & 1u) << 3) |
                             (((c0 >> b) & 1u) << 4) | (((c1 >> b) & 1u) << 5) | (((d0 >> b) & 1u) << 6) | (((d1 >> b) & 1u) << 7);
                    if (!cellrow) cs &= 0x03u;
                    const u32 own = ((ez >> b) & 1u) | (((ey >> b) & 1u) << 1) | (((ex >> b) & 1u) << 2);
                    const u32 cv = (cellrow && (zb + b + 1u < Z)) ? 1u : 0u;
                    entries[o] = make_uint2(r, (zb + b) | (cs << 16) | (own << 24) | (cv << 27));
                }
                o++;
            }
            off += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

// host-side: the span path needs Z % 128 == 0
static inline bool compact128_ok(const DenseParams &p) { return (p.g.Z & 127) == 0 && p.g.Z >= 128; }

// first entry of row [lo,hi) whose z is >= z
__device__ __forceinline__ u32 row_lower_bound(const uint2 *__restrict__ entries, u32 lo, u32 hi, u32 z) {
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (ent_z(entries[mid].y) < z) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Cell-local data: 8 corner values + the 6 axis positions.
struct CellData {
    float v[8];
    float px[2], py[2], pz[2];
};

// same, for call sites that already hold the flat index n of a base point: explicit fields load values[n + off]
// exactly as before (no index arithmetic is added to the explicit-field instances)
template <bool IMPLICIT>
__device__ __forceinline__ float field_at(const float *__restrict__ values, const DenseParams &p, i64 n, i64 off, u32 x, u32 y, u32 z) {
    if (IMPLICIT) return field_value<true>(values, p, x, y, z);
    return __ldg(values + n + off);
}

template <bool IMPLICIT = false>
__device__ __forceinline__ void load_cell_values(const float *__restrict__ values, const DenseParams &p, u32 r, u32 z, CellData &c) {
    const u32 Z = (u32) p.g.Z;
    if (IMPLICIT) {
        const u32 Y = (u32) p.g.Y, x = r / Y, y = r - x * Y;
#pragma unroll 1
        for (int k = 0; k < 8; k++) c.v[k] = field_value<true>(values, p, x + ((k >> 2) & 1), y + ((k >> 1) & 1), z + (k & 1));
        return;
    }
    const i64 n = (i64) r * Z + z;
    c.v[0] = __ldg(values + n);
    c.v[1] = __ldg(values + n + 1);
    c.v[2] = __ldg(values + n + Z);
    c.v[3] = __ldg(values + n + Z + 1);
    c.v[4] = __ldg(values + n + p.YZ);
    c.v[5] = __ldg(values + n + p.YZ + 1);
    c.v[6] = __ldg(values + n + p.YZ + Z);
    c.v[7] = __ldg(values + n + p.YZ + Z + 1);
}
__device__ __forceinline__ void load_cell_positions(const DenseParams &p, u32 r, u32 z, CellData &c) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    const u32 x = r / Y, y = r - x * Y;
    const u32 xg = x + (u32) p.g.x_off;
    c.px[0] = axis_pos(xg, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
    c.px[1] = axis_pos(xg + 1, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
    c.py[0] = axis_pos(y, Y - 1, p.g.amin[1], p.g.asize[1]);
    c.py[1] = axis_pos(y + 1, Y - 1, p.g.amin[1], p.g.asize[1]);
    c.pz[0] = axis_pos(z, Z - 1, p.g.amin[2], p.g.asize[2]);
    c.pz[1] = axis_pos(z + 1, Z - 1, p.g.amin[2], p.g.asize[2]);
}
// True if no two crossing points of the cell can be bit-equal, so that no triangle of the cell is degenerate
// (src/mc/nagae.cu:71) and the positions need not be evaluated to know it.  Two points on different edges of a
// cell can only coincide near a corner the edges share.  With a = level - v0, d = v1 - v0 (the operands of the
// reference's t = a / d), 2 eps |d| <= |a| <= (1 - 2 eps) |d| gives eps <= t <= 1 - eps after rounding, hence
// the point is at least eps * cell - 2 ulp away from both end points along the edge axis, while a point on
// another edge through the same corner stays within 1 ulp of the corner along that axis.  eps * cell = 16 ulp of
// the largest coordinate (make_dense_params), so the two differ.  NaN fails the comparisons -> not plain.
__device__ __forceinline__ bool cell_is_plain(const CellData &c, u32 status, float level, const float (&eps2)[3]) {
    bool plain = true;
#pragma unroll
    for (int k = 0; k < 12; k++)
        if ((status >> k) & 1u) {
            const int p0 = edge_c0(k), p1 = edge_c1(k);
            const int axis = (p0 ^ p1) == 4 ? 0 : ((p0 ^ p1) == 2 ? 1 : 2);
            const float a = fabsf(__fsub_rn(level, c.v[p0])), d = fabsf(__fsub_rn(c.v[p1], c.v[p0]));
            const float e2 = eps2[axis];
            plain = plain && (a >= e2 * d) && (a <= (1.0f - e2) * d);
        }
    return plain;
}
__device__ __forceinline__ bool cell_is_plain(const CellData &c, u32 status, const DenseParams &p) {
    return cell_is_plain(c, status, p.level, p.eps2);
}
// eps2 of an axis (host): 2 * 16 ulp of the largest coordinate of the axis, relative to the cell size (see cell_is_plain)
static inline float plain_eps2(float amin, float asize, i64 res) {
    const double lo = amin, hi = (double) amin + (double) asize;
    const double big = fabs(lo) > fabs(hi) ? fabs(lo) : fabs(hi);
    const double cell = res > 1 ? fabs((double) asize) / (double) (res - 1) : 0.0;
    double e2 = cell > 0.0 ? 2.0 * 16.0 * big * 1.1920928955078125e-7 / cell : 1.0;
    if (!(e2 < 0.25)) e2 = 1.0;   // cells of a few ulp (or NaN boxes): every cell takes the exact path
    return (float) e2;
}
template <bool IMPLICIT = false>
__device__ __forceinline__ void load_cell(const float *__restrict__ values, const DenseParams &p, u32 r, u32 z, CellData &c) {
    load_cell_values<IMPLICIT>(values, p, r, z, c);
    load_cell_positions(p, r, z, c);
}

// position of the iso-crossing on cell edge e (exact reference arithmetic, see common.cuh)
__device__ __forceinline__ void cell_edge_point(const CellData &c, int e, float level, float &ox, float &oy, float &oz) {
    const int p0 = edge_c0(e), p1 = edge_c1(e);
    const float t = edge_t(c.v[p0], c.v[p1], level);
    ox = lerp_ref(t, c.px[(p0 >> 2) & 1], c.px[(p1 >> 2) & 1]);
    oy = lerp_ref(t, c.py[(p0 >> 1) & 1], c.py[(p1 >> 1) & 1]);
    oz = lerp_ref(t, c.pz[p0 & 1], c.pz[p1 & 1]);
}

// Owner entry + axis of each of the 12 edges of the cell at entry s.
//   lbY / lbX / lbXY: lower bounds of z in rows (x,y+1), (x+1,y), (x+1,y+1).
// slot id = 3*entry + axis (axis: 0 = +z, 1 = +y, 2 = +x).  Only meaningful for sign-change edges.
__device__ __forceinline__ void cell_edge_slots(const uint2 *__restrict__ entries, u32 s, u32 z, u32 lbY, u32 lbX, u32 lbXY,
                                                u32 n_entries, u32 slot[12]) {
    const u32 iY1 = lbY + ((lbY < n_entries && ent_z(entries[lbY].y) == z) ? 1u : 0u);
    const u32 iX1 = lbX + ((lbX < n_entries && ent_z(entries[lbX].y) == z) ? 1u : 0u);
    slot[0] = 3 * s + 0;        // e0  v0-v1  +z from corner 0
    slot[3] = 3 * s + 1;        // e3  v0-v2  +y from corner 0
    slot[8] = 3 * s + 2;        // e8  v0-v4  +x from corner 0
    slot[1] = 3 * (s + 1) + 1;  // e1  v1-v3  +y from corner 1 (x,y,z+1): the next entry of the row
    slot[9] = 3 * (s + 1) + 2;  // e9  v1-v5  +x from corner 1
    slot[2] = 3 * lbY + 0;      // e2  v2-v3  +z from corner 2 (x,y+1,z)
    slot[11] = 3 * lbY + 2;     // e11 v2-v6  +x from corner 2
    slot[10] = 3 * iY1 + 2;     // e10 v3-v7  +x from corner 3 (x,y+1,z+1)
    slot[4] = 3 * lbX + 0;      // e4  v4-v5  +z from corner 4 (x+1,y,z)
    slot[7] = 3 * lbX + 1;      // e7  v4-v6  +y from corner 4
    slot[5] = 3 * iX1 + 1;      // e5  v5-v7  +y from corner 5 (x+1,y,z+1)
    slot[6] = 3 * lbXY + 0;     // e6  v6-v7  +z from corner 6 (x+1,y+1,z)
}

// grid point positions, (X,Y,Z,3) f32 -- replaces get_vtx_pos_op (include/utils.cuh:62-80)
static __global__ void __launch_bounds__(256) k_grid_points(Geom g, float *__restrict__ out) {
    const i64 P = g.X * g.Y * g.Z;
    const u32 Y = (u32) g.Y, Z = (u32) g.Z;
    for (i64 n = (i64) blockIdx.x * blockDim.x + threadIdx.x; n < P; n += (i64) gridDim.x * blockDim.x) {
        u32 z = (u32) (n % Z);
        i64 r = n / Z;
        u32 y = (u32) (r % Y), x = (u32) (r / Y);
        float *o = out + 3 * n;
        o[0] = axis_pos(x + (u32) g.x_off, (u32) g.Xg - 1, g.amin[0], g.asize[0]);
        o[1] = axis_pos(y, Y - 1, g.amin[1], g.asize[1]);
        o[2] = axis_pos(z, Z - 1, g.amin[2], g.asize[2]);
    }
}

static inline u32 compact_heavy_cap(u32 cap) { return cap / FILL_HEAVY + 2; }
// Enqueue the compaction stage: bits (+ the span summary behind them) -> entries (ordered) + row_start + counters[C_S].
// desc must hold compact_desc_count(p) zeroed descriptors; counters zeroed; span_cnt: compact_span_bytes(p), not initialised.
static inline void launch_compact(const u32 *bits, const DenseParams &p, uint2 *entries, u32 cap, u32 *row_start, u64 *desc,
                                  u32 *counters, unsigned char *span_cnt, u32 *heavy_list, cudaStream_t stream) {
    if (compact128_ok(p)) {
        const u32 spr = (u32) (p.g.Z >> 7);
        const unsigned char *sum = span_sum_of(bits, p.P);
        const u32 heavy_cap = compact_heavy_cap(cap);
        if (spr <= 32u) {       // block-cooperative count / fill (thread per candidate span)
            constexpr u32 RF_ROWS = 128;
            const u32 rb = (p.R + RC_ROWS - 1u) / RC_ROWS, fb = (p.R + RF_ROWS - 1u) / RF_ROWS;
            if (spr % 8u == 0u) ISX_LAUNCH_PDL(k_rowcount_blk<8>, rb, RC_ROWS, 0, stream, bits, sum, p, row_start, span_cnt);
            else if (spr % 4u == 0u) ISX_LAUNCH_PDL(k_rowcount_blk<4>, rb, RC_ROWS, 0, stream, bits, sum, p, row_start, span_cnt);
            else ISX_LAUNCH_PDL(k_rowcount_blk<1>, rb, RC_ROWS, 0, stream, bits, sum, p, row_start, span_cnt);
            ISX_LAUNCH_PDL(k_scan_rows, scan_blocks(148), 256, 0, stream, row_start, (u32) p.R, desc, counters, (int) C_TICKET_A, (int) C_S);
            ISX_LAUNCH_PDL(k_rowfill_cnt, fb, RF_ROWS, 0, stream, bits, p, row_start, (const unsigned char *) span_cnt, entries, cap, heavy_list, heavy_cap, counters + C_NHEAVY);
        } else {                // very long rows (Z > 4096): thread per row
            const u32 rb = (p.R + 255u) / 256u;
            if (spr % 8u == 0u) ISX_LAUNCH(k_rowcount_sum<8>, rb, 256, 0, stream, bits, sum, p, row_start);
            else if (spr % 4u == 0u) ISX_LAUNCH(k_rowcount_sum<4>, rb, 256, 0, stream, bits, sum, p, row_start);
            else ISX_LAUNCH(k_rowcount_sum<1>, rb, 256, 0, stream, bits, sum, p, row_start);
            ISX_LAUNCH_PDL(k_scan_rows, scan_blocks(148), 256, 0, stream, row_start, (u32) p.R, desc, counters, (int) C_TICKET_A, (int) C_S);
            ISX_LAUNCH(k_rowfill128, (p.R + 127u) / 128u, 128, 0, stream, bits, p, row_start, sum, entries, cap, heavy_list, heavy_cap,
                       counters + C_NHEAVY);
        }
        ISX_LAUNCH_PDL(k_rowfill_heavy, 148 * 4, 256, 0, stream, bits, p, (const u32 *) row_start, entries, cap, (const u32 *) heavy_list, heavy_cap, (const u32 *) (counters + C_NHEAVY));
    } else {
        ISX_LAUNCH(k_compact, (p.NQ + CP_TILE - 1) / CP_TILE, 256, 0, stream, bits, p, entries, cap, row_start, desc, counters);
    }
}
static inline size_t compact_span_bytes(const DenseParams &p) { return compact128_ok(p) && (p.g.Z >> 7) <= 32 ? (size_t) p.R * (size_t) (p.g.Z >> 7) + 16 : 16; }
static inline size_t compact_desc_count(const DenseParams &p) {
    size_t a = (size_t) p.NQ / CP_TILE + 2, b = (size_t) p.R / RS_TILE + 2;
    return a > b ? a : b;
}

}   // namespace isx
