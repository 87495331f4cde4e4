// common.cuh -- shared device/host helpers for the sm_100a iso-surface extraction kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace isx {

typedef unsigned long long u64;
typedef uint32_t u32;
typedef int64_t i64;

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI returns int status + isoext_last_error())
// ---------------------------------------------------------------------------------------------
enum Status : int {
    OK = 0,
    E_INVALID = -1,     // bad argument
    E_CUDA = -2,        // CUDA runtime error
    E_WORKSPACE = -3,   // workspace / scratch buffer too small
    E_CAPACITY = -4,    // a capacity-bounded list overflowed; counts_out holds the needed sizes
    E_METHOD = -5,      // unknown marching-cubes method
};

std::string &last_error();
int fail(int code, const std::string &msg);

#define ISX_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return isx::fail(isx::E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
    } while (0)

// Kernel-launch accounting and optional per-launch timing of the dominant (volume-streaming) kernel.
// bench.py reads these through isoext_profile_begin/end (api.cu); the product path never depends on them.
extern long long g_kernel_launches;
extern int g_signbits_variant;
// development tuning knobs (isoext_debug_set_tuning): [0] blocks per SM of the chained-scan kernels (default 4)
extern int g_tuning[8];
static inline int scan_blocks(int sms) { return sms * (g_tuning[0] > 0 ? g_tuning[0] : 4); }
struct StreamTimer {
    bool enabled = false;
    static constexpr int kMaxPairs = 4096;
    cudaEvent_t ev[2 * kMaxPairs];
    int used = 0, created = 0;
};
extern StreamTimer g_stream_timer;
void stream_timer_mark(cudaStream_t s);   // record the next event of a (begin,end) pair if enabled

// optional per-kernel timing (development): every launch is bracketed by CUDA events
extern bool g_detail_timing;
void detail_mark(const char *name, cudaStream_t s, bool begin);

#define ISX_LAUNCH(kernel, grid, block, smem, stream, ...)                     \
    do {                                                                       \
        ++isx::g_kernel_launches;                                              \
        if (isx::g_detail_timing) isx::detail_mark(#kernel, (stream), true);   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
        if (isx::g_detail_timing) isx::detail_mark(#kernel, (stream), false);  \
    } while (0)

// Programmatic dependent launch (griddepcontrol, sm_90+): a kernel launched with ISX_LAUNCH_PDL may be scheduled while
// the previous kernel of the stream is still draining; its blocks park at pdl_wait() -- which returns once that kernel
// has completed and its writes are visible -- so the launch latency and the ramp-up (2-4 us per boundary, 12 boundaries
// per extraction) overlap the predecessor's tail.  Rules: pdl_wait() is the FIRST statement of every kernel that is
// ever launched this way (nothing produced by an earlier kernel may be touched before it); pdl_trigger() right behind it
// lets the NEXT kernel do the same.  Both are no-ops under an ordinary launch.  g_tuning[6] = 1 switches the launch
// attribute off (A/B).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define ISX_LAUNCH_PDL(kernel, grid, block, smem, stream, ...)                                     \
    do {                                                                                           \
        ++isx::g_kernel_launches;                                                                  \
        if (isx::g_detail_timing) isx::detail_mark(#kernel, (stream), true);                       \
        cudaLaunchConfig_t cfg_ = {};                                                              \
        cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block);                                    \
        cfg_.dynamicSmemBytes = (smem); cfg_.stream = (stream);                                    \
        cudaLaunchAttribute at_[1];                                                                \
        at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
        at_[0].val.programmaticStreamSerializationAllowed = 1;                                     \
        cfg_.attrs = at_; cfg_.numAttrs = isx::g_tuning[6] == 1 ? 0 : 1;                           \
        cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                            \
        if (isx::g_detail_timing) isx::detail_mark(#kernel, (stream), false);                      \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves typed sub-buffers out of one caller-provided device blob (256-byte aligned pieces).
struct Carver {
    char *base;
    size_t off = 0;
    explicit Carver(void *p) : base(static_cast<char *>(p)) {}
    template <typename T> T *take(size_t n) {
        off = align_up(off, 256);
        T *r = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return r;
    }
    size_t bytes() const { return align_up(off, 256); }
};

// ---------------------------------------------------------------------------------------------
// grid geometry: local slab of a (possibly larger) global grid
// ---------------------------------------------------------------------------------------------
struct Geom {
    i64 X, Y, Z;          // points per axis of the local slab
    i64 x_off;            // global x index of local plane 0
    i64 Xg;               // global number of points along x
    float amin[3], asize[3];
};

// Reference semantics (include/utils.cuh:62-80 of the reference, SASS-checked):
//   pos = fma( div.rn(float(i), float(res-1)), aabb_max-aabb_min, aabb_min )
__device__ __forceinline__ float axis_pos(u32 i, u32 res_minus_1, float amin, float asize) {
    return __fmaf_rn(__fdiv_rn((float) i, (float) res_minus_1), asize, amin);
}

// Reference semantics (src/mc/nagae.cu:54-56, include/math.cuh:77-81; SASS: FADD 1-t, FMUL t*b, FFMA):
//   t = d != 0 ? div.rn(level - v0, v1 - v0) : 0 ;  p = fma(a, 1 - t, rn(t * b))
__device__ __forceinline__ float edge_t(float v0, float v1, float level) {
    float d = __fsub_rn(v1, v0);
    return (d != 0.0f) ? __fdiv_rn(__fsub_rn(level, v0), d) : 0.0f;
}
__device__ __forceinline__ float lerp_ref(float t, float a, float b) {
    return __fmaf_rn(a, __fsub_rn(1.0f, t), __fmul_rn(t, b));
}

// Order-preserving float -> u32 key; -0 and +0 share a key (IEEE == treats them as equal, and the
// reference welds with float comparisons: include/math.cuh:112-126).
__device__ __forceinline__ u32 float_key(float f) {
    u32 b = __float_as_uint(f);
    if ((b << 1) == 0u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(u32 k) {
    u32 b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

// ---------------------------------------------------------------------------------------------
// streaming loads
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

// ---------------------------------------------------------------------------------------------
// decoupled look-back ("chained") exclusive scan over thread blocks.
//
// Blocks take a ticket (dynamic block id) so that block b only ever waits on blocks that have
// already started.  Descriptor word: [63:62] state (0 empty, 1 aggregate, 2 inclusive prefix),
// [61:32] epoch tag, [31:0] value.  A descriptor only counts if its epoch matches, which lets the
// radix sort reuse one descriptor array across passes without re-zeroing it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 desc_pack(u32 state, u32 epoch, u32 value) {
    return ((u64) state << 62) | ((u64) (epoch & 0x3fffffffu) << 32) | value;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Called by ONE warp (all 32 lanes).  `tile` = this block's ticket, `aggregate` = its total.
// Publishes the aggregate, walks back over predecessors and returns the exclusive prefix
// (sum of aggregates of tiles < tile); then publishes the inclusive prefix.
// `desc` is indexed desc[tile * stride].
__device__ __forceinline__ u32 lookback_exclusive(u64 *desc, u32 stride, u32 tile, u32 aggregate, u32 epoch) {
    const u32 lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(desc, desc_pack(2, epoch, aggregate));
        return 0;
    }
    if (lane == 0) st_relaxed_u64(desc + (size_t) tile * stride, desc_pack(1, epoch, aggregate));
    u32 exclusive = 0;
    int look = (int) tile - 1;   // lane 0 inspects tile `look`, lane k inspects look - k
    while (true) {
        int idx = look - (int) lane;
        u64 d = 0;
        u32 state;
        // spin until every inspected predecessor has published something for this epoch
        do {
            if (idx >= 0) {
                d = ld_relaxed_u64(desc + (size_t) idx * stride);
                state = (((u32) (d >> 32)) & 0x3fffffffu) == (epoch & 0x3fffffffu) ? (u32) (d >> 62) : 0u;
            } else {
                state = 2;   // before the first tile: inclusive prefix 0
                d = 0;
            }
        } while (__any_sync(0xffffffffu, state == 0));
        u32 incl_mask = __ballot_sync(0xffffffffu, state == 2);
        u32 v = (u32) d;
        if (incl_mask) {
            int first = __ffs(incl_mask) - 1;   // nearest predecessor with an inclusive prefix
            u32 contrib = (lane <= (u32) first) ? v : 0u;
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            exclusive += contrib;
            break;
        }
        u32 contrib = v;
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        exclusive += contrib;
        look -= 32;
    }
    if (lane == 0) st_relaxed_u64(desc + (size_t) tile * stride, desc_pack(2, epoch, exclusive + aggregate));
    return exclusive;
}

// Block-wide exclusive scan of one u32 per thread (blockDim.x <= 1024, multiple of 32).
// Returns the thread's exclusive prefix; *total receives the block aggregate (all threads).
__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *total, u32 *smem_warp /* >= 33 entries */) {
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    u32 incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (u32) o) incl += t;
    }
    if (lane == 31) smem_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nwarps ? smem_warp[lane] : 0u;
        u32 wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (u32) o) wi += t;
        }
        if (lane < nwarps) smem_warp[lane] = wi - w;
        if (lane == 31) smem_warp[32] = wi;
    }
    __syncthreads();
    u32 res = incl - v + smem_warp[warp];
    *total = smem_warp[32];
    __syncthreads();
    return res;
}

}   // namespace isx
