// peer.cu -- the exchange steps of the slab-sharded path as kernels over NVLink / NVSwitch peer memory.
//
// No counterpart in the reference (it is single-GPU).  One process per GPU; every rank exports two
// allocations through CUDA IPC: its extended value slab and a small SYNC block.  Neighbours map them and
//   * k_peer_pull      waits (acquire, system scope) until the neighbour has published the epoch of this
//                      exchange and then copies the halo planes straight out of the neighbour's slab with
//                      128-bit loads -- wait and transfer are one kernel, there is no host rendezvous and no
//                      staging copy;
//   * k_relabel_peer   turns local vertex ids into global ids: every block reads the owned-vertex counts
//                      of the lower ranks directly from their SYNC blocks (waiting for the epoch tag), sums
//                      them and relabels -- the "all-gather + exclusive scan + relabel" step as one kernel.
// SYNC block (u64 words):  [0] READY epoch (values of that exchange are in place)
//                          [1] DONE epoch  (this rank has finished pulling its halos of that exchange)
//                          [2 + 4*s ..]  ring of PEER_RING count slots {epoch, n_own, n_tri, -}
// Ranks may drift apart by at most world-1 epochs (a pull needs the neighbour's READY of the same epoch),
// so a ring of 32 slots serves any single-node world size.  Every spin is bounded by a wall-clock timeout;
// on expiry the kernel raises *err (host-mapped memory) and stops waiting, and the host layer throws.
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "common.cuh"

namespace isx {

constexpr int PEER_RING = 32;
constexpr int PEER_WORDS = 2 + 4 * PEER_RING;
constexpr u64 PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;   // 20 s

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64 *p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= want (exact: == want); false on timeout
__device__ __forceinline__ bool wait_epoch(const u64 *flag, u64 want, bool exact) {
    const u64 t0 = global_ns();
    while (true) {
        const u64 v = ld_acquire_sys(flag);
        if (exact ? v == want : v >= want) return true;
        if (global_ns() - t0 > PEER_TIMEOUT_NS) return false;
        __nanosleep(200);
    }
}

__global__ void k_peer_publish(u64 *flag, u64 value) {
    __threadfence_system();   // everything this stream wrote before is visible to the peers first
    st_release_sys(flag, value);
}

__global__ void k_peer_publish_counts(u64 *sync, u64 epoch, u64 n_own, u64 n_tri) {
    u64 *slot = sync + 2 + 4 * (epoch % PEER_RING);
    slot[1] = n_own;
    slot[2] = n_tri;
    __threadfence_system();
    st_release_sys(slot, epoch);
}

__global__ void k_peer_wait(const u64 *flag_a, const u64 *flag_b, u64 want, u32 *err) {
    bool ok = true;
    if (flag_a) ok = wait_epoch(flag_a, want, false) && ok;
    if (flag_b) ok = wait_epoch(flag_b, want, false) && ok;
    if (!ok) *err = 1u;
}

struct PullSeg {
    float *dst;
    const float *src;        // peer memory
    i64 n;                   // floats
    const u64 *ready;        // READY word of the peer that owns src
};

__global__ void __launch_bounds__(256) k_peer_pull(PullSeg a, PullSeg b, u64 epoch, u32 *err) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        bool ok = true;
        if (a.n > 0) ok = wait_epoch(a.ready, epoch, false) && ok;
        if (b.n > 0) ok = wait_epoch(b.ready, epoch, false) && ok;
        if (!ok) *err = 1u;
        s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;
    const i64 tid = (i64) blockIdx.x * blockDim.x + threadIdx.x, nth = (i64) gridDim.x * blockDim.x;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const PullSeg s = k ? b : a;
        if (s.n <= 0) continue;
        const bool vec = ((reinterpret_cast<uintptr_t>(s.dst) | reinterpret_cast<uintptr_t>(s.src)) & 15u) == 0;
        const i64 n4 = vec ? s.n / 4 : 0;
        const float4 *src4 = reinterpret_cast<const float4 *>(s.src);
        float4 *dst4 = reinterpret_cast<float4 *>(s.dst);
        i64 i = tid;
        for (; i + 3 * nth < n4; i += 4 * nth) {   // 4 independent 128-bit loads in flight per thread (NVLink latency)
            const float4 v0 = __ldcg(src4 + i), v1 = __ldcg(src4 + i + nth), v2 = __ldcg(src4 + i + 2 * nth),
                         v3 = __ldcg(src4 + i + 3 * nth);
            dst4[i] = v0; dst4[i + nth] = v1; dst4[i + 2 * nth] = v2; dst4[i + 3 * nth] = v3;
        }
        for (; i < n4; i += nth) dst4[i] = __ldcg(src4 + i);
        for (i64 j = 4 * n4 + tid; j < s.n; j += nth) s.dst[j] = __ldcg(s.src + j);
    }
}

struct PeerSyncs {
    const u64 *p[PEER_RING];
};

// Local (extended-slab) vertex ids -> global ids (see k_relabel_faces in mc_dense.cu for the map):
// base_mine = sum of n_own over the lower ranks, read from their SYNC blocks.
__global__ void __launch_bounds__(256) k_relabel_peer(int *F, i64 n, i64 n_lo, i64 n_hi, i64 n_own, PeerSyncs peers, int rank, u64 epoch,
                                                      i64 *bases_out, u32 *err) {
    __shared__ unsigned long long s_vb, s_fb;
    __shared__ int s_ok;
    if (threadIdx.x == 0) { s_vb = 0; s_fb = 0; s_ok = 1; }
    __syncthreads();
    if ((int) threadIdx.x < rank) {
        const u64 *slot = peers.p[threadIdx.x] + 2 + 4 * (epoch % PEER_RING);
        if (wait_epoch(slot, epoch, true)) {
            atomicAdd(&s_vb, (unsigned long long) ld_acquire_sys(slot + 1));
            atomicAdd(&s_fb, (unsigned long long) ld_acquire_sys(slot + 2));
        } else {
            s_ok = 0;
            *err = 1u;
        }
    }
    __syncthreads();
    if (!s_ok) return;
    const i64 base_mine = (i64) s_vb, base_next = base_mine + n_own;
    if (bases_out && blockIdx.x == 0 && threadIdx.x == 0) { bases_out[0] = base_mine; bases_out[1] = (i64) s_fb; }
    for (i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64) gridDim.x * blockDim.x) {
        const i64 id = F[i];
        const i64 g = id < n_lo ? base_mine - (n_lo - id) : (id < n_hi ? base_mine + (id - n_lo) : base_next + (id - n_hi));
        F[i] = (int) g;
    }
}

}   // namespace isx

using namespace isx;

extern "C" {

int isoext_peer_sync_words(void) { return PEER_WORDS; }

int isoext_peer_alloc(size_t bytes, void **d_ptr, unsigned char *handle64) {
    if (!d_ptr || !handle64 || bytes == 0) return fail(E_INVALID, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    ISX_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *d_ptr = p;
    return OK;
}
int isoext_peer_free(void *d_ptr) {
    if (d_ptr) ISX_CUDA(cudaFree(d_ptr));
    return OK;
}
int isoext_peer_open(const unsigned char *handle64, void **d_ptr) {
    if (!d_ptr || !handle64) return fail(E_INVALID, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void *p = nullptr;
    ISX_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return OK;
}
int isoext_peer_close(void *d_ptr) {
    if (d_ptr) ISX_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return OK;
}

int isoext_peer_publish(uint64_t *d_flag, uint64_t value, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ISX_LAUNCH(k_peer_publish, 1, 1, 0, stream, (u64 *) d_flag, (u64) value);
    ISX_CUDA(cudaGetLastError());
    return OK;
}
int isoext_peer_publish_counts(uint64_t *d_sync, uint64_t epoch, int64_t n_own, int64_t n_tri, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    ISX_LAUNCH(k_peer_publish_counts, 1, 1, 0, stream, (u64 *) d_sync, (u64) epoch, (u64) n_own, (u64) n_tri);
    ISX_CUDA(cudaGetLastError());
    return OK;
}
int isoext_peer_wait(const uint64_t *d_flag_a, const uint64_t *d_flag_b, uint64_t want, uint32_t *err_mapped, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d_flag_a && !d_flag_b) return OK;
    ISX_LAUNCH(k_peer_wait, 1, 1, 0, stream, (const u64 *) d_flag_a, (const u64 *) d_flag_b, (u64) want, err_mapped);
    ISX_CUDA(cudaGetLastError());
    return OK;
}
int isoext_peer_halo_pull(float *d_dst0, const float *peer_src0, int64_t n0, const uint64_t *peer_ready0, float *d_dst1,
                          const float *peer_src1, int64_t n1, const uint64_t *peer_ready1, uint64_t epoch, uint32_t *err_mapped,
                          void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n0 <= 0 && n1 <= 0) return OK;
    PullSeg a{d_dst0, peer_src0, n0 > 0 ? n0 : 0, (const u64 *) peer_ready0}, b{d_dst1, peer_src1, n1 > 0 ? n1 : 0, (const u64 *) peer_ready1};
    if ((a.n > 0 && (!a.dst || !a.src || !a.ready)) || (b.n > 0 && (!b.dst || !b.src || !b.ready)))
        return fail(E_INVALID, "halo_pull: null pointer");
    const i64 total4 = (a.n + b.n) / 4 + 1;
    i64 want = (total4 + 256 * 4 - 1) / (256 * 4);
    const int blocks = (int) (want > 148 * 4 ? 148 * 4 : (want < 1 ? 1 : want));
    ISX_LAUNCH(k_peer_pull, blocks, 256, 0, stream, a, b, (u64) epoch, err_mapped);
    ISX_CUDA(cudaGetLastError());
    return OK;
}
int isoext_relabel_faces_peer(int32_t *d_F, int64_t n_ids, int64_t n_lo, int64_t n_hi, int64_t n_own, const uint64_t *const *peer_syncs,
                              int rank, uint64_t epoch, int64_t *d_bases_out, uint32_t *err_mapped, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rank < 0 || rank > PEER_RING) return fail(E_INVALID, "relabel_faces_peer: rank out of range (at most 32 ranks)");
    if (n_ids <= 0 && !d_bases_out) return OK;
    PeerSyncs ps;
    for (int r = 0; r < PEER_RING; r++) ps.p[r] = r < rank ? (const u64 *) peer_syncs[r] : nullptr;
    i64 want = (n_ids + 255) / 256;
    const int blocks = (int) (want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
    ISX_LAUNCH(k_relabel_peer, blocks, 256, 0, stream, d_F, n_ids > 0 ? n_ids : 0, n_lo, n_hi, n_own, ps, rank, (u64) epoch, d_bases_out,
               err_mapped);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

}   // extern "C"
