// setops.cu -- cell-list maintenance of SparseGrid and small index utilities, as kernels.
//
// Replaces the thrust calls of the reference's SparseGrid (src/grid/sparse.cu of the reference):
//   add_cells     :71-97   sort + unique of the new ids, set_union with the sorted list
//   remove_cells  :99-126  set_difference
//   filter_cell_indices :150-179  copy_if on the crossing flag
// and UniformGrid::get_cells (src/grid/uniform.cu:42-51, include/utils.cuh:32-60).
//
//   isoext_ids_sort_unique   ids (any order, duplicates) -> sorted unique list.  64-bit LSD radix sort (radix.cuh, the
//                            8 digit places of a 64-bit key), first-occurrence flags, look-back scan, ordered scatter
//   isoext_ids_difference    a (sorted unique) minus b (any order): b is sorted, every element of a binary-searches it
//   isoext_compact_flagged   stable compaction of 4- or 8-byte items by a byte flag
//   isoext_grid_cells_dense  8 corner point ids per cell
//   isoext_vertex_layer_histogram  welded vertices per cell layer (input of the slab balancer, dist.py)
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "dense.cuh"
#include "radix.cuh"

namespace isx {

struct SetWs {
    u32 *khi, *klo;      // n
    u32 *flag;           // n + 2  (becomes the exclusive scan)
    u32 *counters;       // 8
    u64 *desc;           // n / RS_TILE + 2
    RadixBuffers radix;
};
static size_t carve_set(Carver &c, size_t n, SetWs *out) {
    SetWs w;
    w.counters = c.take<u32>(8);                       // counters + desc are cleared by one memset
    w.desc = c.take<u64>(n / RS_TILE + 2);
    w.khi = c.take<u32>(n);
    w.klo = c.take<u32>(n);
    w.flag = c.take<u32>(n + 2);
    RadixBuffers::carve(c, n, &w.radix);
    if (out) *out = w;
    return c.bytes();
}

static __global__ void __launch_bounds__(256) k_ids_split(const i64 *__restrict__ ids, u32 n, u32 *__restrict__ khi, u32 *__restrict__ klo) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u64 v = (u64) ids[i];
        khi[i] = (u32) (v >> 32);
        klo[i] = (u32) v;
    }
}
// flag[i] = 1 iff the i-th id in sorted order differs from its predecessor
static __global__ void __launch_bounds__(256) k_ids_flag_first(const i64 *__restrict__ ids, const u32 *__restrict__ perm, u32 n,
                                                               u32 *__restrict__ flag) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        flag[i] = (i == 0 || ids[perm[i]] != ids[perm[i - 1]]) ? 1u : 0u;
}
// flag[i] = 1 iff a[i] does not occur in the sorted list b[perm[.]]
static __global__ void __launch_bounds__(256) k_ids_flag_absent(const i64 *__restrict__ a, u32 na, const i64 *__restrict__ b,
                                                                const u32 *__restrict__ perm, u32 nb, u32 *__restrict__ flag) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < na; i += gridDim.x * blockDim.x) {
        const u64 v = (u64) a[i];
        u32 lo = 0, hi = nb;
        while (lo < hi) {
            const u32 mid = (lo + hi) >> 1;
            if ((u64) b[perm[mid]] < v) lo = mid + 1; else hi = mid;
        }
        flag[i] = (lo < nb && (u64) b[perm[lo]] == v) ? 0u : 1u;
    }
}
static __global__ void __launch_bounds__(256) k_flag_from_bytes(const unsigned char *__restrict__ keep, u32 n, u32 *__restrict__ flag) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) flag[i] = keep[i] ? 1u : 0u;
}
// out[scan[i]] = src[perm ? perm[i] : i] for every i whose scan value steps up (scan = exclusive scan of the flags, n+1 entries)
template <typename T>
static __global__ void __launch_bounds__(256) k_scatter_flagged(const T *__restrict__ src, const u32 *__restrict__ perm, u32 n,
                                                                const u32 *__restrict__ scan, T *__restrict__ out) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const u32 o = scan[i];
        if (scan[i + 1] != o) out[o] = src[perm ? perm[i] : i];
    }
}

// (X-1, Y-1, Z-1, 8) corner ids, corner k at (x + (k>>2&1), y + (k>>1&1), z + (k&1))   include/utils.cuh:32-60
template <typename T>
static __global__ void __launch_bounds__(256) k_grid_cells(i64 X, i64 Y, i64 Z, T *__restrict__ out) {
    const i64 cy = Y - 1, cz = Z - 1, nc = (X - 1) * cy * cz;
    for (i64 t = (i64) blockIdx.x * blockDim.x + threadIdx.x; t < nc * 8; t += (i64) gridDim.x * blockDim.x) {
        const i64 c = t >> 3;
        const int k = (int) (t & 7);
        const i64 z = c % cz, r = c / cz, y = r % cy, x = r / cy;
        out[t] = (T) (((x + (k >> 2 & 1)) * Y + (y + (k >> 1 & 1))) * Z + (z + (k & 1)));
    }
}

// layer = clamp(floor((x - amin) / (amax - amin) * layers), 0, layers-1) in double (the formula of dist.vertex_layer_histogram)
static __global__ void __launch_bounds__(256) k_layer_hist(const float *__restrict__ V, i64 n, double amin, double asize, i64 layers,
                                                           u32 *__restrict__ hist) {
    for (i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64) gridDim.x * blockDim.x) {
        double t = floor(((double) V[3 * i] - amin) / asize * (double) layers);
        i64 l = t < 0.0 ? 0 : (t > (double) (layers - 1) ? layers - 1 : (i64) t);
        atomicAdd(&hist[l], 1u);
    }
}

int device_sms();
static int blocks_for(i64 n) {
    i64 want = (n + 255) / 256;
    const i64 cap = (i64) device_sms() * 16;
    return (int) (want < 1 ? 1 : (want > cap ? cap : want));
}

// flags in w.flag[0..n) -> exclusive scan in place (n+1 entries); returns the total through the counter block
static int scan_flags_and_count(const SetWs &w, u32 n, cudaStream_t stream, i64 *total) {
    ISX_LAUNCH(k_scan_rows, 148 * 4, 256, 0, stream, w.flag, n, w.desc, w.counters, 0, 1);
    ISX_CUDA(cudaGetLastError());
    u32 h[2] = {0, 0};
    ISX_CUDA(cudaMemcpyAsync(h, w.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    *total = h[1];
    return OK;
}
static int clear_set(const SetWs &w, size_t n, cudaStream_t stream) {
    ISX_CUDA(cudaMemsetAsync(w.counters, 0, (size_t) ((char *) (w.desc + n / RS_TILE + 2) - (char *) w.counters), stream));
    return OK;
}

}   // namespace isx

using namespace isx;

extern "C" {

size_t isoext_setops_workspace_bytes(int64_t n) {
    Carver c(nullptr);
    return carve_set(c, (size_t) (n > 0 ? n : 1), nullptr);
}

int isoext_ids_sort_unique(const int64_t *ids, int64_t n, int64_t *out, void *workspace, size_t workspace_bytes, void *stream_,
                           int64_t *n_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    *n_out = 0;
    if (n <= 0) return OK;
    if (n >= ((i64) 1 << 31)) return fail(E_INVALID, "too many ids for one call (>= 2^31)");
    Carver c(workspace);
    SetWs w;
    if (carve_set(c, (size_t) n, &w) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    int rc = clear_set(w, (size_t) n, stream);
    if (rc != OK) return rc;
    const int blocks = blocks_for(n);
    ISX_LAUNCH(k_ids_split, blocks, 256, 0, stream, ids, (u32) n, w.khi, w.klo);
    ISX_CUDA(radix_sort96(w.khi, w.klo, w.klo, (u32) n, w.radix, stream, nullptr, 4));
    ISX_LAUNCH(k_ids_flag_first, blocks, 256, 0, stream, ids, w.radix.perm[0], (u32) n, w.flag);
    i64 total = 0;
    rc = scan_flags_and_count(w, (u32) n, stream, &total);
    if (rc != OK) return rc;
    ISX_LAUNCH(k_scatter_flagged<i64>, blocks, 256, 0, stream, ids, w.radix.perm[0], (u32) n, w.flag, out);
    ISX_CUDA(cudaGetLastError());
    *n_out = total;
    return OK;
}

int isoext_ids_difference(const int64_t *a, int64_t na, const int64_t *b, int64_t nb, int64_t *out, void *workspace,
                          size_t workspace_bytes, void *stream_, int64_t *n_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    *n_out = 0;
    if (na <= 0) return OK;
    if (na >= ((i64) 1 << 31) || nb >= ((i64) 1 << 31)) return fail(E_INVALID, "too many ids for one call (>= 2^31)");
    const size_t m = (size_t) (na > nb ? na : nb);
    Carver c(workspace);
    SetWs w;
    if (carve_set(c, m, &w) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    int rc = clear_set(w, m, stream);
    if (rc != OK) return rc;
    if (nb > 0) {
        ISX_LAUNCH(k_ids_split, blocks_for(nb), 256, 0, stream, b, (u32) nb, w.khi, w.klo);
        ISX_CUDA(radix_sort96(w.khi, w.klo, w.klo, (u32) nb, w.radix, stream, nullptr, 4));
    }
    ISX_LAUNCH(k_ids_flag_absent, blocks_for(na), 256, 0, stream, a, (u32) na, b, w.radix.perm[0], (u32) (nb > 0 ? nb : 0), w.flag);
    i64 total = 0;
    rc = scan_flags_and_count(w, (u32) na, stream, &total);
    if (rc != OK) return rc;
    ISX_LAUNCH(k_scatter_flagged<i64>, blocks_for(na), 256, 0, stream, a, (const u32 *) nullptr, (u32) na, w.flag, out);
    ISX_CUDA(cudaGetLastError());
    *n_out = total;
    return OK;
}

int isoext_compact_flagged(const void *src, int elem_bytes, const unsigned char *keep, int64_t n, void *out, void *workspace,
                           size_t workspace_bytes, void *stream_, int64_t *n_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    *n_out = 0;
    if (n <= 0) return OK;
    if (elem_bytes != 4 && elem_bytes != 8) return fail(E_INVALID, "elem_bytes must be 4 or 8");
    if (n >= ((i64) 1 << 31)) return fail(E_INVALID, "too many items for one call (>= 2^31)");
    Carver c(workspace);
    SetWs w;
    if (carve_set(c, (size_t) n, &w) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    int rc = clear_set(w, (size_t) n, stream);
    if (rc != OK) return rc;
    const int blocks = blocks_for(n);
    ISX_LAUNCH(k_flag_from_bytes, blocks, 256, 0, stream, keep, (u32) n, w.flag);
    i64 total = 0;
    rc = scan_flags_and_count(w, (u32) n, stream, &total);
    if (rc != OK) return rc;
    if (elem_bytes == 4)
        ISX_LAUNCH(k_scatter_flagged<u32>, blocks, 256, 0, stream, (const u32 *) src, (const u32 *) nullptr, (u32) n, w.flag, (u32 *) out);
    else
        ISX_LAUNCH(k_scatter_flagged<u64>, blocks, 256, 0, stream, (const u64 *) src, (const u32 *) nullptr, (u32) n, w.flag, (u64 *) out);
    ISX_CUDA(cudaGetLastError());
    *n_out = total;
    return OK;
}

int isoext_grid_cells_dense(int64_t X, int64_t Y, int64_t Z, int wide, void *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (X < 1 || Y < 1 || Z < 1) return fail(E_INVALID, "grid shape must be positive");
    const i64 nc = (X - 1) * (Y - 1) * (Z - 1);
    if (nc <= 0) return OK;
    if (!wide && X * Y * Z > (i64) 0xffffffffLL + 1) return fail(E_INVALID, "point ids do not fit 32 bits: ask for 64-bit output");
    if (wide)
        ISX_LAUNCH(k_grid_cells<i64>, blocks_for(nc * 8), 256, 0, stream, X, Y, Z, (i64 *) out);
    else
        ISX_LAUNCH(k_grid_cells<u32>, blocks_for(nc * 8), 256, 0, stream, X, Y, Z, (u32 *) out);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

int isoext_vertex_layer_histogram(const float *V, int64_t n, float aabb_min_x, float aabb_max_x, int64_t layers, uint32_t *hist,
                                  void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (layers < 1) return fail(E_INVALID, "layers must be positive");
    ISX_CUDA(cudaMemsetAsync(hist, 0, (size_t) layers * sizeof(u32), stream));
    if (n > 0) ISX_LAUNCH(k_layer_hist, blocks_for(n), 256, 0, stream, V, n, (double) aabb_min_x, (double) aabb_max_x - (double) aabb_min_x, layers, hist);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

}   // extern "C"
