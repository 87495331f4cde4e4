// smooth.cu -- separable Gaussian smoothing of a dense scalar field.
//
// Replaces the dense k^3 conv3d of the reference's gaussian_smooth (src/isoext/utils.py:5-39: F.pad(replicate) +
// F.conv3d with the outer product of the 1-D taps): the kernel is separable, so three 1-D passes (z, y, x) with
// clamped indices (= replicate padding) do k*3 instead of k^3 multiply-adds per voxel and never materialise the
// padded volume.  Every pass streams the volume once (the k taps of a voxel hit L1/L2: along z they are
// neighbours in a row, along y / x neighbouring threads read neighbouring z), so the filter is HBM-bound:
// 3 x (4 B read + 4 B written) per voxel.  Same 1-D taps as the reference (computed by the host layer with the
// reference's expressions); the summation order differs from cuDNN's, so results agree to float32 rounding
// (tests: 1e-6 of the field's range), not bit for bit.
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "common.cuh"

namespace isx {

constexpr int SMOOTH_MAX_TAPS = 127;
struct Taps { float w[SMOOTH_MAX_TAPS + 1]; };

// out[i] = sum_t w[t] * in[clamp(c + t - r) along the axis], c = coordinate of i along the axis (n points, stride `stride`)
__global__ void __launch_bounds__(256) k_smooth_axis(const float *__restrict__ in, float *__restrict__ out, i64 total, i64 n, i64 stride,
                                                     int k, Taps taps) {
    const int r = k >> 1;
    for (i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64) gridDim.x * blockDim.x) {
        const i64 c = (i / stride) % n;
        const float *base = in + (i - c * stride);
        float acc = 0.f;
        if (c >= r && c + r < n) {           // interior: no clamping
            const float *q = base + (c - r) * stride;
            for (int t = 0; t < k; t++) acc = __fmaf_rn(taps.w[t], __ldg(q + (i64) t * stride), acc);
        } else {
            for (int t = 0; t < k; t++) {
                i64 j = c + t - r;
                j = j < 0 ? 0 : (j >= n ? n - 1 : j);
                acc = __fmaf_rn(taps.w[t], __ldg(base + j * stride), acc);
            }
        }
        out[i] = acc;
    }
}

int device_sms();

}   // namespace isx

using namespace isx;

extern "C" {

// field (X,Y,Z) f32 -> out (X,Y,Z) f32; tmp: X*Y*Z floats of scratch; taps: k host floats (k odd, <= 127).
// in / out / tmp must be three different buffers.
int isoext_gaussian_smooth_separable(const float *field, int64_t X, int64_t Y, int64_t Z, const float *taps_host, int k, float *tmp,
                                     float *out, void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (X < 1 || Y < 1 || Z < 1) return fail(E_INVALID, "field shape must be positive");
    if (k < 1 || k > SMOOTH_MAX_TAPS || !(k & 1)) return fail(E_INVALID, "kernel size must be odd and <= 127");
    if (field == out || field == tmp || tmp == out) return fail(E_INVALID, "field, tmp and out must be distinct buffers");
    Taps t;
    for (int i = 0; i < k; i++) t.w[i] = taps_host[i];
    const i64 total = X * Y * Z;
    i64 want = (total + 255) / 256;
    const i64 cap = (i64) device_sms() * 32;
    const int blocks = (int) (want > cap ? cap : want);
    ISX_LAUNCH(k_smooth_axis, blocks, 256, 0, stream, field, out, total, Z, (i64) 1, k, t);       // along z
    ISX_LAUNCH(k_smooth_axis, blocks, 256, 0, stream, out, tmp, total, Y, Z, k, t);                // along y
    ISX_LAUNCH(k_smooth_axis, blocks, 256, 0, stream, tmp, out, total, X, Y * Z, k, t);            // along x
    ISX_CUDA(cudaGetLastError());
    return OK;
}

}   // extern "C"
