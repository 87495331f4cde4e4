// api.cu -- error plumbing and library identification of the C-ABI (include/isoext_b200.h).
#include "common.cuh"

#include <cmath>

namespace isx {
std::string &last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const std::string &msg) {
    last_error() = msg;
    return code;
}
long long g_kernel_launches = 0;
int g_signbits_variant = 0;
StreamTimer g_stream_timer;
void stream_timer_mark(cudaStream_t s) {
    StreamTimer &t = g_stream_timer;
    if (!t.enabled || t.used >= 2 * StreamTimer::kMaxPairs) return;
    if (t.used >= t.created) {
        if (cudaEventCreate(&t.ev[t.created]) != cudaSuccess) return;
        t.created++;
    }
    cudaEventRecord(t.ev[t.used++], s);
}
}   // namespace isx

extern "C" {
// tuning knob for the streaming kernel: low byte = variant, next byte = blocks per SM (0 = default)
int isoext_debug_set_signbits_variant(int v) { isx::g_signbits_variant = v; return 0; }
// ---- measurement hooks (bench.py): count kernel launches, time the volume-streaming kernel --------
int isoext_profile_begin(void) {
    isx::g_stream_timer.enabled = true;
    isx::g_stream_timer.used = 0;
    isx::g_kernel_launches = 0;
    return 0;
}
// stream_ms_total: sum of CUDA-event durations of the dominant kernel's launches since begin();
// stream_launches: how many of them; kernel_launches: all kernel launches of this library since begin().
int isoext_profile_end(double *stream_ms_total, int64_t *stream_launches, int64_t *kernel_launches) {
    isx::StreamTimer &t = isx::g_stream_timer;
    t.enabled = false;
    double total = 0;
    int pairs = t.used / 2;
    for (int i = 0; i < pairs; i++) {
        float ms = 0;
        if (cudaEventSynchronize(t.ev[2 * i + 1]) != cudaSuccess) return isx::fail(isx::E_CUDA, "event sync failed");
        if (cudaEventElapsedTime(&ms, t.ev[2 * i], t.ev[2 * i + 1]) != cudaSuccess) return isx::fail(isx::E_CUDA, "event elapsed failed");
        total += ms;
    }
    *stream_ms_total = total;
    *stream_launches = pairs;
    *kernel_launches = isx::g_kernel_launches;
    return 0;
}

const char *isoext_last_error(void) { return isx::last_error().c_str(); }
// Grid-plane position along one axis, computed on the host with the same operations as the device
// (float division, one fused multiply-add): used for the slab ownership thresholds.
float isoext_axis_position(int64_t i, int64_t res, float amin, float amax) {
    float q = (float) (uint32_t) i / (float) (uint32_t) (res - 1);
    return fmaf(q, amax - amin, amin);
}
const char *isoext_build_info(void) { return "isoext_b200 sm_100a " __DATE__; }
int isoext_abi_version(void) { return 1; }
}
