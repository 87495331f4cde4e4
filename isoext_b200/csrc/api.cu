// api.cu -- error plumbing and library identification of the C-ABI (include/isoext_b200.h).
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "common.cuh"

#include <cmath>
#include <map>
#include <vector>

namespace isx {
std::string &last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const std::string &msg) {
    last_error() = msg;
    return code;
}
bool g_detail_timing = false;
struct DetailRec { const char *name; cudaEvent_t e0, e1; };
static std::vector<DetailRec> g_detail;
void detail_mark(const char *name, cudaStream_t s, bool begin) {
    if (begin) {
        DetailRec r{name, nullptr, nullptr};
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, s);
        g_detail.push_back(r);
    } else if (!g_detail.empty()) {
        cudaEventRecord(g_detail.back().e1, s);
    }
}
long long g_kernel_launches = 0;
int g_signbits_variant = 0;
int g_tuning[8] = {0, 0, 0, 0, 0, 0, 0, 0};
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
// development tuning knobs (common.cuh: g_tuning); 0 restores the default
int isoext_debug_set_tuning(int key, int value) {
    if (key < 0 || key >= 8) return -1;
    isx::g_tuning[key] = value;
    return 0;
}
// Development: per-kernel CUDA-event timing.  enable(1) starts collecting; report() synchronises, writes
// "name total_us launches" lines into buf and clears the records.
int isoext_debug_detail_enable(int on) { isx::g_detail_timing = on != 0; return 0; }
// timeline variant: one line per launch "name start_us end_us" relative to the first recorded event
int isoext_debug_detail_timeline(char *buf, int buf_size) {
    cudaDeviceSynchronize();
    std::string out;
    if (!isx::g_detail.empty()) {
        cudaEvent_t first = isx::g_detail.front().e0;
        for (auto &r : isx::g_detail) {
            float t0 = 0, t1 = 0;
            if (r.e1 && cudaEventElapsedTime(&t0, first, r.e0) == cudaSuccess && cudaEventElapsedTime(&t1, first, r.e1) == cudaSuccess) {
                char line[256];
                snprintf(line, sizeof(line), "%-28s %10.1f %10.1f\n", r.name, t0 * 1e3, t1 * 1e3);
                out += line;
            }
        }
        for (auto &r : isx::g_detail) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        isx::g_detail.clear();
    }
    snprintf(buf, buf_size, "%s", out.c_str());
    return 0;
}
int isoext_debug_detail_report(char *buf, int buf_size) {
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<double, int>> acc;
    std::vector<std::string> order;
    for (auto &r : isx::g_detail) {
        float ms = 0;
        if (r.e1 && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            if (!acc.count(r.name)) order.push_back(r.name);
            acc[r.name].first += ms * 1e3;
            acc[r.name].second += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    isx::g_detail.clear();
    std::string out;
    for (auto &n : order) {
        char line[256];
        snprintf(line, sizeof(line), "%-28s %10.1f us total %6d launches %8.2f us each\n", n.c_str(), acc[n].first, acc[n].second,
                 acc[n].first / acc[n].second);
        out += line;
    }
    snprintf(buf, buf_size, "%s", out.c_str());
    return 0;
}
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
