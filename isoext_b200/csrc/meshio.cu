// meshio.cu -- mesh output (host code; no kernel).  SURVEY.md 8f-3.
//
// Replaces the Python loop of the reference's write_obj (src/isoext/utils.py:42-63: `v.tolist()`, one f-string per
// vertex / face, `writelines`), which takes minutes on a multi-million-vertex mesh.  The bytes written are IDENTICAL
// to the reference's: Python prints a float32 coordinate as repr(float(x)) -- the shortest decimal string that
// round-trips the DOUBLE value -- in fixed notation unless the decimal exponent is < -4 or >= 16
// (CPython: PyOS_double_to_string(x, 'r', 0, Py_DTSF_ADD_DOT_0)).  std::to_chars gives the same shortest digits;
// the notation rule is restated below.  Formatting runs on all host cores, the file is written in one pass.
#include "../../include/isoext_b200.h"   // the C-ABI prototypes are compiler-checked against the definitions
#include "common.cuh"

#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// repr(float) of CPython for a double
inline void py_float_repr(double x, std::string &out) {
    if (std::isnan(x)) { out += "nan"; return; }
    if (std::isinf(x)) { out += x < 0 ? "-inf" : "inf"; return; }
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // shortest round-trip digits
    *r.ptr = 0;                                              // to_chars does not terminate the string
    const char *p = buf, *end = r.ptr;
    if (*p == '-') { out += '-'; p++; }
    const char *e = p;
    while (e < end && *e != 'e') e++;
    char digits[32];
    int nd = 0;
    for (const char *q = p; q < e; q++)
        if (*q != '.') digits[nd++] = *q;
    const int exp10 = std::atoi(e + 1);
    while (nd > 1 && digits[nd - 1] == '0') nd--;          // (to_chars never pads, but stay safe)
    const int decpt = exp10 + 1;                            // value = 0.d1d2... * 10^decpt
    if (decpt <= -4 || decpt > 16) {                        // exponent notation: d[.ddd]e+XX (at least two exponent digits)
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        const int ex = decpt - 1;
        char eb[8];
        std::snprintf(eb, sizeof(eb), "e%c%02d", ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
        out += eb;
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t) -decpt, '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t) (decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

template <typename Fn> void parallel_chunks(int64_t n, std::vector<std::string> &parts, Fn fn) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int) (hw ? (hw > 32 ? 32 : hw) : 4);
    if (n < 1 << 14) nt = 1;
    parts.assign(nt, std::string());
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const int64_t a = n * t / nt, b = n * (t + 1) / nt;
        th.emplace_back([&, t, a, b]() { fn(a, b, parts[t]); });
    }
    for (auto &x : th) x.join();
}

}   // namespace

extern "C" {

// v: nv x 3 float32, f: nf x 3 int32 (zero-based), both in HOST memory.  An empty mesh leaves an empty file.
int isoext_write_obj(const char *path, const float *v, int64_t nv, const int32_t *f, int64_t nf) {
    FILE *fp = std::fopen(path, "wb");
    if (!fp) return isx::fail(isx::E_INVALID, std::string("cannot open ") + path);
    if (nv <= 0 || nf <= 0 || !v || !f) { std::fclose(fp); return isx::OK; }
    std::vector<std::string> parts;
    parallel_chunks(nv, parts, [&](int64_t a, int64_t b, std::string &o) {
        o.reserve((size_t) (b - a) * 48);
        for (int64_t i = a; i < b; i++) {
            o += "v ";
            py_float_repr((double) v[3 * i], o); o += ' ';
            py_float_repr((double) v[3 * i + 1], o); o += ' ';
            py_float_repr((double) v[3 * i + 2], o); o += '\n';
        }
    });
    bool ok = true;
    for (auto &s : parts) ok = ok && std::fwrite(s.data(), 1, s.size(), fp) == s.size();
    parallel_chunks(nf, parts, [&](int64_t a, int64_t b, std::string &o) {
        o.reserve((size_t) (b - a) * 28);
        char buf[16];
        for (int64_t i = a; i < b; i++) {
            o += 'f';
            for (int k = 0; k < 3; k++) {
                o += ' ';
                auto r = std::to_chars(buf, buf + sizeof(buf), (long long) f[3 * i + k] + 1);   // OBJ ids are one-based
                o.append(buf, r.ptr - buf);
            }
            o += '\n';
        }
    });
    for (auto &s : parts) ok = ok && std::fwrite(s.data(), 1, s.size(), fp) == s.size();
    ok = (std::fclose(fp) == 0) && ok;
    return ok ? isx::OK : isx::fail(isx::E_INVALID, std::string("short write to ") + path);
}

// Binary little-endian PLY (extension: the compact format for multi-million-vertex meshes).
int isoext_write_ply(const char *path, const float *v, int64_t nv, const int32_t *f, int64_t nf) {
    FILE *fp = std::fopen(path, "wb");
    if (!fp) return isx::fail(isx::E_INVALID, std::string("cannot open ") + path);
    std::fprintf(fp, "ply\nformat binary_little_endian 1.0\ncomment isoext_b200\nelement vertex %lld\nproperty float x\nproperty float y\n"
                     "property float z\nelement face %lld\nproperty list uchar int vertex_indices\nend_header\n",
                 (long long) (nv > 0 ? nv : 0), (long long) (nf > 0 ? nf : 0));
    bool ok = true;
    if (nv > 0) ok = std::fwrite(v, 12, (size_t) nv, fp) == (size_t) nv;
    if (nf > 0) {
        std::vector<unsigned char> rec((size_t) nf * 13);
        for (int64_t i = 0; i < nf; i++) {
            rec[13 * i] = 3;
            std::memcpy(&rec[13 * i + 1], f + 3 * i, 12);
        }
        ok = ok && std::fwrite(rec.data(), 13, (size_t) nf, fp) == (size_t) nf;
    }
    ok = (std::fclose(fp) == 0) && ok;
    return ok ? isx::OK : isx::fail(isx::E_INVALID, std::string("short write to ") + path);
}

}   // extern "C"
