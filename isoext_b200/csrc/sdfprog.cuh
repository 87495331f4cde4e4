// sdfprog.cuh -- analytic signed-distance programs evaluated inside the extraction kernels.
//
// The reference's workflow evaluates `isoext.sdf` objects with torch on grid.get_points() (12 B/point written and read
// back), stores the field with set_values (4 B/point written) and the extraction reads it again (4 B/point).  For the
// built-in primitives and combinators (src/isoext/sdf.py:69-221 of the reference: Sphere, Torus, Cuboid, Union,
// SmoothUnion, Intersection, Negation, Translation, Rotation) the host layer compiles the object tree to a small
// postfix program instead and the kernels evaluate it where they would have loaded a value: the field is never
// written to HBM.  The volume pass becomes a compute pass with Lipschitz culling (k_sdf_bits, dense.cuh): all these
// operators are 1-Lipschitz, so one evaluation at the centre of a 32-point word decides the whole word unless the
// surface is within half a word.
//
// Arithmetic: float32 with explicit _rn intrinsics, written the way torch evaluates the same expressions
// (norm = sqrt of a left-to-right sum of squares, max / min / abs exact); tests/test_sdf_program_gpu.py measures the
// agreement with torch on the GPU.  The parity contract of the fused path is: same mesh, bit for bit, as materialising
// the field with isoext_sdf_eval_dense (the SAME device function) and extracting it in two steps.
#pragma once
#include "common.cuh"

namespace isx {

enum SdfOp : unsigned char {
    SDF_END = 0, SDF_RESET = 1, SDF_TRANSLATE = 2, SDF_ROTATE = 3, SDF_SPHERE = 4, SDF_TORUS = 5, SDF_CUBOID = 6,
    SDF_UNION = 7, SDF_INTER = 8, SDF_SMOOTH = 9, SDF_NEG = 10
};
constexpr int SDF_MAX_OPS = 240, SDF_MAX_CONSTS = 256, SDF_STACK = 12;

// Device-resident program (the host layer uploads exactly this layout, little endian).
struct SdfProg {
    u32 n_ops, n_consts;
    float lipschitz;          // bound of |f(a) - f(b)| / |a - b| (1 for every supported operator; a little slack is added)
    u32 reserved;
    unsigned char ops[SDF_MAX_OPS];
    float consts[SDF_MAX_CONSTS];
};

__device__ __forceinline__ float sdf_norm3(float a, float b, float c) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)));
}
__device__ __forceinline__ float sdf_norm2(float a, float b) { return __fsqrt_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b))); }

// f(px, py, pz)
static __device__ __noinline__ float sdf_eval(const SdfProg *__restrict__ prog, float px, float py, float pz) {
    float st[SDF_STACK];
    int sp = 0;
    float x = px, y = py, z = pz;
    const float *c = prog->consts;
    const u32 n = prog->n_ops;
    for (u32 i = 0; i < n;) {
        const unsigned char op = prog->ops[i++];
        switch (op) {
            case SDF_RESET: x = px; y = py; z = pz; break;
            case SDF_TRANSLATE: x = __fsub_rn(x, c[0]); y = __fsub_rn(y, c[1]); z = __fsub_rn(z, c[2]); c += 3; break;
            case SDF_ROTATE: {   // p @ R, R row-major: p'_j = p0 R0j + p1 R1j + p2 R2j
                const float nx = __fadd_rn(__fadd_rn(__fmul_rn(x, c[0]), __fmul_rn(y, c[3])), __fmul_rn(z, c[6]));
                const float ny = __fadd_rn(__fadd_rn(__fmul_rn(x, c[1]), __fmul_rn(y, c[4])), __fmul_rn(z, c[7]));
                const float nz = __fadd_rn(__fadd_rn(__fmul_rn(x, c[2]), __fmul_rn(y, c[5])), __fmul_rn(z, c[8]));
                x = nx; y = ny; z = nz; c += 9;
                break;
            }
            case SDF_SPHERE: st[sp++] = __fsub_rn(sdf_norm3(x, y, z), c[0]); c += 1; break;
            case SDF_TORUS: {
                const float ring = __fsub_rn(sdf_norm2(x, y), c[0]);
                st[sp++] = __fsub_rn(sdf_norm2(ring, z), c[1]);
                c += 2;
                break;
            }
            case SDF_CUBOID: {
                const float qx = __fsub_rn(fabsf(x), c[0]), qy = __fsub_rn(fabsf(y), c[1]), qz = __fsub_rn(fabsf(z), c[2]);
                const float inside = fmaxf(fmaxf(qx, qy), qz);
                const float outside = sdf_norm3(fmaxf(qx, 0.f), fmaxf(qy, 0.f), fmaxf(qz, 0.f));
                st[sp++] = __fadd_rn(outside, fminf(inside, 0.f));
                c += 3;
                break;
            }
            case SDF_UNION: case SDF_INTER: {
                const int k = prog->ops[i++];
                float a = st[sp - k];
                for (int j = 1; j < k; j++) a = op == SDF_UNION ? fminf(a, st[sp - k + j]) : fmaxf(a, st[sp - k + j]);
                sp -= k;
                st[sp++] = a;
                break;
            }
            case SDF_SMOOTH: {   // acc = -k log(exp(-acc/k) + exp(-other/k)), pairwise from the left
                const int k = prog->ops[i++];
                const float kk = c[0];
                c += 1;
                float a = st[sp - k];
                for (int j = 1; j < k; j++) {
                    const float o = st[sp - k + j];
                    a = __fmul_rn(-kk, logf(__fadd_rn(expf(__fdiv_rn(-a, kk)), expf(__fdiv_rn(-o, kk)))));
                }
                sp -= k;
                st[sp++] = a;
                break;
            }
            case SDF_NEG: st[sp - 1] = -st[sp - 1]; break;
            default: i = n; break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.f;
}

}   // namespace isx
