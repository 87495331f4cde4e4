// dcmath.cuh -- per-cell device math shared by the dense and sparse dual-contouring kernels:
// trilinear-gradient normals, QEF accumulation + pseudo-inverse solve + clip, diagonal lengths.
#pragma once
#include "dense.cuh"

namespace isx {

// ---------------------------------------------------------------------------------------------
// normals: central differences of the cell's trilinear interpolant, in the reference's exact
// float32 operation order and FMA contraction pattern (decoded from the SASS nvcc 12.9 emits for
// compute_normals_op, src/its.cu:186-268):  a*(1-t) + b*t  ->  fma(a, 1-t, rn(b*t)), except the
// z-stage shared by the x+- and y+- samples, where c00 and c01 are  fma(b, t, rn(a*(1-t))).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mix_a(float a, float b, float t, float omt) {   // fma(a, 1-t, rn(b*t))
    return __fmaf_rn(a, omt, __fmul_rn(b, t));
}
__device__ __forceinline__ float mix_b(float a, float b, float t, float omt) {   // fma(b, t, rn(a*(1-t)))
    return __fmaf_rn(b, t, __fmul_rn(a, omt));
}
__device__ __forceinline__ float clamp01(float t) { return fmaxf(0.01f, fminf(0.99f, t)); }

__device__ __forceinline__ void cell_normal(const CellData &c, float px_, float py_, float pz_, float &nx, float &ny, float &nz) {
    const float sx = __fsub_rn(c.px[1], c.px[0]), sy = __fsub_rn(c.py[1], c.py[0]), sz = __fsub_rn(c.pz[1], c.pz[0]);
    const float tx = clamp01(__fdiv_rn(__fsub_rn(px_, c.px[0]), sx));
    const float ty = clamp01(__fdiv_rn(__fsub_rn(py_, c.py[0]), sy));
    const float tz = clamp01(__fdiv_rn(__fsub_rn(pz_, c.pz[0]), sz));
    const float eps = 0.02f;
    const float xp = fminf(__fadd_rn(tx, eps), 0.99f), xm = fmaxf(__fsub_rn(tx, eps), 0.01f);
    const float yp = fminf(__fadd_rn(ty, eps), 0.99f), ym = fmaxf(__fsub_rn(ty, eps), 0.01f);
    const float zp = fminf(__fadd_rn(tz, eps), 0.99f), zm = fmaxf(__fsub_rn(tz, eps), 0.01f);
    const float otx = __fsub_rn(1.0f, tx), oty = __fsub_rn(1.0f, ty), otz = __fsub_rn(1.0f, tz);
    const float *v = c.v;
    // z stage at tz (shared by the x and y samples)
    const float c00 = mix_b(v[0], v[1], tz, otz), c01 = mix_b(v[2], v[3], tz, otz);
    const float c10 = mix_a(v[4], v[5], tz, otz), c11 = mix_a(v[6], v[7], tz, otz);
    // d/dx
    const float c0 = mix_a(c00, c01, ty, oty), c1 = mix_a(c10, c11, ty, oty);
    const float fxp = mix_a(c0, c1, xp, __fsub_rn(1.0f, xp)), fxm = mix_a(c0, c1, xm, __fsub_rn(1.0f, xm));
    const float gx = __fdiv_rn(__fsub_rn(fxp, fxm), __fmul_rn(sx, __fsub_rn(xp, xm)));
    // d/dy
    const float oyp = __fsub_rn(1.0f, yp), oym = __fsub_rn(1.0f, ym);
    const float fyp = mix_a(mix_a(c00, c01, yp, oyp), mix_a(c10, c11, yp, oyp), tx, otx);
    const float fym = mix_a(mix_a(c00, c01, ym, oym), mix_a(c10, c11, ym, oym), tx, otx);
    const float gy = __fdiv_rn(__fsub_rn(fyp, fym), __fmul_rn(sy, __fsub_rn(yp, ym)));
    // d/dz
    const float ozp = __fsub_rn(1.0f, zp), ozm = __fsub_rn(1.0f, zm);
    const float p0 = mix_a(mix_a(v[0], v[1], zp, ozp), mix_a(v[2], v[3], zp, ozp), ty, oty);
    const float p1 = mix_a(mix_a(v[4], v[5], zp, ozp), mix_a(v[6], v[7], zp, ozp), ty, oty);
    const float m0 = mix_a(mix_a(v[0], v[1], zm, ozm), mix_a(v[2], v[3], zm, ozm), ty, oty);
    const float m1 = mix_a(mix_a(v[4], v[5], zm, ozm), mix_a(v[6], v[7], zm, ozm), ty, oty);
    const float fzp = mix_a(p0, p1, tx, otx), fzm = mix_a(m0, m1, tx, otx);
    const float gz = __fdiv_rn(__fsub_rn(fzp, fzm), __fmul_rn(sz, __fsub_rn(zp, zm)));
    // |g| = sqrt(fma(gz,gz, fma(gx,gx, gy*gy)))
    const float len = __fsqrt_rn(__fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy))));
    if (len > 1e-8f) {
        nx = __fdiv_rn(gx, len);
        ny = __fdiv_rn(gy, len);
        nz = __fdiv_rn(gz, len);
    } else {
        nx = 0.0f; ny = 0.0f; nz = 1.0f;
    }
}


// ---------------------------------------------------------------------------------------------
// QEF + solve + clip for one cell with k crossings stored at points/normals[o0 .. o0+k).
// Accumulation mirrors get_qef_op's float32 rounding (SASS of src/dc.cu:27-64 of the reference):
//   d = fma(n.z,p.z, fma(n.x,p.x, rn(n.y*p.y)));  ATA_ij = fma(n_i,n_j,ATA_ij);  ATb_i = fma(n_i,d,ATb_i)
//   p_avg = sum(p) / float(k);  ATA_ii += reg;  ATb_i = fma(p_avg_i, reg, ATb_i)
// The solve replaces cuSOLVER gesvdjBatched + cuBLAS gemv (src/batched_la.cu:151-179) with an
// in-register FP64 Jacobi eigen-decomposition of the symmetric 3x3 matrix and the same thresholded
// pseudo-inverse (sigma_j > svd_tol * sigma_max ? 1/sigma_j : 0); the result is clipped to the cell
// AABB [lo,hi] (src/dc.cu:93-98).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void jacobi_rotate(double &app, double &aqq, double &apq, double &arp, double &arq,
                                              double (&V)[3][3], int p_, int q_) {
    if (apq == 0.0) return;
    const double theta = (aqq - app) / (2.0 * apq);
    const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
    const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
    const double tau = s / (1.0 + c);
    const double h = t * apq;
    app -= h;
    aqq += h;
    apq = 0.0;
    const double g = arp, hq = arq;
    arp = g - s * (hq + g * tau);
    arq = hq + s * (g - hq * tau);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const double vp = V[k][p_], vq = V[k][q_];
        V[k][p_] = vp - s * (vq + vp * tau);
        V[k][q_] = vq + s * (vp - vq * tau);
    }
}

__device__ __forceinline__ void qef_solve_clip(const float *__restrict__ points, const float *__restrict__ normals, u32 o0, u32 k,
                                               float reg, float svd_tol, const float lo[3], const float hi[3], float *out) {
    float a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, b0 = 0, b1 = 0, b2 = 0, sx = 0, sy = 0, sz = 0;
    for (u32 i = 0; i < k; i++) {
        const size_t o = 3 * (size_t) (o0 + i);
        const float nx = normals[o], ny = normals[o + 1], nz = normals[o + 2];
        const float qx = points[o], qy = points[o + 1], qz = points[o + 2];
        sx = __fadd_rn(sx, qx); sy = __fadd_rn(sy, qy); sz = __fadd_rn(sz, qz);
        const float d = __fmaf_rn(nz, qz, __fmaf_rn(nx, qx, __fmul_rn(ny, qy)));
        a00 = __fmaf_rn(nx, nx, a00); a01 = __fmaf_rn(nx, ny, a01); a02 = __fmaf_rn(nx, nz, a02);
        a11 = __fmaf_rn(ny, ny, a11); a12 = __fmaf_rn(ny, nz, a12); a22 = __fmaf_rn(nz, nz, a22);
        b0 = __fmaf_rn(nx, d, b0); b1 = __fmaf_rn(ny, d, b1); b2 = __fmaf_rn(nz, d, b2);
    }
    const float kf = (float) k;
    const float ax = __fdiv_rn(sx, kf), ay = __fdiv_rn(sy, kf), az = __fdiv_rn(sz, kf);
    a00 = __fadd_rn(a00, reg); a11 = __fadd_rn(a11, reg); a22 = __fadd_rn(a22, reg);
    b0 = __fmaf_rn(ax, reg, b0); b1 = __fmaf_rn(ay, reg, b1); b2 = __fmaf_rn(az, reg, b2);

    // symmetric eigen-decomposition A = V diag(w) V^T (cyclic Jacobi, FP64)
    double d0 = a00, d1 = a11, d2 = a22, e01 = a01, e02 = a02, e12 = a12;
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
    for (int sweep = 0; sweep < 12; sweep++) {
        // converged far below double precision: the off-diagonal part perturbs the eigenvalues by ~ e^2 / gap (the
        // absolute 1e-40 kept two more sweeps of 15 FP64 divisions / square roots per cell running)
        if (fabs(e01) + fabs(e02) + fabs(e12) <= 1e-20 * (fabs(d0) + fabs(d1) + fabs(d2))) break;
        jacobi_rotate(d0, d1, e01, e02, e12, V, 0, 1);   // (p,q)=(0,1); r=2: a_rp=e02, a_rq=e12
        jacobi_rotate(d0, d2, e02, e01, e12, V, 0, 2);   // (0,2); r=1: a_rp=e01, a_rq=e12
        jacobi_rotate(d1, d2, e12, e01, e02, V, 1, 2);   // (1,2); r=0: a_rp=e01, a_rq=e02
    }
    const double wmax = fmax(d0, fmax(d1, d2));
    const double thr = (double) svd_tol * wmax;
    const double wv[3] = {d0, d1, d2};
    double xs = 0, ys = 0, zs = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (wv[j] > thr) {
            const double c = (V[0][j] * (double) b0 + V[1][j] * (double) b1 + V[2][j] * (double) b2) / wv[j];
            xs += V[0][j] * c; ys += V[1][j] * c; zs += V[2][j] * c;
        }
    }
    out[0] = fminf(fmaxf((float) xs, lo[0]), hi[0]);
    out[1] = fminf(fmaxf((float) ys, lo[1]), hi[1]);
    out[2] = fminf(fmaxf((float) zs, lo[2]), hi[2]);
}

// |a - b| in the reference's rounding: sqrt(fma(dz,dz, fma(dx,dx, rn(dy*dy))))
__device__ __forceinline__ float dist_ref(const float *a, const float *b) {
    const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
}


// Number of sign-change edges of a case.
__device__ __forceinline__ u32 case_edge_count(u32 cs) { return __popc(edge_mask_of_case(cs)); }

}   // namespace isx
