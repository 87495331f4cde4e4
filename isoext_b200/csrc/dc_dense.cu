// dc_dense.cu -- intersections, normals and dual contouring on a dense grid for sm_100a.
//
// Replaces get_intersection / compute_intersection_normals (src/its.cu:93-284), get_qef +
// BatchedLASolver::lsq_svd + fix_dual_v_op (src/dc.cu:14-99, src/batched_la.cu:104-179: cuSOLVER
// gesvdjBatched + 2x cuBLAS gemvStridedBatched + 2 thrust passes), Grid::get_dual_quads
// (src/grid/uniform.cu:60-83: sort_by_key + unique_by_key over 4x-duplicated edges) and
// get_triangles_op + welding (src/dc.cu:101-157,204-217) of the reference.
//
// The front end (sign bits -> ordered entry list + row_start) is shared with marching cubes
// (dense.cuh).  Everything after it is surface-sized:
//   k_its_scan     look-back scan over entries: cell slots + CSR offsets of the (cell,edge) crossings
//   k_its_emit     crossing points (bit-exact lerp), optional normals, is_out bits, CSR arrays
//   k_its_normals  normals only (when the caller supplied an Intersection without normals)
//   k_dc_solve     per cell: QEF accumulation in the reference's float32 rounding pattern, 3x3
//                  symmetric eigen-solve in registers (FP64 Jacobi), pseudo-inverse, clip to the cell
//   k_dc_quads     per owned sign-change edge: the 4 incident cells through row_start (each edge is
//                  owned by exactly one entry -> no sort/unique of duplicated edges), marks used cells
//   k_dc_scan      look-back scan: quad offsets, candidate ids of the used cells
//   k_dc_keys      sortable position keys of the used dual vertices
//   radix_sort96 + k_unique (weld.cuh): reference order = lexicographic positions
//   k_dc_faces     orientation flip, shorter-diagonal split, final ids
#include "dense.cuh"
#include "radix.cuh"
#include "weld.cuh"

namespace isx {

constexpr int IT_ITEMS = 8;
constexpr int IT_TILE = 256 * IT_ITEMS;

// ---- phase-1 workspace ------------------------------------------------------------------------
struct ItsWs {
    u32 *counters;
    u32 *bits;
    u64 *descA, *descC, *descI;
};
static size_t carve_its_ws(Carver &c, const DenseParams &p, size_t cap, ItsWs *out) {
    ItsWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.bits = c.take<u32>(signbits_words(p.P));
    b.descA = c.take<u64>((size_t) p.NQ / CP_TILE + 2);
    b.descC = c.take<u64>(cap / IT_TILE + 2);
    b.descI = c.take<u64>(cap / IT_TILE + 2);
    if (out) *out = b;
    return c.bytes();
}

// Number of sign-change edges of a case.
__device__ __forceinline__ u32 case_edge_count(u32 cs) { return __popc(edge_mask_of_case(cs)); }

// ---------------------------------------------------------------------------------------------
// scan over entries: cellslot (rank among active-cell entries) and its_off (CSR offset)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_its_scan(u32 cap, u32 *__restrict__ counters, const uint2 *__restrict__ entries,
                                                  u32 *__restrict__ cellslot, u32 *__restrict__ its_off,
                                                  u64 *__restrict__ descC, u64 *__restrict__ descI) {
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preC, s_preI;
    const u32 S = counters[C_S];
    if (S > cap) return;
    const u32 ntiles = (S + IT_TILE - 1) / IT_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_B], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 s0 = tile * IT_TILE + threadIdx.x * IT_ITEMS;
        u32 isc[IT_ITEMS], ni[IT_ITEMS], sumC = 0, sumI = 0;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            isc[j] = 0;
            ni[j] = 0;
            if (s < S) {
                const u32 w = entries[s].y;
                if (ent_cell(w)) {
                    isc[j] = 1;
                    ni[j] = case_edge_count(ent_case(w));
                }
            }
            sumC += isc[j];
            sumI += ni[j];
        }
        u32 totC, totI;
        u32 exC = block_exclusive_scan(sumC, &totC, sw);
        u32 exI = block_exclusive_scan(sumI, &totI, sw);
        const u32 warp = threadIdx.x >> 5;
        if (warp == 0) {
            u32 pre = lookback_exclusive(descC, 1, tile, totC, 1u);
            if (threadIdx.x == 0) s_preC = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descI, 1, tile, totI, 1u);
            if ((threadIdx.x & 31) == 0) s_preI = pre;
        }
        __syncthreads();
        exC += s_preC;
        exI += s_preI;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            if (s < S) {
                cellslot[s] = isc[j] ? exC : 0xffffffffu;
                its_off[s] = exI;
                exC += isc[j];
                exI += ni[j];
                if (s == S - 1) {
                    counters[C_T] = exC;    // number of active cells
                    counters[C_I] = exI;    // number of intersections
                }
            }
        }
    }
}

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
// per active cell: crossing points in edge order 0..11 (src/its.cu:37-90), normals, is_out bits
// ---------------------------------------------------------------------------------------------
template <bool POINTS, bool NORMALS>
__global__ void __launch_bounds__(128) k_its_emit(const float *__restrict__ values, DenseParams p,
                                                  const uint2 *__restrict__ entries, u32 S, const u32 *__restrict__ cellslot,
                                                  const u32 *__restrict__ its_off, float *__restrict__ points,
                                                  float *__restrict__ normals, unsigned char *__restrict__ isout,
                                                  u32 *__restrict__ cell_offsets, i64 *__restrict__ cell_indices, u32 n_cells,
                                                  u32 n_its) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 w = e.y, r = e.x, z = ent_z(w);
        if (POINTS) {
            // is_out of the owned edges: v_lo <= v_hi (src/its.cu:78), independent of the level
            const u32 own = ent_own(w);
            u32 io = 0;
            if (own) {
                const i64 n = (i64) r * Z + z;
                const float v0 = __ldg(values + n);
                if ((own & 1u) && v0 <= __ldg(values + n + 1)) io |= 1u;
                if ((own & 2u) && v0 <= __ldg(values + n + Z)) io |= 2u;
                if ((own & 4u) && v0 <= __ldg(values + n + p.YZ)) io |= 4u;
            }
            isout[s] = (unsigned char) io;
        }
        if (!ent_cell(w)) continue;
        const u32 slot = cellslot[s];
        u32 o = its_off[s];
        if (POINTS) {
            const u32 x = r / Y, y = r - x * Y;
            cell_offsets[slot] = o;
            cell_indices[slot] = ((i64) (x + (u32) p.g.x_off) * (Y - 1) + y) * (Z - 1) + z;
            if (slot == n_cells - 1) cell_offsets[n_cells] = n_its;
        }
        CellData c;
        load_cell(values, p, r, z, c);
        const u32 status = edge_mask_of_case(ent_case(w));
#pragma unroll
        for (int k = 0; k < 12; k++) {
            if (!((status >> k) & 1u)) continue;
            float qx, qy, qz;
            if (POINTS) {
                cell_edge_point(c, k, p.level, qx, qy, qz);
                points[3 * (size_t) o + 0] = qx;
                points[3 * (size_t) o + 1] = qy;
                points[3 * (size_t) o + 2] = qz;
            } else {
                qx = points[3 * (size_t) o + 0];
                qy = points[3 * (size_t) o + 1];
                qz = points[3 * (size_t) o + 2];
            }
            if (NORMALS) {
                float nx, ny, nz;
                cell_normal(c, qx, qy, qz, nx, ny, nz);
                normals[3 * (size_t) o + 0] = nx;
                normals[3 * (size_t) o + 1] = ny;
                normals[3 * (size_t) o + 2] = nz;
            }
            o++;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// QEF + solve + clip, one thread per active cell.
// Accumulation mirrors get_qef_op's float32 rounding (SASS of src/dc.cu:27-64):
//   d = fma(n.z,p.z, fma(n.x,p.x, rn(n.y*p.y)));  ATA_ij = fma(n_i,n_j,ATA_ij);  ATb_i = fma(n_i,d,ATb_i)
//   p_avg = sum(p) / float(k);  ATA_ii += reg;  ATb_i = fma(p_avg_i, reg, ATb_i)
// The solve replaces cuSOLVER gesvdjBatched + cuBLAS gemv (src/batched_la.cu:151-179) with an
// in-register FP64 Jacobi eigen-decomposition of the symmetric 3x3 matrix and the same thresholded
// pseudo-inverse (sigma_j > svd_tol * sigma_max ? 1/sigma_j : 0).
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

__global__ void __launch_bounds__(128) k_dc_solve(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ cellslot, const u32 *__restrict__ its_off,
                                                  const float *__restrict__ points, const float *__restrict__ normals,
                                                  float reg, float svd_tol, float *__restrict__ dual_v) {
    const u32 Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 w = e.y;
        if (!ent_cell(w)) continue;
        const u32 slot = cellslot[s];
        const u32 o0 = its_off[s], k = case_edge_count(ent_case(w));
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
            if (fabs(e01) + fabs(e02) + fabs(e12) < 1e-40) break;
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
        // clip to the cell AABB (src/dc.cu:93-98)
        const u32 r = e.x, z = ent_z(w), x = r / Y, y = r - x * Y;
        const u32 xg = x + (u32) p.g.x_off;
        const float lx = axis_pos(xg, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
        const float hx = axis_pos(xg + 1, (u32) p.g.Xg - 1, p.g.amin[0], p.g.asize[0]);
        const float ly = axis_pos(y, Y - 1, p.g.amin[1], p.g.asize[1]), hy = axis_pos(y + 1, Y - 1, p.g.amin[1], p.g.asize[1]);
        const float lz = axis_pos(z, Z - 1, p.g.amin[2], p.g.asize[2]), hz = axis_pos(z + 1, Z - 1, p.g.amin[2], p.g.asize[2]);
        dual_v[3 * (size_t) slot + 0] = fminf(fmaxf((float) xs, lx), hx);
        dual_v[3 * (size_t) slot + 1] = fminf(fmaxf((float) ys, ly), hy);
        dual_v[3 * (size_t) slot + 2] = fminf(fmaxf((float) zs, lz), hz);
    }
}

// ---------------------------------------------------------------------------------------------
// quads: the 4 cells around an owned sign-change edge (src/shared_luts.cu:78-82 gives the cyclic
// order: cell = edge origin - offset_k).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 find_cell(const uint2 *__restrict__ entries, const u32 *__restrict__ row_start,
                                         const u32 *__restrict__ cellslot, u32 row, u32 z) {
    const u32 lo = row_start[row], hi = row_start[row + 1];
    const u32 i = row_lower_bound(entries, lo, hi, z);
    if (i >= hi || ent_z(entries[i].y) != z) return 0xffffffffu;
    return cellslot[i];   // 0xffffffff when the entry is not an active cell
}

// returns false if the quad must be skipped (a neighbour outside the cell grid or not active)
__device__ __forceinline__ bool quad_cells(const DenseParams &p, const uint2 *__restrict__ entries,
                                           const u32 *__restrict__ row_start, const u32 *__restrict__ cellslot, u32 r, u32 z,
                                           int axis, u32 q[4]) {
    const u32 X = (u32) p.g.X, Y = (u32) p.g.Y, Z = (u32) p.g.Z;
    const u32 x = r / Y, y = r - x * Y;
    // offsets (dx,dy,dz) per k; axis: 0 = +z edge, 1 = +y edge, 2 = +x edge
    int dx[4], dy[4], dz[4];
    if (axis == 2) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 1; dz[1] = 0; dx[2] = 0; dy[2] = 1; dz[2] = 1; dx[3] = 0; dy[3] = 0; dz[3] = 1; }
    else if (axis == 1) { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 0; dy[1] = 0; dz[1] = 1; dx[2] = 1; dy[2] = 0; dz[2] = 1; dx[3] = 1; dy[3] = 0; dz[3] = 0; }
    else { dx[0] = 0; dy[0] = 0; dz[0] = 0; dx[1] = 1; dy[1] = 0; dz[1] = 0; dx[2] = 1; dy[2] = 1; dz[2] = 0; dx[3] = 0; dy[3] = 1; dz[3] = 0; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (x < (u32) dx[k] || y < (u32) dy[k] || z < (u32) dz[k]) return false;   // include/utils.cuh:129-131
        const u32 cx = x - dx[k], cy = y - dy[k], cz = z - dz[k];
        if (cx + 1 >= X || cy + 1 >= Y || cz + 1 >= Z) return false;               // (reference: undefined behaviour)
        q[k] = find_cell(entries, row_start, cellslot, cx * Y + cy, cz);
        if (q[k] == 0xffffffffu) return false;
    }
    return true;
}

__global__ void __launch_bounds__(128) k_dc_quads(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ row_start, const u32 *__restrict__ cellslot,
                                                  unsigned char *__restrict__ qmask, unsigned char *__restrict__ used) {
    const u32 Y = (u32) p.g.Y;
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const uint2 e = entries[s];
        const u32 own = ent_own(e.y);
        u32 m = 0;
        const u32 x = e.x / Y;
        // quads are attributed to the slab that owns the edge's origin plane
        if (own && x >= p.emit_lo && x < p.emit_hi + 1) {
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (!((own >> a) & 1u)) continue;
                u32 q[4];
                if (quad_cells(p, entries, row_start, cellslot, e.x, ent_z(e.y), a, q)) {
                    m |= 1u << a;
                    used[q[0]] = 1; used[q[1]] = 1; used[q[2]] = 1; used[q[3]] = 1;
                }
            }
        }
        qmask[s] = (unsigned char) m;
    }
}

// scan over entries: quad offsets + candidate ids of used cells
__global__ void __launch_bounds__(256) k_dc_scan(u32 S, u32 *__restrict__ counters, const u32 *__restrict__ cellslot,
                                                 const unsigned char *__restrict__ qmask, const unsigned char *__restrict__ used,
                                                 u32 *__restrict__ quad_off, u32 *__restrict__ cand_of_cell,
                                                 u64 *__restrict__ descQ, u64 *__restrict__ descU) {
    __shared__ u32 sw[33];
    __shared__ u32 s_tile, s_preQ, s_preU;
    const u32 ntiles = (S + IT_TILE - 1) / IT_TILE;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&counters[C_TICKET_D], 1u);
        __syncthreads();
        const u32 tile = s_tile;
        if (tile >= ntiles) break;
        const u32 s0 = tile * IT_TILE + threadIdx.x * IT_ITEMS;
        u32 nq[IT_ITEMS], us[IT_ITEMS], slot[IT_ITEMS], sumQ = 0, sumU = 0;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            nq[j] = 0; us[j] = 0; slot[j] = 0xffffffffu;
            if (s < S) {
                nq[j] = __popc((u32) qmask[s]);
                slot[j] = cellslot[s];
                if (slot[j] != 0xffffffffu) us[j] = used[slot[j]] ? 1u : 0u;
            }
            sumQ += nq[j];
            sumU += us[j];
        }
        u32 totQ, totU;
        u32 exQ = block_exclusive_scan(sumQ, &totQ, sw);
        u32 exU = block_exclusive_scan(sumU, &totU, sw);
        const u32 warp = threadIdx.x >> 5;
        if (warp == 0) {
            u32 pre = lookback_exclusive(descQ, 1, tile, totQ, 1u);
            if (threadIdx.x == 0) s_preQ = pre;
        } else if (warp == 1) {
            u32 pre = lookback_exclusive(descU, 1, tile, totU, 1u);
            if ((threadIdx.x & 31) == 0) s_preU = pre;
        }
        __syncthreads();
        exQ += s_preQ;
        exU += s_preU;
#pragma unroll
        for (int j = 0; j < IT_ITEMS; j++) {
            const u32 s = s0 + j;
            if (s < S) {
                quad_off[s] = exQ;
                if (slot[j] != 0xffffffffu) cand_of_cell[slot[j]] = us[j] ? exU : 0xffffffffu;
                exQ += nq[j];
                exU += us[j];
                if (s == S - 1) {
                    counters[C_Q] = exQ;
                    counters[C_VC] = exU;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_dc_keys(u32 n_cells, const u32 *__restrict__ cand_of_cell, const float *__restrict__ dual_v,
                                                 u32 *__restrict__ kx, u32 *__restrict__ ky, u32 *__restrict__ kz) {
    for (u32 c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
        const u32 id = cand_of_cell[c];
        if (id == 0xffffffffu) continue;
        kx[id] = float_key(dual_v[3 * (size_t) c + 0]);
        ky[id] = float_key(dual_v[3 * (size_t) c + 1]);
        kz[id] = float_key(dual_v[3 * (size_t) c + 2]);
    }
}

// |a - b| in the reference's rounding: sqrt(fma(dz,dz, fma(dx,dx, rn(dy*dy))))
__device__ __forceinline__ float dist_ref(const float *a, const float *b) {
    const float dx = __fsub_rn(a[0], b[0]), dy = __fsub_rn(a[1], b[1]), dz = __fsub_rn(a[2], b[2]);
    return __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))));
}

__global__ void __launch_bounds__(128) k_dc_faces(DenseParams p, const uint2 *__restrict__ entries, u32 S,
                                                  const u32 *__restrict__ row_start, const u32 *__restrict__ cellslot,
                                                  const unsigned char *__restrict__ qmask, const unsigned char *__restrict__ isout,
                                                  const u32 *__restrict__ quad_off, const u32 *__restrict__ cand_of_cell,
                                                  const u32 *__restrict__ cand_rank, const float *__restrict__ dual_v,
                                                  int *__restrict__ F, int *__restrict__ quads_out) {
    for (u32 s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
        const u32 m = qmask[s];
        if (!m) continue;
        const uint2 e = entries[s];
        const u32 io = isout[s];
        u32 qi = quad_off[s];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (!((m >> a) & 1u)) continue;
            u32 q[4];
            quad_cells(p, entries, row_start, cellslot, e.x, ent_z(e.y), a, q);
            if (!((io >> a) & 1u)) {   // edge points inward: reverse the loop (src/dc.cu:123-127)
                const u32 t0 = q[0], t1 = q[1];
                q[0] = q[3]; q[1] = q[2]; q[2] = t1; q[3] = t0;
            }
            if (quads_out) {
                quads_out[4 * (size_t) qi + 0] = (int) q[0]; quads_out[4 * (size_t) qi + 1] = (int) q[1];
                quads_out[4 * (size_t) qi + 2] = (int) q[2]; quads_out[4 * (size_t) qi + 3] = (int) q[3];
            }
            const float *v0 = dual_v + 3 * (size_t) q[0], *v1 = dual_v + 3 * (size_t) q[1];
            const float *v2 = dual_v + 3 * (size_t) q[2], *v3 = dual_v + 3 * (size_t) q[3];
            int id[4];
#pragma unroll
            for (int k = 0; k < 4; k++) id[k] = (int) cand_rank[cand_of_cell[q[k]]];
            int *f = F + 6 * (size_t) qi;
            if (dist_ref(v0, v2) > dist_ref(v1, v3)) {   // split along v1-v3 (src/dc.cu:139-147)
                f[0] = id[1]; f[1] = id[3]; f[2] = id[0];
                f[3] = id[3]; f[4] = id[1]; f[5] = id[2];
            } else {                                      // split along v0-v2 (src/dc.cu:148-155)
                f[0] = id[2]; f[1] = id[0]; f[2] = id[1];
                f[3] = id[0]; f[4] = id[2]; f[5] = id[3];
            }
            qi++;
        }
    }
}

// ---- DC workspaces -----------------------------------------------------------------------------
struct DcWs {
    u32 *counters;
    unsigned char *qmask;   // S
    unsigned char *used;    // n_cells
    u32 *quad_off;          // S
    u32 *cand_of_cell;      // n_cells
    u64 *descQ, *descU;
};
static size_t carve_dc_ws(Carver &c, size_t S, size_t n_cells, DcWs *out) {
    DcWs b;
    b.counters = c.take<u32>(C_COUNT);
    b.qmask = c.take<unsigned char>(S + 1);
    b.used = c.take<unsigned char>(n_cells + 1);
    b.quad_off = c.take<u32>(S + 1);
    b.cand_of_cell = c.take<u32>(n_cells + 1);
    b.descQ = c.take<u64>(S / IT_TILE + 2);
    b.descU = c.take<u64>(S / IT_TILE + 2);
    if (out) *out = b;
    return c.bytes();
}
struct DcScratch {
    u32 *kx, *ky, *kz, *cand_rank;
    u64 *descV;
    RadixBuffers radix;
};
static size_t carve_dc_scratch(Carver &c, size_t nc, DcScratch *out) {
    DcScratch s;
    s.kx = c.take<u32>(nc);
    s.ky = c.take<u32>(nc);
    s.kz = c.take<u32>(nc);
    s.cand_rank = c.take<u32>(nc);
    s.descV = c.take<u64>(nc / UQ_TILE + 2);
    RadixBuffers::carve(c, nc, &s.radix);
    if (out) *out = s;
    return c.bytes();
}

int make_dense_params(i64 X, i64 Y, i64 Z, i64 x_off, i64 Xg, const float *amin, const float *amax, float level,
                      i64 emit_lo, i64 emit_hi, DenseParams *out);
int device_sms();

}   // namespace isx

using namespace isx;

extern "C" {

size_t isoext_its_dense_workspace_bytes(int64_t X, int64_t Y, int64_t Z, int64_t cap_entries) {
    DenseParams p;
    float z3[3] = {0, 0, 0}, o3[3] = {1, 1, 1};
    if (make_dense_params(X, Y, Z, 0, X, z3, o3, 0.f, 0, X - 1, &p) != OK) return 0;
    Carver c(nullptr);
    return carve_its_ws(c, p, (size_t) cap_entries, nullptr);
}

// Phase 1 of get_intersection: classify + compact + CSR scan.
//   entries (cap_entries+1 x uint2), row_start (X*Y+2 x u32), cellslot / its_off (cap_entries x u32):
//   caller-owned outputs that the Intersection keeps.
//   counts_out[0..2] = entries S, active cells, intersections I.
int isoext_its_dense_count(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                           const float *aabb_min, const float *aabb_max, float level, void *workspace, size_t workspace_bytes,
                           int64_t cap_entries, void *entries, uint32_t *row_start, uint32_t *cellslot, uint32_t *its_off,
                           void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, 0, X - 1, &p);
    if (rc != OK) return rc;
    if ((reinterpret_cast<uintptr_t>(values) & 15u) != 0) return fail(E_INVALID, "values must be 16-byte aligned");
    if (cap_entries < 1 || cap_entries >= ((i64) 1 << 29)) return fail(E_INVALID, "cap_entries out of range");
    Carver c(workspace);
    ItsWs b;
    if (carve_its_ws(c, p, (size_t) cap_entries, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const u32 cap = (u32) cap_entries;
    uint2 *ent = static_cast<uint2 *>(entries);
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.descA, 0, ((size_t) p.NQ / CP_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descC, 0, ((size_t) cap / IT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descI, 0, ((size_t) cap / IT_TILE + 2) * sizeof(u64), stream));
    const int sms = device_sms();
    {
        i64 groups = p.P >> 7;
        i64 want = (groups + 8 * SB_UNROLL - 1) / (8 * SB_UNROLL);
        int blocks = (int) (want < 1 ? 1 : (want > (i64) sms * 8 ? (i64) sms * 8 : want));
        stream_timer_mark(stream);
        ISX_LAUNCH(k_signbits, blocks, 256, 0, stream, values, b.bits, p.P, level);
        stream_timer_mark(stream);
    }
    if ((p.g.Z & 127) == 0) {
        const u32 nspans = p.R * (u32) (p.g.Z >> 7);
        ISX_LAUNCH(k_compact128, (nspans + SP_TILE - 1) / SP_TILE, 256, 0, stream, b.bits, p, ent, cap, row_start, b.descA, b.counters);
    } else {
        ISX_LAUNCH(k_compact, (p.NQ + CP_TILE - 1) / CP_TILE, 256, 0, stream, b.bits, p, ent, cap, row_start, b.descA, b.counters);
    }
    ISX_LAUNCH(k_its_scan, sms * 4, 256, 0, stream, cap, b.counters, ent, cellslot, its_off, b.descC, b.descI);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_S];
    counts_out[1] = h[C_T];
    counts_out[2] = h[C_I];
    if (h[C_S] > cap) return fail(E_CAPACITY, "entry capacity exceeded; retry with cap_entries >= counts_out[0]");
    return OK;
}

// Phase 2 of get_intersection: points (I x 3), normals (I x 3, only if compute_normals), is_out bits
// per entry, CSR cell_offsets (n_cells+1, u32) and cell_indices (n_cells, i64).
int isoext_its_dense_emit(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                          const float *aabb_min, const float *aabb_max, float level, int compute_normals, const void *entries,
                          int64_t n_entries, const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, int64_t n_its,
                          float *points, float *normals, unsigned char *isout, uint32_t *cell_offsets, int64_t *cell_indices,
                          void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, level, 0, X - 1, &p);
    if (rc != OK) return rc;
    if (n_entries <= 0) return OK;
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const int sms = device_sms();
    if (compute_normals)
        ISX_LAUNCH((k_its_emit<true, true>), sms * 8, 128, 0, stream, values, p, ent, (u32) n_entries, cellslot, its_off, points,
                   normals, isout, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its);
    else
        ISX_LAUNCH((k_its_emit<true, false>), sms * 8, 128, 0, stream, values, p, ent, (u32) n_entries, cellslot, its_off, points,
                   normals, isout, cell_offsets, cell_indices, (u32) n_cells, (u32) n_its);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

// compute_intersection_normals (src/its.cu:270-284) for an existing Intersection.
int isoext_its_dense_normals(const float *values, int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global,
                             const float *aabb_min, const float *aabb_max, const void *entries, int64_t n_entries,
                             const uint32_t *cellslot, const uint32_t *its_off, const float *points, float *normals,
                             void *stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X - 1, &p);
    if (rc != OK) return rc;
    if (n_entries <= 0) return OK;
    ISX_LAUNCH((k_its_emit<false, true>), device_sms() * 8, 128, 0, stream, values, p, static_cast<const uint2 *>(entries),
               (u32) n_entries, cellslot, its_off, const_cast<float *>(points), normals, nullptr, nullptr, nullptr, 0u, 0u);
    ISX_CUDA(cudaGetLastError());
    return OK;
}

size_t isoext_dc_dense_workspace_bytes(int64_t n_entries, int64_t n_cells) {
    Carver c(nullptr);
    return carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, nullptr);
}
size_t isoext_dc_dense_scratch_bytes(int64_t n_candidates) {
    Carver c(nullptr);
    return carve_dc_scratch(c, (size_t) (n_candidates > 0 ? n_candidates : 1), nullptr);
}

// Phase 1 of dual_contouring: dual vertices (n_cells x 3, clipped) + quad / candidate counts.
//   counts_out[0..1] = quads Q, used dual vertices Vc.
int isoext_dc_dense_count(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                          const float *aabb_max, const void *entries, int64_t n_entries, const uint32_t *row_start,
                          const uint32_t *cellslot, const uint32_t *its_off, int64_t n_cells, const float *points,
                          const float *normals, float reg, float svd_tol, float *dual_v, void *workspace,
                          size_t workspace_bytes, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X - 1, &p);
    if (rc != OK) return rc;
    counts_out[0] = counts_out[1] = 0;
    if (n_entries <= 0 || n_cells <= 0) return OK;
    Carver c(workspace);
    DcWs b;
    if (carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const u32 S = (u32) n_entries;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(b.counters, 0, C_COUNT * sizeof(u32), stream));
    ISX_CUDA(cudaMemsetAsync(b.used, 0, (size_t) n_cells + 1, stream));
    ISX_CUDA(cudaMemsetAsync(b.descQ, 0, ((size_t) S / IT_TILE + 2) * sizeof(u64), stream));
    ISX_CUDA(cudaMemsetAsync(b.descU, 0, ((size_t) S / IT_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_dc_solve, sms * 8, 128, 0, stream, p, ent, S, cellslot, its_off, points, normals, reg, svd_tol, dual_v);
    ISX_LAUNCH(k_dc_quads, sms * 8, 128, 0, stream, p, ent, S, row_start, cellslot, b.qmask, b.used);
    ISX_LAUNCH(k_dc_scan, sms * 4, 256, 0, stream, S, b.counters, cellslot, b.qmask, b.used, b.quad_off, b.cand_of_cell, b.descQ, b.descU);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_Q];
    counts_out[1] = h[C_VC];
    return OK;
}

// Phase 2: V (capacity Vc x 3), F (2Q x 3 int32), optional quads_out (Q x 4 cell slots, oriented).
//   counts_out[0] = welded vertices.
int isoext_dc_dense_emit(int64_t X, int64_t Y, int64_t Z, int64_t x_offset, int64_t X_global, const float *aabb_min,
                         const float *aabb_max, const void *entries, int64_t n_entries, const uint32_t *row_start,
                         const uint32_t *cellslot, const unsigned char *isout, int64_t n_cells, const float *dual_v,
                         void *workspace, size_t workspace_bytes, void *scratch, size_t scratch_bytes, int64_t n_candidates,
                         float *V, int32_t *F, int32_t *quads_out, void *stream_, int64_t *counts_out) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    DenseParams p;
    int rc = make_dense_params(X, Y, Z, x_offset, X_global, aabb_min, aabb_max, 0.f, 0, X - 1, &p);
    if (rc != OK) return rc;
    counts_out[0] = 0;
    if (n_candidates <= 0) return OK;
    Carver c(workspace);
    DcWs b;
    if (carve_dc_ws(c, (size_t) n_entries, (size_t) n_cells, &b) > workspace_bytes) return fail(E_WORKSPACE, "workspace too small");
    Carver cs(scratch);
    DcScratch s;
    if (carve_dc_scratch(cs, (size_t) n_candidates, &s) > scratch_bytes) return fail(E_WORKSPACE, "scratch too small");
    const uint2 *ent = static_cast<const uint2 *>(entries);
    const u32 S = (u32) n_entries, nc = (u32) n_candidates;
    const int sms = device_sms();
    ISX_CUDA(cudaMemsetAsync(s.descV, 0, ((size_t) nc / UQ_TILE + 2) * sizeof(u64), stream));
    ISX_LAUNCH(k_dc_keys, sms * 8, 256, 0, stream, (u32) n_cells, b.cand_of_cell, dual_v, s.kx, s.ky, s.kz);
    ISX_CUDA(radix_sort96(s.kx, s.ky, s.kz, nc, s.radix, stream));
    ISX_LAUNCH(k_unique, sms * 4, 256, 0, stream, nc, s.radix.perm[0], s.kx, s.ky, s.kz, s.cand_rank, V, b.counters, s.descV,
               host_float_key(-INFINITY), host_float_key(INFINITY));
    ISX_LAUNCH(k_dc_faces, sms * 8, 128, 0, stream, p, ent, S, row_start, cellslot, b.qmask, isout, b.quad_off, b.cand_of_cell,
               s.cand_rank, dual_v, F, quads_out);
    ISX_CUDA(cudaGetLastError());
    u32 h[C_COUNT];
    ISX_CUDA(cudaMemcpyAsync(h, b.counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISX_CUDA(cudaStreamSynchronize(stream));
    counts_out[0] = h[C_V];
    return OK;
}

}   // extern "C"
