/* oracle/oracle_c.c -- TEST INFRASTRUCTURE (the parity checker), NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the iso-surface extraction path of
 * GuangyanCai/isoext v0.5.1 (reference mounted at /root/reference; citations are relative to it).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * The product (isoext_b200/) never imports, links or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against every known-answer
 * count the reference records in its executed notebooks (doc/marching_cubes.ipynb:60-61,113-114,
 * 144-146; doc/grids.ipynb:172; doc/occupancy_grids.ipynb:88; doc/quickstart.ipynb:21;
 * doc/dual_contouring.ipynb:51-52,104) and tests/test_mc_gpu.py / tests/test_dc_parity3_gpu.py (GPU) checks it bit-for-bit
 * against the reference's own CUDA build (oracle/_ref/libisoext_ref.so).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math  (no FMA contraction: every fused
 * operation below is an explicit fmaf(), mirroring the SASS nvcc 12.9 emits for the reference
 * with its default -fmad=true: FADD 1-t; FMUL t*b; FFMA a*(1-t)+t*b; IEEE div).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_luts.h"

/* corner i = (x + (i>>2&1), y + (i>>1&1), z + (i&1))           include/shared_luts.cuh:31-33 */
/* edge e joins corners ORC_EDGES[e][0] < ORC_EDGES[e][1]       src/shared_luts.cu:3-4        */
static const int ORC_EDGES[12][2] = {{0, 1}, {1, 3}, {2, 3}, {0, 2}, {4, 5}, {5, 7},
                                     {6, 7}, {4, 6}, {0, 4}, {1, 5}, {3, 7}, {2, 6}};

typedef struct { float x, y, z; } f3;

typedef struct {
    float *v;      /* nv x 3 */
    int32_t *f;    /* nf x 3 */
    int64_t nv, nf, n_active;
} orc_mesh;

typedef struct {
    int64_t X, Y, Z;
    float amin[3], asize[3];
} orc_geom;

static orc_geom make_geom(int64_t X, int64_t Y, int64_t Z, const float *amin, const float *amax) {
    orc_geom g;
    g.X = X; g.Y = Y; g.Z = Z;
    for (int a = 0; a < 3; a++) {
        g.amin[a] = amin[a];
        g.asize[a] = amax[a] - amin[a]; /* float subtraction on the host: include/utils.cuh:69 */
    }
    return g;
}

/* include/utils.cuh:71-79 -- pos = aabb_min + (i / float(res-1)) * aabb_size, contracted to one FFMA */
static inline float axis_pos(int64_t i, int64_t res, float amin, float asize) {
    float q = (float) (uint32_t) i / (float) (uint32_t) (res - 1);
    return fmaf(q, asize, amin);
}

static inline f3 point_pos(const orc_geom *g, int64_t x, int64_t y, int64_t z) {
    f3 p;
    p.x = axis_pos(x, g->X, g->amin[0], g->asize[0]);
    p.y = axis_pos(y, g->Y, g->amin[1], g->asize[1]);
    p.z = axis_pos(z, g->Z, g->amin[2], g->asize[2]);
    return p;
}

/* src/grid/uniform.cu:22-30 */
void orc_points_dense(int64_t X, int64_t Y, int64_t Z, const float *amin, const float *amax, float *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    for (int64_t x = 0; x < X; x++)
        for (int64_t y = 0; y < Y; y++)
            for (int64_t z = 0; z < Z; z++) {
                f3 p = point_pos(&g, x, y, z);
                float *o = out + 3 * ((x * Y + y) * Z + z);
                o[0] = p.x; o[1] = p.y; o[2] = p.z;
            }
}

/* src/grid/sparse.cu:21-36 -- corner positions of the given cells; cell index = x(Y-1)(Z-1)+y(Z-1)+z */
void orc_points_cells(int64_t X, int64_t Y, int64_t Z, const float *amin, const float *amax,
                      const int64_t *cell_idx, int64_t n, float *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    for (int64_t c = 0; c < n; c++) {
        int64_t i = cell_idx[c];
        int64_t z = i % (Z - 1); i /= (Z - 1);
        int64_t y = i % (Y - 1);
        int64_t x = i / (Y - 1);
        for (int k = 0; k < 8; k++) {
            f3 p = point_pos(&g, x + (k >> 2 & 1), y + (k >> 1 & 1), z + (k & 1));
            float *o = out + 3 * (8 * c + k);
            o[0] = p.x; o[1] = p.y; o[2] = p.z;
        }
    }
}

/* include/utils.cuh:94-99 */
static inline int case_of(const float *cv, float level) {
    int c = 0;
    for (int i = 0; i < 8; i++) c |= ((cv[i] - level < 0) ? 1 : 0) << i;
    return c;
}

static inline int sign_change_edges(int c) {
    int m = 0;
    for (int e = 0; e < 12; e++)
        if (((c >> ORC_EDGES[e][0]) ^ (c >> ORC_EDGES[e][1])) & 1) m |= 1 << e;
    return m;
}

/* src/mc/nagae.cu:54-56 + include/math.cuh:77-81 */
static inline f3 edge_point(f3 a, f3 b, float v0, float v1, float level) {
    float denom = v1 - v0;
    float t = (denom != 0.0f) ? (level - v0) / denom : 0.0f;
    float s = 1 - t;
    f3 p;
    p.x = fmaf(a.x, s, t * b.x);
    p.y = fmaf(a.y, s, t * b.y);
    p.z = fmaf(a.z, s, t * b.z);
    return p;
}

static inline int f3_ne(f3 a, f3 b) { return a.x != b.x || a.y != b.y || a.z != b.z; }

typedef struct { f3 *d; int64_t n, cap; } f3vec;
static void f3vec_push(f3vec *v, f3 p) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->d = (f3 *) realloc(v->d, (size_t) v->cap * sizeof(f3));
    }
    v->d[v->n++] = p;
}

/* src/mc/nagae.cu:34-80 / src/mc/lorensen.cu:34-80 */
static void process_cube(const float *cv, const f3 *cp, int case_num, float level, int method, f3vec *out) {
    int status = sign_change_edges(case_num); /* == edge_status_table[case] (tools/gen_luts.py) */
    f3 cube_v[12];
    for (int e = 0; e < 12; e++)
        if (status >> e & 1) {
            int p0 = ORC_EDGES[e][0], p1 = ORC_EDGES[e][1];
            cube_v[e] = edge_point(cp[p0], cp[p1], cv[p0], cv[p1], level);
        }
    const signed char *row = method == 0 ? ORC_TRI_NAGAE[case_num] : ORC_TRI_LORENSEN[case_num];
    int max_len = method == 0 ? 15 : 12;
    for (int i = 0; i < max_len; i += 3) {
        if (row[i] == -1) break;
        f3 v0 = cube_v[row[i]], v1 = cube_v[row[i + 1]], v2 = cube_v[row[i + 2]];
        if (f3_ne(v0, v1) && f3_ne(v0, v2) && f3_ne(v1, v2)) {
            f3vec_push(out, v0); f3vec_push(out, v1); f3vec_push(out, v2);
        }
    }
}

/* include/math.cuh:112-126 */
static int f3_less(const f3 *a, const f3 *b) {
    if (a->x < b->x) return 1;
    if (b->x < a->x) return 0;
    if (a->y < b->y) return 1;
    if (b->y < a->y) return 0;
    return a->z < b->z;
}
static int f3_cmp(const void *pa, const void *pb) {
    const f3 *a = (const f3 *) pa, *b = (const f3 *) pb;
    if (f3_less(a, b)) return -1;
    if (f3_less(b, a)) return 1;
    return 0;
}

/* src/utils.cu:32-59 -- V = sorted unique corner positions, F = lower_bound rank of each corner */
static void weld(f3vec *corners, orc_mesh *out) {
    int64_t n = corners->n;
    out->nf = n / 3;
    if (n == 0) { out->nv = 0; out->v = NULL; out->f = NULL; return; }
    f3 *s = (f3 *) malloc((size_t) n * sizeof(f3));
    memcpy(s, corners->d, (size_t) n * sizeof(f3));
    qsort(s, (size_t) n, sizeof(f3), f3_cmp);
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++)
        if (i == 0 || f3_ne(s[i], s[m - 1])) s[m++] = s[i];
    out->nv = m;
    out->v = (float *) malloc((size_t) m * 3 * sizeof(float));
    memcpy(out->v, s, (size_t) m * sizeof(f3));
    out->f = (int32_t *) malloc((size_t) n * sizeof(int32_t));
    for (int64_t i = 0; i < n; i++) {
        int64_t lo = 0, hi = m;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (f3_less(&s[mid], &corners->d[i])) lo = mid + 1; else hi = mid;
        }
        out->f[i] = (int32_t) lo;
    }
    free(s);
}

void orc_mesh_free(orc_mesh *m) {
    free(m->v); free(m->f);
    m->v = NULL; m->f = NULL;
}

/* src/mc/mc.cu:17-68 on a UniformGrid.  method: 0 = nagae, 1 = lorensen.
 * x0,x1: cell-x range [x0,x1) to extract (whole grid: 0, X-1) -- used to check slab sharding. */
int orc_mc_dense(const float *values, int64_t X, int64_t Y, int64_t Z, const float *amin,
                 const float *amax, float level, int method, int64_t x0, int64_t x1, orc_mesh *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    f3vec corners = {0};
    int64_t n_active = 0;
    for (int64_t x = x0; x < x1; x++)
        for (int64_t y = 0; y + 1 < Y; y++)
            for (int64_t z = 0; z + 1 < Z; z++) {
                float cv[8];
                for (int i = 0; i < 8; i++)
                    cv[i] = values[((x + (i >> 2 & 1)) * Y + (y + (i >> 1 & 1))) * Z + z + (i & 1)];
                int c = case_of(cv, level);
                if (c == 0 || c == 255) continue;
                n_active++;
                f3 cp[8];
                for (int i = 0; i < 8; i++)
                    cp[i] = point_pos(&g, x + (i >> 2 & 1), y + (i >> 1 & 1), z + (i & 1));
                process_cube(cv, cp, c, level, method, &corners);
            }
    out->n_active = n_active;
    weld(&corners, out);
    free(corners.d);
    return 0;
}

/* src/mc/mc.cu:17-68 on a SparseGrid: cells = sorted cell indices, values8 = (N,8) corner values */
int orc_mc_sparse(const float *values8, const int64_t *cell_idx, int64_t N, int64_t X, int64_t Y,
                  int64_t Z, const float *amin, const float *amax, float level, int method, orc_mesh *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    f3vec corners = {0};
    int64_t n_active = 0;
    for (int64_t s = 0; s < N; s++) {
        const float *cv = values8 + 8 * s;
        int c = case_of(cv, level);
        if (c == 0 || c == 255) continue;
        n_active++;
        int64_t i = cell_idx[s];
        int64_t z = i % (Z - 1); i /= (Z - 1);
        int64_t y = i % (Y - 1);
        int64_t x = i / (Y - 1);
        f3 cp[8];
        for (int k = 0; k < 8; k++) cp[k] = point_pos(&g, x + (k >> 2 & 1), y + (k >> 1 & 1), z + (k & 1));
        process_cube(cv, cp, c, level, method, &corners);
    }
    out->n_active = n_active;
    weld(&corners, out);
    free(corners.d);
    return 0;
}

/* Case histogram + active count only (cheap, used for full-size property checks). */
int64_t orc_case_histogram(const float *values, int64_t X, int64_t Y, int64_t Z, float level, int64_t *hist256) {
    memset(hist256, 0, 256 * sizeof(int64_t));
    int64_t n_active = 0;
    for (int64_t x = 0; x + 1 < X; x++)
        for (int64_t y = 0; y + 1 < Y; y++)
            for (int64_t z = 0; z + 1 < Z; z++) {
                float cv[8];
                for (int i = 0; i < 8; i++)
                    cv[i] = values[((x + (i >> 2 & 1)) * Y + (y + (i >> 1 & 1))) * Z + z + (i & 1)];
                int c = case_of(cv, level);
                hist256[c]++;
                if (c != 0 && c != 255) n_active++;
            }
    return n_active;
}

/* ======================================================================================
 * Intersections + dual contouring (src/its.cu, src/dc.cu, src/batched_la.cu)
 * ====================================================================================== */
typedef struct {
    float *points;          /* I x 3 */
    float *normals;         /* I x 3 (zeros unless computed) */
    uint64_t *edges;        /* I x 2 global point ids (p_lo, p_hi); uniform-grid ids for both grid kinds */
    uint8_t *is_out;        /* I */
    int64_t *cell_indices;  /* S   (dense: cell index; sparse: slot in the sparse list) */
    int64_t *cell_coords;   /* S x 3 cell (x,y,z) */
    int64_t *cell_offsets;  /* S + 1 */
    int64_t n_points, n_cells;
} orc_its;

void orc_its_free(orc_its *t) {
    free(t->points); free(t->normals); free(t->edges); free(t->is_out);
    free(t->cell_indices); free(t->cell_coords); free(t->cell_offsets);
    memset(t, 0, sizeof(*t));
}

/* src/its.cu:172-268 for one intersection point p of a cell.
 * Float32, in the operation order AND fused-multiply-add contraction pattern that nvcc 12.9 (default
 * -fmad=true) emits for the reference's compute_normals_op (read off its SASS):
 *   a*(1-t) + b*t  ->  fmaf(a, 1-t, b*t)          (first product fused)
 * except the z-stage at t.z that the compiler shares between the x+- and y+- samples, where
 *   c00, c01       ->  fmaf(b, t, a*(1-t))        (second product fused);
 *   |g| = sqrtf(fmaf(gz,gz, fmaf(gx,gx, gy*gy)));  all divisions and the sqrt are IEEE-rounded. */
static inline float mix_a(float a, float b, float t, float omt) { return fmaf(a, omt, b * t); }
static inline float mix_b(float a, float b, float t, float omt) { return fmaf(b, t, a * omt); }
static inline float clamp01(float t) { return fmaxf(0.01f, fminf(0.99f, t)); }

static f3 cell_normal(f3 p, const f3 *cp, const float *v) {
    f3 cmin = cp[0];
    float sx = cp[7].x - cp[0].x, sy = cp[7].y - cp[0].y, sz = cp[7].z - cp[0].z;
    float tx = clamp01((p.x - cmin.x) / sx), ty = clamp01((p.y - cmin.y) / sy), tz = clamp01((p.z - cmin.z) / sz);
    const float eps = 0.02f;
    float xp = fminf(tx + eps, 0.99f), xm = fmaxf(tx - eps, 0.01f);
    float yp = fminf(ty + eps, 0.99f), ym = fmaxf(ty - eps, 0.01f);
    float zp = fminf(tz + eps, 0.99f), zm = fmaxf(tz - eps, 0.01f);
    float otx = 1 - tx, oty = 1 - ty, otz = 1 - tz;
    float c00 = mix_b(v[0], v[1], tz, otz), c01 = mix_b(v[2], v[3], tz, otz);
    float c10 = mix_a(v[4], v[5], tz, otz), c11 = mix_a(v[6], v[7], tz, otz);
    float c0 = mix_a(c00, c01, ty, oty), c1 = mix_a(c10, c11, ty, oty);
    float fxp = mix_a(c0, c1, xp, 1 - xp), fxm = mix_a(c0, c1, xm, 1 - xm);
    float gx = (fxp - fxm) / (sx * (xp - xm));
    float oyp = 1 - yp, oym = 1 - ym;
    float fyp = mix_a(mix_a(c00, c01, yp, oyp), mix_a(c10, c11, yp, oyp), tx, otx);
    float fym = mix_a(mix_a(c00, c01, ym, oym), mix_a(c10, c11, ym, oym), tx, otx);
    float gy = (fyp - fym) / (sy * (yp - ym));
    float ozp = 1 - zp, ozm = 1 - zm;
    float p0 = mix_a(mix_a(v[0], v[1], zp, ozp), mix_a(v[2], v[3], zp, ozp), ty, oty);
    float p1 = mix_a(mix_a(v[4], v[5], zp, ozp), mix_a(v[6], v[7], zp, ozp), ty, oty);
    float m0 = mix_a(mix_a(v[0], v[1], zm, ozm), mix_a(v[2], v[3], zm, ozm), ty, oty);
    float m1 = mix_a(mix_a(v[4], v[5], zm, ozm), mix_a(v[6], v[7], zm, ozm), ty, oty);
    float fzp = mix_a(p0, p1, tx, otx), fzm = mix_a(m0, m1, tx, otx);
    float gz = (fzp - fzm) / (sz * (zp - zm));
    float len = sqrtf(fmaf(gz, gz, fmaf(gx, gx, gy * gy)));
    f3 n;
    if (len > 1e-8f) { n.x = gx / len; n.y = gy / len; n.z = gz / len; }
    else { n.x = 0; n.y = 0; n.z = 1; }
    return n;
}

/* src/its.cu:93-159.  Dense when cell_idx == NULL (values = (X,Y,Z)); otherwise sparse
 * (values = (N,8), cell_idx = N sorted cell indices). */
int orc_get_intersection(const float *values, const int64_t *cell_idx, int64_t N, int64_t X, int64_t Y,
                         int64_t Z, const float *amin, const float *amax, float level,
                         int compute_normals, orc_its *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    int64_t ncells = cell_idx ? N : (X - 1) * (Y - 1) * (Z - 1);
    int64_t cap_c = 1024, cap_p = 4096, S = 0, I = 0;
    out->cell_indices = (int64_t *) malloc(cap_c * sizeof(int64_t));
    out->cell_coords = (int64_t *) malloc(cap_c * 3 * sizeof(int64_t));
    out->cell_offsets = (int64_t *) malloc((cap_c + 1) * sizeof(int64_t));
    out->points = (float *) malloc(cap_p * 3 * sizeof(float));
    out->normals = (float *) malloc(cap_p * 3 * sizeof(float));
    out->edges = (uint64_t *) malloc(cap_p * 2 * sizeof(uint64_t));
    out->is_out = (uint8_t *) malloc(cap_p);
    for (int64_t s = 0; s < ncells; s++) {
        int64_t i = cell_idx ? cell_idx[s] : s;
        int64_t z = i % (Z - 1); i /= (Z - 1);
        int64_t y = i % (Y - 1);
        int64_t x = i / (Y - 1);
        float cv[8];
        uint64_t pid[8];
        for (int k = 0; k < 8; k++) {
            int64_t px = x + (k >> 2 & 1), py = y + (k >> 1 & 1), pz = z + (k & 1);
            pid[k] = (uint64_t) ((px * Y + py) * Z + pz);
            cv[k] = cell_idx ? values[8 * s + k] : values[pid[k]];
        }
        int status = sign_change_edges(case_of(cv, level));
        if (status == 0) continue;
        if (S + 1 >= cap_c) {
            cap_c *= 2;
            out->cell_indices = (int64_t *) realloc(out->cell_indices, cap_c * sizeof(int64_t));
            out->cell_coords = (int64_t *) realloc(out->cell_coords, cap_c * 3 * sizeof(int64_t));
            out->cell_offsets = (int64_t *) realloc(out->cell_offsets, (cap_c + 1) * sizeof(int64_t));
        }
        if (I + 12 >= cap_p) {
            cap_p *= 2;
            out->points = (float *) realloc(out->points, cap_p * 3 * sizeof(float));
            out->normals = (float *) realloc(out->normals, cap_p * 3 * sizeof(float));
            out->edges = (uint64_t *) realloc(out->edges, cap_p * 2 * sizeof(uint64_t));
            out->is_out = (uint8_t *) realloc(out->is_out, cap_p);
        }
        out->cell_indices[S] = s;
        out->cell_coords[3 * S] = x; out->cell_coords[3 * S + 1] = y; out->cell_coords[3 * S + 2] = z;
        out->cell_offsets[S] = I;
        f3 cp[8];
        for (int k = 0; k < 8; k++) cp[k] = point_pos(&g, x + (k >> 2 & 1), y + (k >> 1 & 1), z + (k & 1));
        for (int e = 0; e < 12; e++) {
            if (!(status >> e & 1)) continue;
            int p0 = ORC_EDGES[e][0], p1 = ORC_EDGES[e][1];
            f3 p = edge_point(cp[p0], cp[p1], cv[p0], cv[p1], level);
            out->points[3 * I] = p.x; out->points[3 * I + 1] = p.y; out->points[3 * I + 2] = p.z;
            out->edges[2 * I] = pid[p0]; out->edges[2 * I + 1] = pid[p1];
            out->is_out[I] = cv[p0] <= cv[p1];
            f3 n = {0, 0, 0};
            if (compute_normals) n = cell_normal(p, cp, cv);
            out->normals[3 * I] = n.x; out->normals[3 * I + 1] = n.y; out->normals[3 * I + 2] = n.z;
            I++;
        }
        S++;
    }
    out->cell_offsets[S] = I;
    out->n_cells = S;
    out->n_points = I;
    return 0;
}

/* Symmetric 3x3 eigen-decomposition by cyclic Jacobi in double: A = V diag(w) V^T. */
static void jacobi3(double A[3][3], double V[3][3], double w[3]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = i == j;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (A[p][q] == 0.0) continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; k++) {
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = A[i][i];
}

typedef struct {
    orc_mesh mesh;       /* welded output exactly as src/dc.cu:204-217 */
    double *dual_v;      /* S x 3: clipped dual vertex per active cell (float64 ground truth)   */
    int64_t *quads;      /* Q x 4: slots into the active-cell list, already orientation-flipped */
    uint64_t *quad_edge; /* Q x 2: the (p_lo,p_hi) edge of each quad                            */
    int64_t n_quads, n_skipped;
} orc_dc;

void orc_dc_free(orc_dc *d) {
    orc_mesh_free(&d->mesh);
    free(d->dual_v); free(d->quads); free(d->quad_edge);
    memset(d, 0, sizeof(*d));
}

typedef struct { uint64_t lo, hi; uint8_t out; } edge_rec;
static int edge_cmp(const void *a, const void *b) {
    const edge_rec *x = (const edge_rec *) a, *y = (const edge_rec *) b;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
    return 0;
}
typedef struct { int64_t key, slot; } cell_rec;
static int cell_cmp(const void *a, const void *b) {
    int64_t x = ((const cell_rec *) a)->key, y = ((const cell_rec *) b)->key;
    return x < y ? -1 : x > y;
}

/* src/dc.cu:161-218: QEF in the reference's float32 arithmetic, solve of src/batched_la.cu:151-179 in float64.
 * `its` must carry normals.  Quads with a neighbour cell outside the cell grid (either side) or
 * absent from the active list are skipped (the reference only checks the lower side,
 * include/utils.cuh:128-135, and is undefined otherwise). */
int orc_dual_contouring(const orc_its *its, int64_t X, int64_t Y, int64_t Z, const float *amin,
                        const float *amax, float reg, float svd_tol, orc_dc *out) {
    orc_geom g = make_geom(X, Y, Z, amin, amax);
    int64_t S = its->n_cells;
    out->dual_v = (double *) malloc((size_t) (S > 0 ? S : 1) * 3 * sizeof(double));
    f3 *dvf = (f3 *) malloc((size_t) (S > 0 ? S : 1) * sizeof(f3));
    for (int64_t s = 0; s < S; s++) {
        /* src/dc.cu:27-64 in float32 with the reference's contraction pattern (SASS of get_qef_op):
         *   d = fmaf(n.z,p.z, fmaf(n.x,p.x, n.y*p.y));  ATA_ij = fmaf(n_i,n_j,ATA_ij);  ATb_i = fmaf(n_i,d,ATb_i)
         *   p_avg = sum(p)/float(k);  ATA_ii += reg;  ATb_i = fmaf(p_avg_i, reg, ATb_i) */
        float Af[3][3] = {{0}}, bf[3] = {0}, sum[3] = {0};
        int64_t i0 = its->cell_offsets[s], i1 = its->cell_offsets[s + 1];
        for (int64_t i = i0; i < i1; i++) {
            float n[3] = {its->normals[3 * i], its->normals[3 * i + 1], its->normals[3 * i + 2]};
            float p[3] = {its->points[3 * i], its->points[3 * i + 1], its->points[3 * i + 2]};
            for (int r = 0; r < 3; r++) sum[r] = sum[r] + p[r];
            float d = fmaf(n[2], p[2], fmaf(n[0], p[0], n[1] * p[1]));
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) Af[r][c] = fmaf(n[r], n[c], Af[r][c]);
                bf[r] = fmaf(n[r], d, bf[r]);
            }
        }
        float kf = (float) (uint32_t) (i1 - i0);
        double A[3][3], b[3];
        for (int r = 0; r < 3; r++) {
            float pavg = sum[r] / kf;
            Af[r][r] = Af[r][r] + reg;
            bf[r] = fmaf(pavg, reg, bf[r]);
        }
        for (int r = 0; r < 3; r++) { b[r] = bf[r]; for (int c = 0; c < 3; c++) A[r][c] = Af[r][c]; }
        /* pinv via eigen-decomposition (A symmetric PSD => singular values = eigenvalues) */
        double V[3][3], w[3];
        jacobi3(A, V, w);
        double wmax = fmax(w[0], fmax(w[1], w[2]));
        double xs[3] = {0, 0, 0};
        for (int j = 0; j < 3; j++) {
            if (!(w[j] > (double) svd_tol * wmax)) continue;
            double utb = V[0][j] * b[0] + V[1][j] * b[1] + V[2][j] * b[2];
            for (int r = 0; r < 3; r++) xs[r] += V[r][j] * utb / w[j];
        }
        /* src/dc.cu:93-98 clip to the cell AABB */
        int64_t cx = its->cell_coords[3 * s], cy = its->cell_coords[3 * s + 1], cz = its->cell_coords[3 * s + 2];
        f3 lo = point_pos(&g, cx, cy, cz), hi = point_pos(&g, cx + 1, cy + 1, cz + 1);
        xs[0] = fmin(fmax(xs[0], lo.x), hi.x);
        xs[1] = fmin(fmax(xs[1], lo.y), hi.y);
        xs[2] = fmin(fmax(xs[2], lo.z), hi.z);
        for (int r = 0; r < 3; r++) out->dual_v[3 * s + r] = xs[r];
        dvf[s].x = (float) xs[0]; dvf[s].y = (float) xs[1]; dvf[s].z = (float) xs[2];
    }
    /* unique edges sorted by (p_lo, p_hi): src/grid/uniform.cu:71-76 */
    int64_t I = its->n_points;
    edge_rec *er = (edge_rec *) malloc((size_t) (I > 0 ? I : 1) * sizeof(edge_rec));
    for (int64_t i = 0; i < I; i++) { er[i].lo = its->edges[2 * i]; er[i].hi = its->edges[2 * i + 1]; er[i].out = its->is_out[i]; }
    qsort(er, (size_t) I, sizeof(edge_rec), edge_cmp);
    int64_t E = 0;
    for (int64_t i = 0; i < I; i++)
        if (i == 0 || edge_cmp(&er[i], &er[E - 1]) != 0) er[E++] = er[i];
    /* cell key -> slot map (src/dc.cu:184-191), as a sorted array */
    cell_rec *cr = (cell_rec *) malloc((size_t) (S > 0 ? S : 1) * sizeof(cell_rec));
    for (int64_t s = 0; s < S; s++) {
        cr[s].key = (its->cell_coords[3 * s] * (Y - 1) + its->cell_coords[3 * s + 1]) * (Z - 1) + its->cell_coords[3 * s + 2];
        cr[s].slot = s;
    }
    qsort(cr, (size_t) S, sizeof(cell_rec), cell_cmp);
    /* src/shared_luts.cu:78-82 */
    static const int EN[3][4][3] = {{{0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {0, 0, 1}},
                                    {{0, 0, 0}, {0, 0, 1}, {1, 0, 1}, {1, 0, 0}},
                                    {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}}};
    out->quads = (int64_t *) malloc((size_t) (E > 0 ? E : 1) * 4 * sizeof(int64_t));
    out->quad_edge = (uint64_t *) malloc((size_t) (E > 0 ? E : 1) * 2 * sizeof(uint64_t));
    f3vec corners = {0};
    int64_t Q = 0, skipped = 0;
    for (int64_t e = 0; e < E; e++) {
        uint64_t lo = er[e].lo, d = er[e].hi - er[e].lo;
        int64_t pz = (int64_t) (lo % (uint64_t) Z), py = (int64_t) ((lo / (uint64_t) Z) % (uint64_t) Y), px = (int64_t) (lo / (uint64_t) (Y * Z));
        int axis = d == (uint64_t) (Y * Z) ? 0 : (d == (uint64_t) Z ? 1 : 2); /* include/utils.cuh:117-124 */
        int64_t q[4];
        int ok = 1;
        for (int k = 0; k < 4 && ok; k++) {
            int64_t cx = px - EN[axis][k][0], cy = py - EN[axis][k][1], cz = pz - EN[axis][k][2];
            if (cx < 0 || cy < 0 || cz < 0 || cx >= X - 1 || cy >= Y - 1 || cz >= Z - 1) { ok = 0; break; }
            int64_t key = (cx * (Y - 1) + cy) * (Z - 1) + cz;
            int64_t a = 0, bnd = S;
            while (a < bnd) { int64_t m = (a + bnd) >> 1; if (cr[m].key < key) a = m + 1; else bnd = m; }
            if (a == S || cr[a].key != key) { ok = 0; break; }
            q[k] = cr[a].slot;
        }
        if (!ok) { skipped++; continue; }
        if (!er[e].out) { int64_t t0 = q[0], t1 = q[1]; q[0] = q[3]; q[1] = q[2]; q[2] = t1; q[3] = t0; }
        memcpy(out->quads + 4 * Q, q, sizeof(q));
        out->quad_edge[2 * Q] = er[e].lo; out->quad_edge[2 * Q + 1] = er[e].hi;
        Q++;
        /* src/dc.cu:129-155 */
        f3 v0 = dvf[q[0]], v1 = dvf[q[1]], v2 = dvf[q[2]], v3 = dvf[q[3]];
        /* norm() of include/math.cuh as nvcc contracts it (SASS of get_triangles_op): sqrt(fma(dz,dz, fma(dx,dx, dy*dy))).
         * Pinned on the GPU: tests/test_dc_parity3_gpu.py feeds the reference's own dual vertices through the product's
         * split and obtains the reference's F bit for bit, and the product's split equals this one. */
        float dx02 = v0.x - v2.x, dy02 = v0.y - v2.y, dz02 = v0.z - v2.z;
        float dx13 = v1.x - v3.x, dy13 = v1.y - v3.y, dz13 = v1.z - v3.z;
        float d02 = sqrtf(fmaf(dz02, dz02, fmaf(dx02, dx02, dy02 * dy02)));
        float d13 = sqrtf(fmaf(dz13, dz13, fmaf(dx13, dx13, dy13 * dy13)));
        if (d02 > d13) {
            f3vec_push(&corners, v1); f3vec_push(&corners, v3); f3vec_push(&corners, v0);
            f3vec_push(&corners, v3); f3vec_push(&corners, v1); f3vec_push(&corners, v2);
        } else {
            f3vec_push(&corners, v2); f3vec_push(&corners, v0); f3vec_push(&corners, v1);
            f3vec_push(&corners, v0); f3vec_push(&corners, v2); f3vec_push(&corners, v3);
        }
    }
    out->n_quads = Q;
    out->n_skipped = skipped;
    out->mesh.n_active = S;
    weld(&corners, &out->mesh);
    free(corners.d); free(er); free(cr); free(dvf);
    return 0;
}
