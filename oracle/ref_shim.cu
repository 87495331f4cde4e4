// oracle/ref_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI shim over the *unmodified* reference sources of GuangyanCai/isoext, compiled where
// they lie under /root/reference (see oracle/Makefile; outputs go to oracle/_ref/, git-ignored).
// It stands in for the reference's nanobind module (src/isoext_ext.cu), which cannot be built
// here because nanobind is not installed.  Only tests/, __graft_entry__.smoke() and
// bench.py --impl reference may load the resulting library; the product never does.
//
// Ownership mirrors src/isoext_ext.cu:32-76: result buffers are handed out without a copy and
// released by the caller with ref_free() (cudaFree), exactly like the reference's DLPack capsule.
#include "dc.cuh"
#include "grid/sparse.cuh"
#include "grid/uniform.cuh"
#include "its.cuh"
#include "mc/mc.cuh"
#include "ndarray.cuh"

#include <cstring>
#include <optional>
#include <string>

namespace {
thread_local std::string g_err;

template <typename T> T *steal(NDArray<T> &a) {
    // Equivalent of ours_to_nb(): keep the allocation alive past the NDArray destructor.
    if (a.size() == 0) return nullptr;
    T *p = a.data();
    a.read_only = true;
    return p;
}
inline float3 f3(const float *p) { return make_float3(p[0], p[1], p[2]); }
}   // namespace

#define REF_TRY try {
#define REF_CATCH                                                                             \
    }                                                                                         \
    catch (const std::exception &e) {                                                         \
        g_err = e.what();                                                                     \
        return -1;                                                                            \
    }                                                                                         \
    return 0;

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
void ref_free(void *p) { if (p) cudaFree(p); }
int ref_sync() { return (int) cudaDeviceSynchronize(); }

// ---- grids -----------------------------------------------------------------------------
int ref_uniform_new(unsigned X, unsigned Y, unsigned Z, const float *amin, const float *amax,
                    float default_value, void **out) {
    REF_TRY
    *out = new UniformGrid(make_uint3(X, Y, Z), f3(amin), f3(amax), default_value);
    REF_CATCH
}
int ref_sparse_new(unsigned X, unsigned Y, unsigned Z, const float *amin, const float *amax,
                   float default_value, void **out) {
    REF_TRY
    *out = new SparseGrid(make_uint3(X, Y, Z), f3(amin), f3(amax), default_value);
    REF_CATCH
}
void ref_grid_delete(void *g) { delete static_cast<Grid *>(g); }
unsigned ref_grid_num_cells(void *g) { return static_cast<Grid *>(g)->get_num_cells(); }
unsigned ref_grid_num_points(void *g) { return static_cast<Grid *>(g)->get_num_points(); }

// values: device pointer; shape is (X,Y,Z) for uniform and (N,8) for sparse.
int ref_uniform_set_values(void *g, const float *d_values, size_t X, size_t Y, size_t Z) {
    REF_TRY
    NDArray<float> v(const_cast<float *>(d_values), {X, Y, Z});
    static_cast<UniformGrid *>(g)->set_values(v);
    REF_CATCH
}
int ref_sparse_set_values(void *g, const float *d_values, size_t N) {
    REF_TRY
    NDArray<float> v(const_cast<float *>(d_values), {N, 8});
    static_cast<SparseGrid *>(g)->set_values(v);
    REF_CATCH
}
// Returns a fresh device buffer of num_points float3.
int ref_grid_points(void *g, float **d_points, size_t *n) {
    REF_TRY
    NDArray<float3> p = static_cast<Grid *>(g)->get_points();
    *n = p.size();
    *d_points = reinterpret_cast<float *>(steal(p));
    REF_CATCH
}
// get_cells (src/grid/uniform.cu:42-51 / src/grid/sparse.cu:61-69): fresh device buffer of num_cells * 8 uint
int ref_grid_cells(void *g, unsigned **d_cells, size_t *n) {
    REF_TRY
    NDArray<uint> c = static_cast<Grid *>(g)->get_cells();
    *n = c.size();
    *d_cells = steal(c);
    REF_CATCH
}
int ref_grid_values(void *g, float **d_values, size_t *n) {
    REF_TRY
    NDArray<float> v = static_cast<Grid *>(g)->get_values();
    *n = v.size();
    *d_values = steal(v);
    REF_CATCH
}
int ref_sparse_add_cells(void *g, const int *d_idx, size_t n) {
    REF_TRY
    NDArray<int> a(const_cast<int *>(d_idx), {n});
    NDArray<uint> u = a.cast<uint>();
    static_cast<SparseGrid *>(g)->add_cells(u);
    REF_CATCH
}
int ref_sparse_remove_cells(void *g, const int *d_idx, size_t n) {
    REF_TRY
    NDArray<int> a(const_cast<int *>(d_idx), {n});
    NDArray<uint> u = a.cast<uint>();
    static_cast<SparseGrid *>(g)->remove_cells(u);
    REF_CATCH
}
int ref_sparse_cell_indices(void *g, int **d_idx, size_t *n) {
    REF_TRY
    thrust::device_vector<uint> dv = static_cast<SparseGrid *>(g)->get_cell_indices();
    NDArray<int> r = NDArray<uint>::copy(dv.data().get(), {dv.size()}).cast<int>();
    *n = r.size();
    *d_idx = steal(r);
    REF_CATCH
}
int ref_sparse_points_by_cell_indices(void *g, const int *d_idx, size_t n, float **d_points) {
    REF_TRY
    NDArray<int> a(const_cast<int *>(d_idx), {n});
    NDArray<uint> u = a.cast<uint>();
    NDArray<float3> p = static_cast<SparseGrid *>(g)->get_points_by_cell_indices(u);
    *d_points = reinterpret_cast<float *>(steal(p));
    REF_CATCH
}
int ref_sparse_filter_cell_indices(void *g, const int *d_idx, const float *d_values, size_t n,
                                   float level, int **d_out, size_t *n_out) {
    REF_TRY
    NDArray<int> a(const_cast<int *>(d_idx), {n});
    NDArray<uint> u = a.cast<uint>();
    NDArray<float> v(const_cast<float *>(d_values), {n, 8});
    NDArray<int> r = static_cast<SparseGrid *>(g)->filter_cell_indices(u, v, level).cast<int>();
    *n_out = r.size();
    *d_out = steal(r);
    REF_CATCH
}

// ---- marching cubes (src/mc/mc.cu:17-68) -------------------------------------------------
int ref_marching_cubes(void *g, float level, const char *method, float **d_v, size_t *nv,
                       int **d_f, size_t *nf) {
    REF_TRY
    auto [v, f] = mc::marching_cubes(static_cast<Grid *>(g), level, std::string(method));
    *nv = v.size();
    *nf = f.size() / 3;
    *d_v = reinterpret_cast<float *>(steal(v));
    *d_f = steal(f);
    REF_CATCH
}

// ---- intersections (src/its.cu:93-159) ----------------------------------------------------
int ref_get_intersection(void *g, float level, int compute_normals, void **out) {
    REF_TRY
    *out = new Intersection(get_intersection(static_cast<Grid *>(g), level, compute_normals != 0));
    REF_CATCH
}
void ref_its_delete(void *its) { delete static_cast<Intersection *>(its); }
size_t ref_its_num_points(void *its) { return static_cast<Intersection *>(its)->points.size(); }
size_t ref_its_num_cells(void *its) { return static_cast<Intersection *>(its)->cell_indices.size(); }
int ref_its_has_normals(void *its) { return static_cast<Intersection *>(its)->has_normals(); }
// Borrowed device pointers (owned by the Intersection).
const float *ref_its_points(void *its) { return reinterpret_cast<const float *>(static_cast<Intersection *>(its)->points.data()); }
const float *ref_its_normals(void *its) { return reinterpret_cast<const float *>(static_cast<Intersection *>(its)->normals.data()); }
const unsigned *ref_its_edges(void *its) { return reinterpret_cast<const unsigned *>(static_cast<Intersection *>(its)->edges.data()); }
const unsigned *ref_its_cell_indices(void *its) { return static_cast<Intersection *>(its)->cell_indices.data(); }
const unsigned *ref_its_cell_offsets(void *its) { return static_cast<Intersection *>(its)->cell_offsets.data(); }
const bool *ref_its_is_out(void *its) { return static_cast<Intersection *>(its)->is_out.data(); }
int ref_its_set_normals(void *its, const float *d_normals, size_t n) {
    REF_TRY
    NDArray<float3> nn(reinterpret_cast<float3 *>(const_cast<float *>(d_normals)), {n});
    static_cast<Intersection *>(its)->set_normals(nn);
    REF_CATCH
}

// ---- dual contouring (src/isoext_ext.cu:345-361 + src/dc.cu:161-218) ----------------------
int ref_dual_contouring(void *g, float level, void *its_or_null, float reg, float svd_tol,
                        float **d_v, size_t *nv, int **d_f, size_t *nf) {
    REF_TRY
    Grid *grid = static_cast<Grid *>(g);
    // The binding takes the Intersection by value (std::optional): the caller's object is untouched.
    Intersection its = its_or_null ? Intersection(*static_cast<Intersection *>(its_or_null))
                                   : get_intersection(grid, level, true);
    if (!its.has_normals()) {
        compute_intersection_normals(its, grid);
        its._has_normals = true;
    }
    auto [v, f] = dual_contouring(grid, its, level, reg, svd_tol);
    *nv = v.size();
    *nf = f.size() / 3;
    *d_v = reinterpret_cast<float *>(steal(v));
    *d_f = steal(f);
    REF_CATCH
}

}   // extern "C"
