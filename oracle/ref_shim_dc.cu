// oracle/ref_shim_dc.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Exposes the reference's PER-CELL dual vertices (before the weld), its float32 QEF and the cuSOLVER
// status per cell, so that the dual-contouring parity test can be three-way and per cell
// (ours / the reference build / the float64 oracle) instead of a nearest-vertex distance.
//
// How: the reference keeps get_qef() and fix_dual_v_op in an anonymous namespace of src/dc.cu, so this
// translation unit compiles that UNMODIFIED file a second time, where it lies, inside a namespace of
// its own (every header it includes is `#pragma once` and is pulled in first, outside the namespace).
// The arithmetic executed is therefore the reference's own compiled code (same nvcc, same flags as the
// first copy in dc.o); only the six orchestration lines of src/dc.cu:166-182 are re-stated below so the
// intermediate arrays can be handed out.  tests/test_dc_parity3_gpu.py cross-checks that the welded set
// of these per-cell vertices is bit-identical to what ::dual_contouring returns.
#include "batched_la.cuh"
#include "dc.cuh"
#include "math.cuh"
#include "utils.cuh"

#include <thrust/device_vector.h>
#include <thrust/remove.h>
#include <thrust/sequence.h>

#include <string>
#include <tuple>

namespace ref_dc_tu {
#include "dc.cu"   // resolved through -I$(REF)/src
}

namespace {
thread_local std::string g_err_dc;
template <typename T> T *steal_dc(NDArray<T> &a) {
    if (a.size() == 0) return nullptr;
    T *p = a.data();
    a.read_only = true;
    return p;
}
}   // namespace

extern "C" {

const char *ref_dc_last_error() { return g_err_dc.c_str(); }

// Per active cell (the order of its.cell_indices): dual vertex after the clip (S x 3), before the clip (S x 3),
// the float32 QEF ATA (S x 9) / ATb (S x 3) and cuSOLVER's info (S).  Buffers are released with ref_free().
// `its` must carry normals.  Follows src/dc.cu:166-182.
int ref_dc_dual_vertices(void *g, void *its_p, float reg, float svd_tol, float **d_dual_v, float **d_unclipped,
                         float **d_ATA, float **d_ATb, int **d_info, size_t *n_cells) {
    try {
        Grid *grid = static_cast<Grid *>(g);
        const Intersection &its = *static_cast<Intersection *>(its_p);
        auto [ATA, ATb] = ref_dc_tu::get_qef(its, reg);
        // cuSOLVER gesvdj destroys its input matrix: hand lsq_svd copies so the QEF can be returned as well.
        NDArray<float> ATA_in(ATA);
        NDArray<float> ATb_in(ATb);
        BatchedLASolver solver;
        auto [dual_v, info] = solver.lsq_svd(ATA_in, ATb_in, svd_tol);
        NDArray<float> raw(dual_v);
        NDArray<uint> cells = grid->get_cells();
        NDArray<float3> points = grid->get_points();
        uint S = its.cell_indices.size();
        thrust::for_each(thrust::counting_iterator<uint>(0), thrust::counting_iterator<uint>(S),
                         ref_dc_tu::fix_dual_v_op(reinterpret_cast<float3 *>(dual_v.data()), its.cell_indices.data(),
                                                  points.data(), cells.data()));
        cells.free();
        points.free();
        cudaDeviceSynchronize();
        *n_cells = S;
        *d_dual_v = steal_dc(dual_v);
        *d_unclipped = steal_dc(raw);
        *d_ATA = steal_dc(ATA);
        *d_ATb = steal_dc(ATb);
        *d_info = steal_dc(info);
    } catch (const std::exception &e) {
        g_err_dc = e.what();
        return -1;
    }
    return 0;
}

}   // extern "C"
