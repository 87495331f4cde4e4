"""Three-way, per-cell dual-contouring parity: ours / the reference's own CUDA build / float64 truth.

Used by tests/test_dc_parity3_gpu.py (asserts) and tools/dc_parity3_report.py (prints the table kept under
profiles/).  For one grid (dense UniformGrid or SparseGrid) it collects, PER ACTIVE CELL and in the same
order:

  ours      float32 dual vertex of the product (in-register FP64 Jacobi on the float32 QEF, csrc/dcmath.cuh)
  ref       float32 dual vertex of the reference's compiled code (src/dc.cu:166-182 via oracle/ref_shim_dc.cu:
            get_qef -> cuSOLVER gesvdjBatched + cuBLAS gemv -> clip), plus its float32 QEF (ATA, ATb)
  truth     float64 minimiser of the reference's OWN float32 QEF (numpy eigh, thresholded pseudo-inverse of
            src/batched_la.cu:158-170, clip of src/dc.cu:93-98) -- i.e. what an exact solver returns for the
            numbers the reference itself built
  oracle    oracle/oracle_c.c's float64 solve of its restated float32 QEF (optional; CPU, whole grid)

and the welded meshes.  Everything downstream of the solve is compared exactly by feeding the REFERENCE's
per-cell dual vertices through the product's quad / split / weld stage: the result must be the reference's
(V, F) bit for bit.
"""
from __future__ import annotations

import numpy as np
import torch


def _cell_bounds(cell_idx: np.ndarray, shape, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)):
    """float32 AABB of each cell, include/utils.cuh:62-80 (pos = fma(i/(n-1), size, min); exact for [-1,1])."""
    X, Y, Z = shape
    c = cell_idx.astype(np.int64)
    cz = c % (Z - 1)
    cy = (c // (Z - 1)) % (Y - 1)
    cx = c // ((Z - 1) * (Y - 1))
    lo, hi = np.empty((len(c), 3), np.float32), np.empty((len(c), 3), np.float32)
    for a, (ci, n) in enumerate(((cx, X), (cy, Y), (cz, Z))):
        size = np.float32(aabb_max[a]) - np.float32(aabb_min[a])
        for out, i in ((lo, ci), (hi, ci + 1)):
            q = (i.astype(np.float32) / np.float32(n - 1)).astype(np.float32)
            out[:, a] = (q.astype(np.float64) * np.float64(size) + np.float64(np.float32(aabb_min[a]))).astype(np.float32)
    return lo, hi, np.stack([cx, cy, cz], 1)


def f64_solve(ATA: np.ndarray, ATb: np.ndarray, svd_tol: float, lo: np.ndarray, hi: np.ndarray):
    """Exact-arithmetic stand-in for src/batched_la.cu:151-179 + src/dc.cu:93-98 on float32 inputs.
    Returns (clipped, unclipped, eigenvalues descending)."""
    A = ATA.astype(np.float64)
    A = 0.5 * (A + A.transpose(0, 2, 1))        # the reference's ATA is symmetric up to float32 rounding of n_i*n_j == n_j*n_i (exact)
    w, V = np.linalg.eigh(A)                    # ascending
    b = ATb.astype(np.float64)
    utb = np.einsum("sij,si->sj", V, b)
    keep = w > svd_tol * w[:, -1:]
    coef = np.where(keep, utb / np.where(keep, w, 1.0), 0.0)
    x = np.einsum("sij,sj->si", V, coef)
    return np.minimum(np.maximum(x, lo), hi), x, w[:, ::-1]


def three_way(iso, ref, grid, rgrid, shape, level=0.0, reg=1e-2, svd_tol=1e-6, with_oracle=False, oracle_inputs=None):
    """grid / rgrid: the product's and the reference build's grid objects holding the SAME values.
    Returns a dict of numpy arrays + scalars (see keys below)."""
    from isoext_b200.dc import dc_dense_raw
    from isoext_b200.sparse import SparseGrid, dc_sparse_raw
    sparse = isinstance(grid, SparseGrid)
    raw_fn = dc_sparse_raw if sparse else dc_dense_raw

    its = iso.get_intersection(grid, level, compute_normals=True)
    rits = ref.get_intersection(rgrid, level, True)
    out = {}
    out["its_points_equal"] = bool(torch.equal(rits.get_points().view(torch.int32), its.get_points().view(torch.int32)))
    out["its_normals_equal"] = bool(torch.equal(rits.get_normals().view(torch.int32), its.get_normals().view(torch.int32)))
    out["cells_equal"] = bool(torch.equal(rits.get_cell_indices().to(torch.int64) & 0xFFFFFFFF, its.cell_indices.to(torch.int64)))

    v, f, dual_v, quads = raw_fn(grid, its, reg, svd_tol, want_quads=True)
    r = ref.dc_dual_vertices(rgrid, rits, reg, svd_tol)
    rv, rf = ref.dual_contouring(rgrid, level, rits, reg, svd_tol)
    # downstream of the solve: reference dual vertices through OUR quad / diagonal split / weld
    v2, f2, _, _ = raw_fn(grid, its, reg, svd_tol, dual_v_in=r["dual_v"])
    out["downstream_V_equal"] = bool(v2.shape == rv.shape and torch.equal(v2.view(torch.int32), rv.view(torch.int32)))
    out["downstream_F_equal"] = bool(f2.shape == rf.shape and torch.equal(f2, rf))
    out["n_V_ours"], out["n_V_ref"], out["n_F_ours"], out["n_F_ref"] = len(v), len(rv), len(f), len(rf)

    cell_idx = its.cell_indices.cpu().numpy()
    if sparse:   # its.cell_indices are slots into the active list
        cell_idx = grid._cells.cpu().numpy()[cell_idx]
    lo, hi, coords = _cell_bounds(cell_idx, shape, grid.aabb_min, grid.aabb_max)
    ATA, ATb = r["ATA"].cpu().numpy(), r["ATb"].cpu().numpy()
    truth, truth_raw, sig = f64_solve(ATA, ATb, svd_tol, lo, hi)
    cell = float(max((hi - lo).max(), 0))
    ours = dual_v.cpu().numpy().astype(np.float64)
    refv = r["dual_v"].cpu().numpy().astype(np.float64)
    out.update(cell=cell, n_cells=len(cell_idx), coords=coords, sig=sig, ATA=ATA, ATb=ATb, truth=truth, truth_raw=truth_raw,
               ours=ours, ref=refv, ref_raw=r["raw"].cpu().numpy(), info=r["info"].cpu().numpy(),
               err_ours=np.abs(ours - truth).max(1) / cell, err_ref=np.abs(refv - truth).max(1) / cell,
               ours_vs_ref=np.abs(ours - refv).max(1) / cell,
               quads=quads.cpu().numpy(), v=v, f=f, rv=rv, rf=rf)
    if with_oracle:
        import oracle
        vals, cells = oracle_inputs
        o = oracle.get_intersection(vals, shape=shape if sparse else None, cell_idx=cells, level=level, compute_normals=True)
        od = oracle.dual_contouring(o, shape, reg, svd_tol, grid.aabb_min, grid.aabb_max)
        out["oracle"] = od["dual_v"]
        out["err_oracle_vs_truth"] = np.abs(od["dual_v"] - truth).max(1) / cell
        out["err_ours_vs_oracle"] = np.abs(ours - od["dual_v"]).max(1) / cell
        out["oracle_quads_equal"] = bool(np.array_equal(od["quads"], out["quads"].astype(np.int64)))
        # the oracle's split / weld (src/dc.cu:139-155, src/utils.cu:32-59) vs ours on the ORACLE's dual vertices
        od32 = torch.from_numpy(od["dual_v"].astype(np.float32)).to(dual_v.device)
        v3, f3, _, _ = raw_fn(grid, its, reg, svd_tol, dual_v_in=od32)
        out["oracle_downstream_equal"] = bool(
            od["v"].shape == tuple(v3.shape) and np.array_equal(od["v"].view(np.uint32), v3.cpu().numpy().view(np.uint32))
            and od["f"].shape == tuple(f3.shape) and np.array_equal(od["f"], f3.cpu().numpy()))
        out["dual_v_bits_differ_from_oracle"] = int((od32.view(torch.int32) != dual_v.view(torch.int32)).any(dim=1).sum())
        out["oracle_F_equal"] = bool(od["f"].shape == tuple(f.shape) and np.array_equal(od["f"], f.cpu().numpy()))
        out["oracle_V_equal"] = bool(od["v"].shape == tuple(v.shape) and np.array_equal(od["v"].view(np.uint32), v.cpu().numpy().view(np.uint32)))
    return out


def summary(res, tag):
    """JSON-able digest: error histograms (in cells) of ours and of the reference against the float64 truth,
    the worst reference cell with its QEF, and the exact-equality flags."""
    eo, er = res["err_ours"], res["err_ref"]
    edges = [0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 0.5, 10]
    hist = lambda e: {f"<{edges[i + 1]:g}": int(((e >= edges[i]) & (e < edges[i + 1])).sum()) for i in range(len(edges) - 1)}
    w = int(np.argmax(er))
    cond = res["sig"][:, 0] / np.maximum(res["sig"][:, 2], 1e-300)
    big = er > 1e-2
    d = {
        "case": tag, "cell_size": res["cell"], "active_cells": int(res["n_cells"]),
        "its_points_equal": res["its_points_equal"], "its_normals_equal": res["its_normals_equal"], "cells_equal": res["cells_equal"],
        "downstream_of_solve_V_bit_equal_to_reference": res["downstream_V_equal"],
        "downstream_of_solve_F_equal_to_reference": res["downstream_F_equal"],
        "V_ours": res["n_V_ours"], "V_ref": res["n_V_ref"], "F_ours": res["n_F_ours"], "F_ref": res["n_F_ref"],
        "ours_vs_f64_max_cells": float(eo.max()), "ours_vs_f64_hist_cells": hist(eo),
        "ours_vs_ref_max_cells": float(res["ours_vs_ref"].max()), "ours_vs_ref_cells_beyond_1e-4": int((res["ours_vs_ref"] > 1e-4).sum()),
        "ref_vs_f64_max_cells": float(er.max()), "ref_vs_f64_p50": float(np.percentile(er, 50)),
        "ref_vs_f64_p99": float(np.percentile(er, 99)), "ref_vs_f64_p999": float(np.percentile(er, 99.9)),
        "ref_vs_f64_hist_cells": hist(er), "ref_cells_beyond_1e-4": int((er > 1e-4).sum()),
        "gesvdj_info_nonzero": int((res["info"] != 0).sum()),
        "median_cond_of_cells_with_ref_err_gt_1e-2": float(np.median(cond[big])) if big.any() else None,
        "median_cond_all": float(np.median(cond)),
        "worst_ref_cell": {
            "coords": res["coords"][w].tolist(), "ATA": res["ATA"][w].tolist(), "ATb": res["ATb"][w].tolist(),
            "singular_values_f64": res["sig"][w].tolist(), "f64_unclipped": res["truth_raw"][w].tolist(),
            "ref_unclipped": res["ref_raw"][w].tolist(), "f64": res["truth"][w].tolist(), "ref": res["ref"][w].tolist(),
            "ours": res["ours"][w].tolist(), "gesvdj_info": int(res["info"][w]),
            "residual_f64": float(np.linalg.norm(res["ATA"][w].astype(np.float64) @ res["truth_raw"][w] - res["ATb"][w])),
            "residual_ref": float(np.linalg.norm(res["ATA"][w].astype(np.float64) @ res["ref_raw"][w].astype(np.float64) - res["ATb"][w])),
        },
    }
    for k in ("oracle_quads_equal", "oracle_F_equal", "oracle_V_equal", "oracle_downstream_equal", "dual_v_bits_differ_from_oracle"):
        if k in res:
            d[k] = res[k]
    if "err_ours_vs_oracle" in res:
        d["ours_vs_oracle_max_cells"] = float(res["err_ours_vs_oracle"].max())
        d["oracle_vs_f64_of_ref_qef_max_cells"] = float(res["err_oracle_vs_truth"].max())
    return d
