"""GPU: the BASELINE.json configs that are not bench.py's headline line, ours vs the reference CUDA build.

    python tools/bench_extra.py [c1] [c4] [c5] [c5big] [its]      (default: c1 c4 c5)

  c1     64^3 sphere marching_cubes (latency case: us per call)
  c4     512^3 box-minus-sphere dual_contouring (defaults: reg=1e-2, svd_tol=1e-6, automatic normals)
  c5     SparseGrid narrow band at 1024^3 (sphere r=.7 and the c3 CSG): marching_cubes + dual_contouring
  c5big  the same at 4096^3-equivalent (int64 cell ids; the reference cannot represent it) -- ours only
Prints one JSON object per case; timing = CUDA events round the public call, field/band resident,
3 warm-ups, median of 10 (reference: 1 warm-up, median of 3).  The reference arm needs oracle/_ref."""
import json
import statistics
import sys

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch

import fields
import isoext_b200 as iso
from isoext_b200 import sdf as S

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.exists("MEASURED_PEAKS.json") else 6650.0
try:
    from oracle import ref
    HAVE_REF = ref.available()
except Exception:
    ref, HAVE_REF = None, False


def timed(fn, warm=3, reps=10):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), out


def fill_dense(grid, fn, n):
    ax = fields.axis(n).cuda(); view = grid.values_view()
    for a in range(0, n, 16):
        P = torch.stack(torch.meshgrid(ax[a:a + 16], ax, ax, indexing="ij"), dim=-1); view[a:a + 16] = fn(P); del P
    return view


def band_cells(fn, n, slab=8):
    """Cell ids of all cells with a sign change (case not in {0,255}) of fn on the n^3 grid -- the narrow band
    the reference's population recipe would end with, computed directly (bench helper, torch)."""
    ax = fields.axis(n).cuda()
    out = []
    C1 = n - 1
    for a in range(0, C1, slab):
        b = min(C1, a + slab)
        P = torch.stack(torch.meshgrid(ax[a:b + 1], ax, ax, indexing="ij"), dim=-1)
        neg = fn(P) < 0
        del P
        cnt = torch.zeros((b - a, C1, C1), dtype=torch.int8, device="cuda")
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    cnt += neg[dx:dx + b - a, dy:dy + C1, dz:dz + C1]
        idx = ((cnt > 0) & (cnt < 8)).nonzero()
        out.append(((idx[:, 0] + a) * C1 + idx[:, 1]) * C1 + idx[:, 2])
        del neg, cnt, idx
    return torch.cat(out)


def nn_dist(A, B, h):
    """max over a in A of the distance to the nearest b in B (searching the 27 bins of size h round a;
    inf if none) -- welded DC vertex sets of two solvers differ by solver noise, which also permutes the
    position-sorted ids, so full-size DC parity is stated as a two-sided nearest-vertex distance."""
    M = int(4.0 / h) + 4
    binof = lambda P: torch.floor((P.double() + 2.0) / h).long()
    key = lambda b: (b[:, 0] * M + b[:, 1]) * M + b[:, 2]
    kb, order = torch.sort(key(binof(B)))
    Bs = B[order].double()
    ba, best = binof(A), torch.full((len(A),), float("inf"), dtype=torch.float64, device=A.device)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                ka = key(ba + torch.tensor([dx, dy, dz], device=A.device))
                lo = torch.searchsorted(kb, ka)
                for j in range(6):
                    idx = (lo + j).clamp(max=len(kb) - 1)
                    ok = kb[idx] == ka
                    d = (A.double() - Bs[idx]).norm(dim=-1)
                    best = torch.where(ok & (d < best), d, best)
    return float(best.max())


def line(**kw):
    print(json.dumps(kw), flush=True)


def case_c1():
    n = 64
    g = iso.UniformGrid([n] * 3); fill_dense(g, fields.sphere(0.5), n)
    ms, (v, f) = timed(lambda: iso.marching_cubes(g), reps=50)
    r = {"case": "c1 64^3 sphere marching_cubes", "ours_us": ms * 1e3, "V": len(v), "T": len(f)}
    if HAVE_REF:
        rg = ref.UniformGrid([n] * 3); rg.set_values(g.values_view().contiguous())
        rms, (rv, rf) = timed(lambda: ref.marching_cubes(rg), warm=2, reps=10)
        r.update(ref_us=rms * 1e3, speedup=rms / ms, identical=bool(torch.equal(rv.view(torch.int32), v.view(torch.int32)) and torch.equal(rf, f)))
    line(**r)


def case_c4():
    n = 512
    g = iso.UniformGrid([n] * 3); fill_dense(g, fields.csg_box_minus_sphere(), n)
    ms, (v, f) = timed(lambda: iso.dual_contouring(g))
    ms_mc, _ = timed(lambda: iso.marching_cubes(g))
    balg = 4.0 * n ** 3 + 12.0 * len(v) + 12.0 * len(f)
    r = {"case": "c4 512^3 box-minus-sphere dual_contouring", "ours_ms": ms, "gvox_s": n ** 3 / ms / 1e6,
         "whole_path_frac_of_hbm_peak": balg / (ms * 1e-3) / 1e9 / PEAK, "V": len(v), "T": len(f), "mc_same_field_ms": ms_mc}
    if HAVE_REF:
        rg = ref.UniformGrid([n] * 3); rg.set_values(g.values_view().contiguous())
        rms, (rv, rf) = timed(lambda: ref.dual_contouring(rg), warm=1, reps=3)
        h = 2.0 / (n - 1)
        r.update(ref_ms=rms, speedup=rms / ms, same_counts=bool(len(rv) == len(v) and len(rf) == len(f)),
                 nearest_vertex_dist_ours_to_ref_in_cells=nn_dist(v, rv, h) / h,
                 nearest_vertex_dist_ref_to_ours_in_cells=nn_dist(rv, v, h) / h,
                 note="ids are position-sorted, so solver noise permutes them; quad topology parity is tested in tests/test_dc_gpu.py")
        del rv, rf
    line(**r)


def sparse_case(tag, n, name, fn, with_ref):
    cells = band_cells(fn, n)
    g = iso.SparseGrid([n] * 3)
    g.add_cells(cells if n > 1290 else cells.to(torch.int32))
    N = g.get_num_cells()
    pts = g.get_points()
    vals = torch.empty((N, 8), dtype=torch.float32, device="cuda")
    for a in range(0, N, 1 << 22):
        vals[a:a + (1 << 22)] = fn(pts[a:a + (1 << 22)])
    del pts
    g.set_values(vals)
    idx_bytes = 8 if g.get_cell_indices().dtype == torch.int64 else 4
    ms, (v, f) = timed(lambda: iso.marching_cubes(g))
    r = {"case": f"{tag} SparseGrid {n}^3-equivalent {name} band", "N_cells": N, "mc_ms": ms, "mc_Mcells_s": N / ms / 1e3,
         "mc_equiv_gvox_s": float(n) ** 3 / ms / 1e6, "V": len(v), "T": len(f),
         "mc_frac_of_hbm_peak": (N * (32 + idx_bytes) + 12.0 * len(v) + 12.0 * len(f)) / (ms * 1e-3) / 1e9 / PEAK}
    dms, (dv, df) = timed(lambda: iso.dual_contouring(g))
    r.update(dc_ms=dms, dc_Mcells_s=N / dms / 1e3, dc_V=len(dv), dc_T=len(df))
    if with_ref and HAVE_REF:
        rg = ref.SparseGrid([n] * 3); rg.add_cells(cells.to(torch.int32).contiguous()); rg.set_values(vals)
        rms, (rv, rf) = timed(lambda: ref.marching_cubes(rg), warm=1, reps=3)
        r.update(ref_mc_ms=rms, mc_speedup=rms / ms,
                 mc_identical=bool(len(rv) == len(v) and torch.equal(rv.view(torch.int32), v.view(torch.int32)) and torch.equal(rf, f)))
        del rv, rf
        rdms, (rdv, rdf) = timed(lambda: ref.dual_contouring(rg), warm=1, reps=3)
        h = 2.0 / (n - 1)
        r.update(ref_dc_ms=rdms, dc_speedup=rdms / dms, dc_same_counts=bool(len(rdv) == len(dv) and len(rdf) == len(df)),
                 dc_nearest_vertex_dist_in_cells=max(nn_dist(dv, rdv, h), nn_dist(rdv, dv, h)) / h)
    line(**r)


def case_c5(big=False):
    n = 4096 if big else 1024
    for name, fn in (("sphere r=.7", S.SphereSDF(0.7)), ("CSG box-minus-sphere", fields.csg_box_minus_sphere())):
        sparse_case("c5big" if big else "c5", n, name, fn, with_ref=not big)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c4", "c5"]
    print(f"# {torch.cuda.get_device_name(0)}; reference build available: {HAVE_REF}; HBM peak used {PEAK} GB/s", flush=True)
    for w in which:
        {"c1": case_c1, "c4": case_c4, "c5": case_c5, "c5big": lambda: case_c5(True)}[w]()
