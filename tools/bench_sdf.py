"""GPU: fused SDF evaluation -> extraction (ImplicitGrid) vs the two-step path (field resident in HBM), and the sparse
population from an SDF.  One JSON line per case; CUDA events, 3 warm-ups, median of 10."""
import json
import statistics
import sys

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch

import fields
import isoext_b200 as iso
from isoext_b200 import sdf as S


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


def case(n, name, sdf, dc=False):
    ig = iso.ImplicitGrid([n] * 3, sdf)
    t_fused, (v, f) = timed(lambda: iso.marching_cubes(ig))
    r = {"case": f"{n}^3 {name} marching_cubes", "fused_ms": t_fused, "fused_gvox_s": n ** 3 / t_fused / 1e6, "V": len(v), "T": len(f)}
    t_eval, _ = timed(lambda: ig.eval_slab(0, min(n, 256)), reps=5)
    r["materialize_ms_full_grid"] = t_eval * n / min(n, 256)
    if n ** 3 * 4 < 60e9:
        ug = ig.materialize()
        t_two, (uv, uf) = timed(lambda: iso.marching_cubes(ug))
        r.update(two_step_ms_field_resident=t_two, identical=bool(torch.equal(v.view(torch.int32), uv.view(torch.int32)) and torch.equal(f, uf)),
                 speedup_vs_resident_field=t_two / t_fused)
        if dc:
            t_dcf, (dv, df) = timed(lambda: iso.dual_contouring(ig))
            t_dct, (udv, udf) = timed(lambda: iso.dual_contouring(ug))
            r.update(dc_fused_ms=t_dcf, dc_two_step_ms=t_dct, dc_identical=bool(torch.equal(dv.view(torch.int32), udv.view(torch.int32)) and torch.equal(df, udf)))
        del ug
    print(json.dumps(r), flush=True)
    del ig
    torch.cuda.empty_cache()


def populate(n, name, sdf, chunk):
    ig = iso.ImplicitGrid([n] * 3, sdf)
    t, g = timed(lambda: iso.SparseGrid([n] * 3).populate_from_dense(ig, 0.0, x_chunk=chunk), warm=1, reps=3)
    print(json.dumps({"case": f"SparseGrid {n}^3-equivalent {name} band populated from the SDF program", "ms": t, "cells": g.get_num_cells(),
                      "Mcells_s": g.get_num_cells() / t / 1e3, "x_chunk": chunk}), flush=True)


if __name__ == "__main__":
    print(f"# {torch.cuda.get_device_name(0)}", flush=True)
    case(512, "torus", fields.torus(), dc=True)
    case(1024, "torus", fields.torus())
    case(1024, "quickstart (sphere minus 3 rotated tori)", fields.quickstart())
    case(2048, "CSG box-minus-sphere", fields.csg_box_minus_sphere())
    populate(1024, "sphere r=.7", S.SphereSDF(0.7), None)
    populate(4096, "sphere r=.7", S.SphereSDF(0.7), 130)
    populate(4096, "CSG box-minus-sphere", fields.csg_box_minus_sphere(), 130)
