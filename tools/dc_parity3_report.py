"""GPU: per-cell three-way dual-contouring parity report (ours / reference CUDA build / float64 truth).

    python tools/dc_parity3_report.py [c4] [c5sphere] [c5csg] [small]  > profiles/r2_dc_parity3.jsonl

One JSON object per config: error histograms in units of a cell for the product and for the reference
build against the float64 minimiser of the reference's own float32 QEF, the worst reference cell with its
QEF, and the bit-equality flags of everything upstream / downstream of the solve (tests/dc_parity3.py)."""
import json
import sys

sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
import torch

import fields
import isoext_b200 as iso
from dc_parity3 import summary, three_way
from isoext_b200 import sdf as S
from oracle import ref


def dense(n, fn, tag):
    from bench_extra import fill_dense
    g = iso.UniformGrid([n] * 3)
    vals = fill_dense(g, fn, n)
    rg = ref.UniformGrid([n] * 3); rg.set_values(vals.contiguous())
    print(json.dumps(summary(three_way(iso, ref, g, rg, (n, n, n)), tag)), flush=True)


def band(n, fn, tag):
    from bench_extra import band_cells
    cells = band_cells(fn, n).to(torch.int32)
    g = iso.SparseGrid([n] * 3); g.add_cells(cells)
    N = g.get_num_cells()
    pts = g.get_points()
    vals = torch.empty((N, 8), dtype=torch.float32, device="cuda")
    for a in range(0, N, 1 << 21):
        vals[a:a + (1 << 21)] = fn(pts[a:a + (1 << 21)])
    del pts
    g.set_values(vals)
    rg = ref.SparseGrid([n] * 3); rg.add_cells(cells.contiguous()); rg.set_values(vals)
    print(json.dumps(summary(three_way(iso, ref, g, rg, (n, n, n)), tag)), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "c4", "c5sphere", "c5csg"]
    csg = fields.csg_box_minus_sphere()
    if "small" in which:
        dense(64, S.CuboidSDF([1, 1, 1]), "64^3 cuboid (sharp features)")
        dense(64, csg, "64^3 CSG")
        dense(128, S.SphereSDF(0.5), "128^3 sphere")
    if "c4" in which:
        dense(512, csg, "c4 512^3 CSG box-minus-sphere")
    if "c5sphere" in which:
        band(1024, S.SphereSDF(0.7), "c5 1024^3-equivalent sphere r=.7 band")
    if "c5csg" in which:
        band(1024, csg, "c5 1024^3-equivalent CSG band")
