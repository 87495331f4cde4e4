"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import sdf as S
from isoext_b200 import dist as idist

def dense(vals):
    g = iso.UniformGrid(list(vals.shape)); g.set_values(vals.cuda()); return g

for vals in (fields.eval_field(S.SphereSDF(0.5), (24, 24, 24)), fields.noise((9, 10, 130), 1), fields.eval_field(fields.csg_box_minus_sphere(), (20, 24, 128)),
             fields.eval_field(S.CuboidSDF([1, 1, 1]), (17, 17, 17)), torch.ones(6, 6, 6),
             fields.noise((3, 4, 1024), 7), fields.eval_field(S.SphereSDF(0.8), (3, 5, 4224)), fields.noise((5, 6, 384), 5)):   # span-summary paths
    g = dense(vals)
    for m in ("nagae", "lorensen"):
        for _ in range(2):                       # second call takes the single-sync fast path
            v, f = iso.marching_cubes(g, 0.0, m)
    its = iso.get_intersection(g, 0.0, True)
    v, f = iso.dual_contouring(g)
    g.get_points()
# big-bucket fallback (axis-aligned face larger than SEG_CAP inside one slab)
g = dense(fields.eval_field(fields.csg_box_minus_sphere(), (40, 160, 128)))
for _ in range(2): iso.marching_cubes(g)
iso.dual_contouring(g)
# second-level sort buckets (x-face inside one layer), the radix last resort (Z > 4096 on such a face) and a slanted sheet
ax = [fields.axis(n) for n in (6, 96, 96)]; X, Y, Z = torch.meshgrid(*ax, indexing="ij")
for vals in ((X - 0.13).contiguous(), (X + 0.013 * Y + 0.007 * Z - 0.05).contiguous()):
    g = dense(vals)
    for _ in range(3): iso.marching_cubes(g)
ax = [fields.axis(n) for n in (5, 3, 9000)]; X, Y, Z = torch.meshgrid(*ax, indexing="ij")
g = dense((X - 0.13).contiguous())
for _ in range(3): iso.marching_cubes(g)
# slabs
vals = fields.eval_field(S.CuboidSDF([1, 1, 1]), (33, 20, 24)).cuda()
for r in range(3):
    sg = idist.SlabGrid(list(vals.shape), rank=r, world=3); p = sg.plan
    sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1]); idist.marching_cubes_local(sg)
# sparse
sg = iso.SparseGrid([20, 20, 20]); sdf = S.SphereSDF(0.5)
c = sg.get_potential_cell_indices(20 ** 3)[0]; c = c[c < 19 ** 3]
keep = sg.filter_cell_indices(c, sdf(sg.get_points_by_cell_indices(c))); sg.add_cells(keep); sg.set_values(sdf(sg.get_points()))
iso.marching_cubes(sg); iso.dual_contouring(sg); iso.get_intersection(sg, 0.0, True)
torch.cuda.synchronize()
print("sanitize cases done")
