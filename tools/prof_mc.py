"""Tiny driver for ncu: N iterations of dense MC at a given size/field."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
which = sys.argv[3] if len(sys.argv) > 3 else "torus"
fn = {"torus": fields.torus(), "csg": fields.csg_box_minus_sphere(), "sphere": fields.sphere()}[which]
vals = fields.eval_field(fn, (n, n, n)).cuda()
g = iso.UniformGrid([n] * 3); g.set_values(vals)
for _ in range(iters):
    v, f = iso.marching_cubes(g)
torch.cuda.synchronize()
print(len(v), len(f))
