"""Workload for ncu captures: one 1024^3 torus extraction from a resident field and one from the SDF program."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ig = iso.ImplicitGrid([n] * 3, fields.torus())
ug = ig.materialize()
for _ in range(3):
    iso.marching_cubes(ug); iso.marching_cubes(ig)
torch.cuda.synchronize()
