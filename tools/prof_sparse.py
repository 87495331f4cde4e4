"""Workload for ncu captures: sparse MC + DC on the 1024^3-equivalent sphere band."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import isoext_b200 as iso
from isoext_b200 import sdf as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = iso.SparseGrid([n] * 3)
g.populate_from_dense(iso.ImplicitGrid([n] * 3, S.SphereSDF(0.7)))
for _ in range(2):
    iso.marching_cubes(g); iso.dual_contouring(g)
torch.cuda.synchronize()
