"""Generate tests/golden/*.npz from the reference's OWN CUDA build (oracle/_ref, needs a GPU):

    gpurun -- 'python tools/make_golden.py gpurun_out/golden'   then copy gpurun_out/golden/* to tests/golden/

Each fixture holds a small input field and the reference's outputs on it, so that the CPU oracle
(and through it the product) stays pinned to real reference outputs even where no GPU / no
/root/reference is available."""
import sys
from pathlib import Path

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch

import fields
from isoext_b200 import sdf as S
from oracle import ref

out = Path(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
out.mkdir(parents=True, exist_ok=True)

MC_CASES = {
    "sphere16_nagae": (fields.eval_field(S.SphereSDF(0.5), (16, 16, 16)), 0.0, "nagae", (-1, -1, -1), (1, 1, 1)),
    "sphere16_lorensen_lvl": (fields.eval_field(S.SphereSDF(0.5), (16, 16, 16)), 0.1, "lorensen", (-1, -1, -1), (1, 1, 1)),
    "torus_12x20x16": (fields.eval_field(fields.torus(), (12, 20, 16)), 0.0, "nagae", (-1, -1, -1), (1, 1, 1)),
    "cuboid17_exact_hits": (fields.eval_field(S.CuboidSDF([1, 1, 1]), (17, 17, 17)), 0.0, "nagae", (-1, -1, -1), (1, 1, 1)),
    "cuboid17_exact_hits_lorensen": (fields.eval_field(S.CuboidSDF([1, 1, 1]), (17, 17, 17)), 0.0, "lorensen", (-1, -1, -1), (1, 1, 1)),
    "noise10_nagae": (fields.noise((10, 10, 10), 7), 0.0, "nagae", (-1, -1, -1), (1, 1, 1)),
    "noise10_lorensen": (fields.noise((10, 10, 10), 8), 0.0, "lorensen", (0, -2, 5), (3, 1, 6)),
    "csg20": (fields.eval_field(fields.csg_box_minus_sphere(), (20, 20, 20)), 0.0, "nagae", (-1, -1, -1), (1, 1, 1)),
}
for name, (vals, level, method, lo, hi) in MC_CASES.items():
    g = ref.UniformGrid(list(vals.shape), lo, hi)
    g.set_values(vals.cuda())
    v, f = ref.marching_cubes(g, level, method)
    np.savez_compressed(out / f"mc_{name}.npz", values=vals.numpy(), level=np.float32(level), method=method,
                        aabb_min=np.float32(lo), aabb_max=np.float32(hi), v=v.cpu().numpy(), f=f.cpu().numpy())
    print("mc", name, tuple(v.shape), tuple(f.shape))

DC_CASES = {
    "sphere14": (fields.eval_field(S.SphereSDF(0.5), (14, 14, 14)), 0.0),
    "cuboid16": (fields.eval_field(S.CuboidSDF([1, 1, 1]), (16, 16, 16)), 0.0),
    "csg18": (fields.eval_field(fields.csg_box_minus_sphere(), (18, 18, 18)), 0.0),
    "torus_12x16x20": (fields.eval_field(fields.torus(), (12, 16, 20)), 0.05),
}
for name, (vals, level) in DC_CASES.items():
    g = ref.UniformGrid(list(vals.shape))
    g.set_values(vals.cuda())
    its = ref.get_intersection(g, level, True)
    v, f = ref.dual_contouring(g, level, None, 1e-2, 1e-6)
    np.savez_compressed(out / f"dc_{name}.npz", values=vals.numpy(), level=np.float32(level),
                        its_points=its.get_points().cpu().numpy(), its_normals=its.get_normals().cpu().numpy(),
                        its_edges=its.get_edges().cpu().numpy(), its_is_out=its.get_is_out().cpu().numpy(),
                        its_cell_indices=its.get_cell_indices().cpu().numpy(),
                        its_cell_offsets=its.get_cell_offsets().cpu().numpy(),
                        v=v.cpu().numpy(), f=f.cpu().numpy())
    print("dc", name, tuple(v.shape), tuple(f.shape), its.num_points())
