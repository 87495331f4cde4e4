import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import fields, oracle
import isoext_b200 as iso
from isoext_b200 import dist as idist, sdf as S
from test_dist_cpu import simulate_rank
vals = fields.eval_field(S.CuboidSDF([1, 1, 1]), (65, 65, 65))
world = 8
for r in range(world):
    sg = idist.SlabGrid([65] * 3, rank=r, world=world)
    p = sg.plan
    sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1].cuda())
    v_own, f, n_lo, n_hi = idist.marching_cubes_local(sg)
    ov, of, on_lo, on_hi = simulate_rank(vals.numpy(), r, world, "nagae")
    print(r, p["c_lo"], p["c_hi"], "gpu own", len(v_own), "nlo/nhi", n_lo, n_hi, "T", len(f), "| sim own", len(ov), on_lo, on_hi, "T", len(of), "thr", sg.thresholds)
    if len(v_own) != len(ov) or (len(f) and not np.array_equal(f.cpu().numpy(), of)):
        # compare the extended lists
        v_ext, f_ext, _ = oracle.mc_dense(vals.numpy(), 0.0, "nagae", x_range=(p["ext_lo"], p["ext_hi"]))
        from isoext_b200.mc import mc_dense_raw
        gv, gf, a, b, _ = mc_dense_raw(sg._ext, (p["n_ext"], 65, 65), sg.aabb_min, sg.aabb_max, 0.0, 0, sg._ws, x_offset=p["ext_lo"], x_global=65, emit_range=(0, p["n_ext"] - 1), x_thresholds=sg.thresholds)
        print("   ext: gpu V", len(gv), "F", len(gf), "oracle V", len(v_ext), "F", len(f_ext), "equalV", gv.shape == v_ext.shape and np.array_equal(gv.cpu().numpy(), v_ext))
        xs = np.unique(v_ext[:, 0]); gx = np.unique(gv.cpu().numpy()[:, 0])
        print("   oracle distinct x:", len(xs), "gpu distinct x:", len(gx), "missing x:", [float(x) for x in xs if x not in gx][:5])
