"""Development check (GPU): isoext_b200 dense MC vs CPU oracle vs the reference CUDA build."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import isoext_b200 as iso
import oracle, fields
from oracle import ref
from isoext_b200 import sdf as S

def ours(vals, level, method, aabb=((-1,-1,-1),(1,1,1))):
    g = iso.UniformGrid(list(vals.shape), aabb[0], aabb[1])
    g.set_values(vals.cuda())
    v, f = iso.marching_cubes(g, level, method)
    return g, v, f

def cmp(name, vals, level=0.0, method="nagae", aabb=((-1,-1,-1),(1,1,1)), use_ref=True):
    g, v, f = ours(vals, level, method, aabb)
    ov, of, na = oracle.mc_dense(vals.numpy(), level, method, aabb[0], aabb[1])
    if v is None:
        ok = len(of) == 0
        print(f"{name:34s} {method:9s} lvl={level:+.2f} EMPTY ok={ok}")
        return ok
    v_, f_ = v.cpu().numpy(), f.cpu().numpy()
    okv = v_.shape == ov.shape and np.array_equal(v_.view(np.uint32), ov.view(np.uint32))
    okf = f_.shape == of.shape and np.array_equal(f_, of)
    msg = f"{name:34s} {method:9s} lvl={level:+.2f} V={len(v_)}/{len(ov)} F={len(f_)}/{len(of)} V_bits={okv} F_eq={okf}"
    if use_ref and ref.available():
        rg = ref.UniformGrid(list(vals.shape), aabb[0], aabb[1]); rg.set_values(vals.cuda())
        rv, rf = ref.marching_cubes(rg, level, method)
        rokv = rv is not None and rv.shape == v.shape and bool((rv.view(torch.int32) == v.view(torch.int32)).all())
        rokf = rf is not None and rf.shape == f.shape and bool((rf == f).all())
        msg += f" | ref V_bits={rokv} F_eq={rokf}"
        okv &= rokv; okf &= rokf
    print(msg, flush=True)
    return okv and okf

allok = True
s5 = fields.eval_field(S.SphereSDF(0.5), (64, 64, 64))
allok &= cmp("64 sphere .5", s5)
allok &= cmp("64 sphere .5", s5, method="lorensen")
allok &= cmp("64 sphere .5", s5, level=0.1)
allok &= cmp("64 sphere .5", s5, level=-0.1)
allok &= cmp("8 empty", torch.ones(8, 8, 8))
allok &= cmp("aniso 16x32x48", fields.eval_field(S.SphereSDF(0.5), (16, 32, 48)))
allok &= cmp("aniso 8x64x16", fields.eval_field(S.SphereSDF(0.5), (8, 64, 16)))
allok &= cmp("odd 33x35x37 torus", fields.eval_field(fields.torus(), (33, 35, 37)))
allok &= cmp("65 cuboid (exact hits)", fields.eval_field(S.CuboidSDF([1, 1, 1]), (65, 65, 65)))
allok &= cmp("65 cuboid lorensen", fields.eval_field(S.CuboidSDF([1, 1, 1]), (65, 65, 65)), method="lorensen")
allok &= cmp("64 gyroid", fields.eval_field(fields.gyroid(6.0), (64, 64, 64)))
allok &= cmp("noise 40^3", fields.noise((40, 40, 40)))
allok &= cmp("noise 37x41x131 lorensen", fields.noise((37, 41, 131), 1), method="lorensen")
allok &= cmp("boundary-crossing sphere 1.2", fields.eval_field(S.SphereSDF(1.2), (48, 48, 48)))
allok &= cmp("aabb [0,3]x[-2,1]x[5,6]", fields.eval_field(S.SphereSDF(0.5), (40, 40, 40)), aabb=((0, -2, 5), (3, 1, 6)))
allok &= cmp("2x2x2", torch.tensor([[[-1., 1], [1, 1]], [[1, 1], [1, -1]]]))
allok &= cmp("3x2x200 noise", fields.noise((3, 2, 200), 2))
allok &= cmp("noise 20x24x128", fields.noise((20, 24, 128), 3))
allok &= cmp("noise 7x5x256 lorensen", fields.noise((7, 5, 256), 4), method="lorensen")
allok &= cmp("sphere 33x40x128", fields.eval_field(S.SphereSDF(0.8), (33, 40, 128)))
allok &= cmp("256 torus", fields.eval_field(fields.torus(), (256, 256, 256)))
allok &= cmp("256 csg", fields.eval_field(fields.csg_box_minus_sphere(), (256, 256, 256)))
allok &= cmp("256 quickstart", fields.eval_field(fields.quickstart(), (256, 256, 256)))
print("ALL OK" if allok else "SOME FAILED")

# get_points parity
g = iso.UniformGrid([31, 17, 53], (-1, 0, 2), (1, 5, 2.5))
p = g.get_points().cpu().numpy()
op = oracle.points_dense((31, 17, 53), (-1, 0, 2), (1, 5, 2.5))
print("points bits equal oracle:", np.array_equal(p.view(np.uint32), op.view(np.uint32)))
if ref.available():
    rg = ref.UniformGrid([31, 17, 53], (-1, 0, 2), (1, 5, 2.5))
    print("points bits equal ref:", bool((rg.get_points().view(torch.int32) == g.get_points().view(torch.int32)).all()))

# timing at 512^3
vals = fields.eval_field(fields.torus(), (512, 512, 512)).cuda()
g = iso.UniformGrid([512] * 3); g.set_values(vals)
for _ in range(3): v, f = iso.marching_cubes(g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): v, f = iso.marching_cubes(g)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"512^3 torus ours: {ms:.3f} ms  {512**3/ms/1e6:.1f} Gvox/s  V={len(v)} F={len(f)}")
if ref.available():
    rg = ref.UniformGrid([512] * 3); rg.set_values(vals)
    for _ in range(2): rv, rf = ref.marching_cubes(rg)
    torch.cuda.synchronize()
    t = time.time()
    for _ in range(3): rv, rf = ref.marching_cubes(rg)
    torch.cuda.synchronize()
    rms = (time.time() - t) / 3 * 1e3
    print(f"512^3 torus ref : {rms:.3f} ms  {512**3/rms/1e6:.1f} Gvox/s V={len(rv)} F={len(rf)}")
    print("512 parity vs ref: V", bool((rv.view(torch.int32) == v.view(torch.int32)).all()), "F", bool((rf == f).all()))
