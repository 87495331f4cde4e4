import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch, fields, oracle
import isoext_b200 as iso
from test_mc_gpu import CASES
for name in sys.argv[1:] or ["xface_6x200x200_second_level_by_y"]:
    vals, level, aabb = CASES[name]()
    ov, of, _ = oracle.mc_dense(vals.numpy(), level, "nagae", aabb[0], aabb[1])
    g = iso.UniformGrid(list(vals.shape), aabb[0], aabb[1]); g.set_values(vals.cuda())
    for it in range(4):
        v, f = iso.marching_cubes(g, level)
        vv, ff = v.cpu().numpy(), f.cpu().numpy()
        okv = vv.shape == ov.shape and np.array_equal(vv.view(np.uint32), ov.view(np.uint32))
        okf = ff.shape == of.shape and np.array_equal(ff, of)
        print(name, "call", it, "V", vv.shape, ov.shape, okv, "F", ff.shape, of.shape, okf, "hints", dict(g._hints))
        if not okv and vv.shape == ov.shape:
            bad = np.nonzero((vv.view(np.uint32) != ov.view(np.uint32)).any(axis=1))[0]
            print("  first bad vertex rows", bad[:5], len(bad), vv[bad[:3]], ov[bad[:3]])
            srt = np.lexsort((vv[:, 2], vv[:, 1], vv[:, 0]))
            print("  ours sorted already?", np.array_equal(srt, np.arange(len(vv))), " same set?", np.array_equal(np.sort(vv.view(np.uint32), axis=0), np.sort(ov.view(np.uint32), axis=0)))
    other = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(vals.shape)).astype(np.float32))
    oov, oof, _ = oracle.mc_dense(other.numpy(), 0.0, "nagae", aabb[0], aabb[1])
    for tag, field, lv, (ev, ef) in (("noise", other, 0.0, (oov, oof)), ("back", vals, level, (ov, of))):
        g.set_values(field.cuda())
        for it in range(3):
            v, f = iso.marching_cubes(g, lv)
            vv, ff = v.cpu().numpy(), f.cpu().numpy()
            okv = vv.shape == ev.shape and np.array_equal(vv.view(np.uint32), ev.view(np.uint32))
            okf = ff.shape == ef.shape and np.array_equal(ff, ef)
            print(tag, "call", it, "V", vv.shape, ev.shape, okv, "F", ff.shape, ef.shape, okf, "hints", dict(g._hints))
            if not okv and vv.shape == ev.shape:
                bad = np.nonzero((vv.view(np.uint32) != ev.view(np.uint32)).any(axis=1))[0]
                print("  bad vertex rows", bad[:6], len(bad)); print(vv[bad[:3]]); print(ev[bad[:3]])
                a, b = vv[:-1], vv[1:]
                lt = (a[:, 0] < b[:, 0]) | ((a[:, 0] == b[:, 0]) & ((a[:, 1] < b[:, 1]) | ((a[:, 1] == b[:, 1]) & (a[:, 2] < b[:, 2]))))
                print("  ours strictly sorted:", bool(lt.all()), "unsorted at", np.nonzero(~lt)[0][:6], " same set:", np.array_equal(np.sort(vv.view(np.uint32).view([('x','u4'),('y','u4'),('z','u4')]).ravel()), np.sort(ev.view(np.uint32).view([('x','u4'),('y','u4'),('z','u4')]).ravel())))
# fresh grid, noise only (first call = count + emit)
g2 = iso.UniformGrid(list(vals.shape), aabb[0], aabb[1]); g2.set_values(other.cuda())
v, f = iso.marching_cubes(g2, 0.0)
print("fresh noise first call:", np.array_equal(v.cpu().numpy().view(np.uint32), oov.view(np.uint32)) if v.shape == oov.shape else (v.shape, oov.shape), np.array_equal(f.cpu().numpy(), oof) if f.shape == oof.shape else (f.shape, oof.shape), dict(g2._hints))
