"""GPU: generic segmented sort on sparse MC / DC band and dense DC: per-kernel times."""
import ctypes as C, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import sdf as S, _lib
lib = _lib.lib()
n = 1024
g = iso.SparseGrid([n] * 3)
g.populate_from_dense(iso.ImplicitGrid([n] * 3, S.SphereSDF(0.7)))
d = iso.UniformGrid([512] * 3)
ax = fields.axis(512).cuda(); view = d.values_view(); fn = fields.csg_box_minus_sphere()
for x0 in range(0, 512, 32):
    P = torch.stack(torch.meshgrid(ax[x0:x0 + 32], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 32] = fn(P); del P
def detail(run):
    lib.isoext_debug_detail_enable(1)
    for _ in range(5): run()
    buf = C.create_string_buffer(1 << 16)
    lib.isoext_debug_detail_report(buf, len(buf))
    lib.isoext_debug_detail_enable(0)
    tot = 0.0; out = []
    for l in buf.value.decode().splitlines():
        w = l.split()
        if not w: continue
        tot += float(w[-3])
        out.append(f"{w[0][:24]}={w[-3]}")
    return tot, " ".join(out)
for grp in [int(a) for a in sys.argv[1:]] or [0]:
    lib.isoext_debug_set_tuning(3, grp)
    for name, run in (("sparse mc", lambda: iso.marching_cubes(g)), ("sparse dc", lambda: iso.dual_contouring(g)), ("dense dc 512 csg", lambda: iso.dual_contouring(d))):
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        tot, det = detail(run)
        print(f"g={grp} {name}: {e0.elapsed_time(e1)/10*1e3:.1f} us, kernel sum {tot:.0f} | {det}", flush=True)
