"""GPU: A/B of development knobs on dense marching cubes.  usage: ab_knobs.py n field key=value[,value...] ..."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import _lib
lib = _lib.lib()
n = int(sys.argv[1]); which = sys.argv[2]
fn = {"torus": fields.torus(), "csg": fields.csg_box_minus_sphere(), "sphere": fields.sphere()}[which]
g = iso.UniformGrid([n] * 3)
ax = fields.axis(n).cuda(); view = g.values_view()
for x0 in range(0, n, 16):
    P = torch.stack(torch.meshgrid(ax[x0:x0 + 16], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 16] = fn(P); del P
ref = None
for spec in sys.argv[3:]:
    key, vals = spec.split("=")
    for v in vals.split(","):
        lib.isoext_debug_set_tuning(int(key), int(v))
        for _ in range(3): V, F = iso.marching_cubes(g)
        if ref is None: ref = (V.clone(), F.clone())
        same = torch.equal(V.view(torch.int32), ref[0].view(torch.int32)) and torch.equal(F, ref[1])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if n <= 1024 else 5
        e0.record()
        for _ in range(reps): iso.marching_cubes(g)
        e1.record(); torch.cuda.synchronize()
        lib.isoext_debug_detail_enable(1)
        for _ in range(5): iso.marching_cubes(g)
        buf = C.create_string_buffer(1 << 16)
        lib.isoext_debug_detail_report(buf, len(buf))
        lib.isoext_debug_detail_enable(0)
        det = " ".join(f"{l.split()[0][:22]}={l.split()[-3]}" for l in buf.value.decode().splitlines() if l.strip() and "signbits" not in l)
        print(f"{which} {n}^3 knob{key}={v}: {e0.elapsed_time(e1)/reps*1e3:.1f} us same={same} | {det}", flush=True)
    lib.isoext_debug_set_tuning(int(key), 0)
