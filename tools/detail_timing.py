"""GPU: warm per-kernel timings (CUDA events around every launch) of one workload."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import _lib
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
which = sys.argv[2] if len(sys.argv) > 2 else "torus"
op = sys.argv[3] if len(sys.argv) > 3 else "mc"
fn = {"torus": fields.torus(), "csg": fields.csg_box_minus_sphere(), "sphere": fields.sphere()}[which]
g = iso.UniformGrid([n] * 3)
ax = fields.axis(n).cuda(); view = g.values_view()
for x0 in range(0, n, 32):
    P = torch.stack(torch.meshgrid(ax[x0:x0 + 32], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 32] = fn(P); del P
run = (lambda: iso.marching_cubes(g)) if op == "mc" else (lambda: iso.dual_contouring(g))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"{op} {which} {n}^3: {e0.elapsed_time(e1)/10*1e3:.1f} us per call (no detail timing)")
lib.isoext_debug_detail_enable(1)
for _ in range(10): run()
buf = C.create_string_buffer(1 << 16)
lib.isoext_debug_detail_report(buf, len(buf))
lib.isoext_debug_detail_enable(0)
print(buf.value.decode())
