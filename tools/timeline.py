"""GPU: event timeline (start/end per kernel, across streams) of ONE marching_cubes call."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import _lib
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
which = sys.argv[2] if len(sys.argv) > 2 else "torus"
fn = {"torus": fields.torus(), "csg": fields.csg_box_minus_sphere()}[which]
g = iso.UniformGrid([n] * 3)
ax = fields.axis(n).cuda(); view = g.values_view()
for x0 in range(0, n, 32):
    P = torch.stack(torch.meshgrid(ax[x0:x0 + 32], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 32] = fn(P); del P
for _ in range(4): iso.marching_cubes(g)
torch.cuda.synchronize()
lib.isoext_debug_detail_enable(1)
iso.marching_cubes(g)
buf = C.create_string_buffer(1 << 18)
lib.isoext_debug_detail_timeline(buf, len(buf))
lib.isoext_debug_detail_enable(0)
print(buf.value.decode())
