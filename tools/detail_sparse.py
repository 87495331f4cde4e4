"""GPU: warm per-kernel timings of sparse marching cubes / dual contouring on the c5 band (1024^3-equivalent sphere)."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import sdf as S, _lib
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
prog = S.SphereSDF(0.7)
g = iso.SparseGrid([n] * 3)
g.populate_from_dense(iso.ImplicitGrid([n] * 3, prog))
print("cells", g.get_num_cells())
for op, run in (("mc", lambda: iso.marching_cubes(g)), ("dc", lambda: iso.dual_contouring(g))):
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    print(f"sparse {op} {n}^3 band: {e0.elapsed_time(e1)/10*1e3:.1f} us per call (no detail timing)")
    lib.isoext_debug_detail_enable(1)
    for _ in range(10): run()
    buf = C.create_string_buffer(1 << 16)
    lib.isoext_debug_detail_report(buf, len(buf))
    lib.isoext_debug_detail_enable(0)
    print(buf.value.decode())
