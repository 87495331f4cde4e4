import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import isoext_b200 as iso
from isoext_b200 import sdf as S, _lib
lib = _lib.lib()
n = 1024
g = iso.SparseGrid([n] * 3)
g.populate_from_dense(iso.ImplicitGrid([n] * 3, S.SphereSDF(0.7)))
ref = None
for knob in (0, 100, 0, 100):
    lib.isoext_debug_set_tuning(3, knob)
    for _ in range(3): v, f = iso.marching_cubes(g)
    if ref is None: ref = (v.clone(), f.clone())
    same = torch.equal(v, ref[0]) and torch.equal(f, ref[1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): iso.marching_cubes(g)
    e1.record(); torch.cuda.synchronize()
    print(f"sparse mc knob3={knob}: {e0.elapsed_time(e1)/10*1e3:.1f} us same={same}", flush=True)
