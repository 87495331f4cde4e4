"""GPU: time k_signbits variants inside the real marching_cubes call (CUDA events via the profile hooks)."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import _lib
lib = _lib.lib()
for n in (512, 1024):
    vals = fields.eval_field(fields.torus(), (n, n, n)).cuda() if n <= 512 else None
    g = iso.UniformGrid([n] * 3)
    if vals is None:
        ax = fields.axis(n).cuda(); fn = fields.torus(); view = g.values_view()
        for x0 in range(0, n, 32):
            P = torch.stack(torch.meshgrid(ax[x0:x0 + 32], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 32] = fn(P); del P
    else:
        g.set_values(vals)
    ref_v = None
    for per_sm in (4, 8, 16, 32):
        for var in (0, 1, 4, 2, 3, 5):
            lib.isoext_debug_set_signbits_variant(var | (per_sm << 8))
            for _ in range(3): v, f = iso.marching_cubes(g)
            torch.cuda.synchronize()
            lib.isoext_profile_begin()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): v, f = iso.marching_cubes(g)
            e1.record(); torch.cuda.synchronize()
            ms, cnt, la = C.c_double(), C.c_int64(), C.c_int64()
            lib.isoext_profile_end(C.byref(ms), C.byref(cnt), C.byref(la))
            k = ms.value / cnt.value
            if ref_v is None: ref_v = (v.clone(), f.clone())
            ok = torch.equal(v, ref_v[0]) and torch.equal(f, ref_v[1])
            print(f"n={n} per_sm={per_sm:2d} var={var}: k_signbits {k*1e3:8.1f} us  {4*n**3/k/1e6:7.1f} GB/s  total {e0.elapsed_time(e1)/20*1e3:8.1f} us  same={ok}", flush=True)
    lib.isoext_debug_set_signbits_variant(0)
