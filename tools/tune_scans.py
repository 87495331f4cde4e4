"""GPU: per-kernel timings of the dense MC pipeline for several grid sizes of the chained-scan kernels."""
import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, fields
import isoext_b200 as iso
from isoext_b200 import _lib
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
g = iso.UniformGrid([n] * 3)
ax = fields.axis(n).cuda(); view = g.values_view(); fn = fields.torus()
for x0 in range(0, n, 32):
    P = torch.stack(torch.meshgrid(ax[x0:x0 + 32], ax, ax, indexing="ij"), dim=-1); view[x0:x0 + 32] = fn(P); del P
for bpsm in (4, 1, 2, 3, 6, 8, 16):
    lib.isoext_debug_set_tuning(0, bpsm)
    for _ in range(3): iso.marching_cubes(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): iso.marching_cubes(g)
    e1.record(); torch.cuda.synchronize()
    lib.isoext_debug_detail_enable(1)
    for _ in range(5): iso.marching_cubes(g)
    buf = C.create_string_buffer(1 << 16)
    lib.isoext_debug_detail_report(buf, len(buf))
    lib.isoext_debug_detail_enable(0)
    rows = [l.split() for l in buf.value.decode().splitlines() if l.strip()]
    pick = {r[0]: r for r in rows if r and r[0] in ("k_scan_rows", "k_scan_entries", "k_unique")}
    print(f"scan blocks/SM {bpsm:2d}: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us per extraction | " +
          " | ".join(f"{k} {pick[k][-3] if k in pick else '?'} us" for k in ("k_scan_rows", "k_scan_entries", "k_unique")), flush=True)
lib.isoext_debug_set_tuning(0, 0)
