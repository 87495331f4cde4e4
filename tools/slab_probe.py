"""GPU (one device): per-kernel timing of ONE simulated rank's slab of the 2048^3 CSG workload at a given world size.
    python tools/slab_probe.py [world=8] [rank=1] [n=2048]
Builds only that rank's extended slab, runs the rank-local extraction (warm), prints the hints the single-call
path carries (n_big = candidates in oversized sort buckets, radix = last resort enqueued) and the per-kernel times."""
import ctypes as C
import sys

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch

import bench
from isoext_b200 import _lib
from isoext_b200 import dist as idist

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
lib = _lib.lib()
dev = torch.device("cuda", 0)
sg = idist.SlabGrid([n, n, n], rank=rank, world=world)
p = sg.plan
bench.build_field_gpu(bench.field_fn("csg"), n, p["ext_lo"], p["ext_hi"] + 1, dev, out=sg._ext)
for _ in range(4):
    v, f, n_lo, n_hi = idist.marching_cubes_local(sg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    idist.marching_cubes_local(sg)
e1.record(); torch.cuda.synchronize()
print(f"# world {world} rank {rank}: planes [{p['ext_lo']}, {p['ext_hi']}], {len(v)} owned vertices, {len(f)} triangles, "
      f"local step {e0.elapsed_time(e1) / 10 * 1e3:.1f} us; hints {sg._hints}")
lib.isoext_debug_detail_enable(1)
for _ in range(5):
    idist.marching_cubes_local(sg)
buf = C.create_string_buffer(1 << 16)
lib.isoext_debug_detail_report(buf, len(buf))
lib.isoext_debug_detail_enable(0)
print(buf.value.decode())
