"""Per-phase wall time of one distributed marching_cubes step (run under torchrun).
    torchrun --nproc-per-node N tools/dist_timing.py [n] [field] [iters]
Phases are bracketed with torch.cuda.synchronize() (so they do not overlap), per rank."""
import os, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch, torch.distributed as dist
import bench
import isoext_b200 as iso
from isoext_b200 import dist as idist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
field = sys.argv[2] if len(sys.argv) > 2 else "csg"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rank, lr, world = bench.dist_env()
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
sg = idist.SlabGrid([n, n, n], group=dist.group.WORLD)
x0, x1 = sg.owned_point_range()
sg.set_owned_values(bench.build_field_gpu(bench.field_fn(field), n, x0, x1, dev))
for _ in range(3):
    idist.marching_cubes(sg)
torch.cuda.synchronize(); dist.barrier()

def tick():
    torch.cuda.synchronize(); return time.perf_counter()

acc = {}
for it in range(iters):
    dist.barrier(); t0 = tick()
    sg.exchange_halos(); t1 = tick()
    v_own, f, n_lo, n_hi = idist.marching_cubes_local(sg, 0.0, "nagae"); t2 = tick()
    t3 = t2
    idist.globalize_faces_(sg, f, n_lo, n_hi); t4 = tick()
    for k, d in (("halo", t1 - t0), ("local", t2 - t1), ("bases", t3 - t2), ("relabel", t4 - t3), ("total", t4 - t0)):
        acc.setdefault(k, []).append(d * 1e3)
row = [min(acc[k]) for k in ("halo", "local", "bases", "relabel", "total")]
out = torch.tensor(row + [float(v_own.shape[0]), float(f.shape[0])], device=dev, dtype=torch.float64)
allr = [torch.empty_like(out) for _ in range(world)]
dist.all_gather(allr, out)
if rank == 0:
    print(f"# {n}^3 {field}, world {world}, transport {'peer' if sg._peer is not None else 'nccl'}; best-of-{iters} ms per phase "
          f"(phases serialised by syncs; relabel includes the count exchange)")
    print("rank   halo  local  bases relabel  total      V_own       T")
    for r, t in enumerate(allr):
        t = t.tolist()
        print(f"{r:4d} {t[0]:6.3f} {t[1]:6.3f} {t[2]:6.3f} {t[3]:7.3f} {t[4]:6.3f} {int(t[5]):10d} {int(t[6]):10d}")
# detail of the local stage on every rank: per-kernel times
from isoext_b200 import _lib
import ctypes as C
lib = _lib.lib()
lib.isoext_debug_detail_enable(1)
idist.marching_cubes_local(sg, 0.0, "nagae"); torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 14)
lib.isoext_debug_detail_report(buf, len(buf))
lib.isoext_debug_detail_enable(0)
for r in range(world):
    dist.barrier()
    if r == rank and (world <= 2 or r in (0, 1, world // 2)):
        print(f"--- rank {r} kernel detail\n{buf.value.decode()}", flush=True)
# whole step, no intermediate syncs
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    idist.marching_cubes(sg)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"whole step (max over ranks): {float(ms):.3f} ms = {n ** 3 / float(ms) / 1e6:.0f} Gvox/s", flush=True)
sg.close()
dist.destroy_process_group()
