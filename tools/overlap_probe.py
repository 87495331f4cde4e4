"""GPU experiment: do two surface-sized tails overlap when they run on separate streams?

Two threads, each with its own grid (512^3 torus: 0.09 ms volume stream + 0.25 ms surface tail) and its own CUDA
stream, extract concurrently (ctypes releases the GIL inside the C-ABI calls).  If the latency-bound tail kernels
overlap, the pair finishes in well under 2x the time of one.  Decides whether splitting the tail into concurrent
branches (candidate keys + sort || triangle analysis) can pay."""
import sys
import threading
import time

sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch

import fields
import isoext_b200 as iso


def build(n, fn):
    g = iso.UniformGrid([n] * 3)
    ax = fields.axis(n).cuda(); view = g.values_view()
    for a in range(0, n, 16):
        P = torch.stack(torch.meshgrid(ax[a:a + 16], ax, ax, indexing="ij"), dim=-1); view[a:a + 16] = fn(P); del P
    for _ in range(3):
        iso.marching_cubes(g)
    return g


def loop(g, reps, stream):
    with torch.cuda.stream(stream):
        for _ in range(reps):
            iso.marching_cubes(g)
        stream.synchronize()


for n in (256, 512):
    gs = [build(n, fields.torus()) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    reps = 200
    torch.cuda.synchronize()
    t0 = time.perf_counter(); loop(gs[0], reps, streams[0]); one = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=loop, args=(gs[i], reps, streams[i])) for i in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    both = (time.perf_counter() - t0) / reps
    print(f"{n}^3 torus: one stream {one * 1e3:.3f} ms/extraction; two concurrent streams {both * 1e3:.3f} ms per PAIR "
          f"({2 * one / both:.2f}x throughput)", flush=True)
