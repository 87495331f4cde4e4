"""GPU: host->device bandwidth from pinned memory: torch pin_memory vs cudaHostAlloc (default / write-combined), 1..4 streams."""
import ctypes as C, time, sys
import torch
rt = C.CDLL("libcudart.so.12")
n = 1 << 30   # floats (4 GiB)
dev = torch.device("cuda:0")
dst = torch.empty(n, dtype=torch.float32, device=dev)

def bw(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return 4 * n * reps / (time.perf_counter() - t0) / 1e9

host = torch.empty(n, dtype=torch.float32, pin_memory=True); host.fill_(1.0)
print(f"torch pin_memory, one copy_: {bw(lambda: dst.copy_(host, non_blocking=True)):.1f} GB/s", flush=True)
for ns in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    def multi():
        step = n // ns
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    print(f"torch pin_memory, {ns} streams: {bw(multi):.1f} GB/s", flush=True)
for chunk_mb in (64, 256):
    def chunked():
        step = chunk_mb << 18
        for a in range(0, n, step):
            dst[a:a + step].copy_(host[a:a + step], non_blocking=True)
    print(f"torch pin_memory, chunks of {chunk_mb} MiB on one stream: {bw(chunked):.1f} GB/s", flush=True)
del host
for flags, name in ((0, "cudaHostAllocDefault"), (4, "cudaHostAllocWriteCombined"), (1, "cudaHostAllocPortable")):
    p = C.c_void_p()
    rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(4 * n), C.c_uint(flags))
    if rc != 0:
        print(name, "failed", rc); continue
    C.memset(p, 0, 4 * n) if flags != 4 else rt.cudaMemset  # touch (skip slow WC memset from the CPU? do it anyway below)
    if flags == 4:
        C.memset(p, 0, 4 * n)
    def cp():
        rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(4 * n), C.c_int(1), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    print(f"{name}: {bw(cp):.1f} GB/s", flush=True)
    rt.cudaFreeHost(p)
# device -> host
hostb = torch.empty(n // 8, dtype=torch.float32, pin_memory=True)
t = bw(lambda: hostb.copy_(dst[: n // 8], non_blocking=True)) / 8
print(f"D2H pinned: {t:.1f} GB/s")
