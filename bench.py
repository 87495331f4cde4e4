#!/usr/bin/env python3
"""bench.py -- headline benchmark of the iso-surface extraction hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path (marching_cubes, default LUT method "nagae", level 0) over one
synthetic analytic field:
  N = 1  : BASELINE.json configs[1], 512^3 dense torus SDF (R=.5, r=.2) on [-1,1]^3      ("c2")
  N > 1  : BASELINE.json configs[2], 2048^3 CSG box-minus-sphere, slab-sharded on dim 0      ("c3")
           one process per GPU (torchrun), halo pull + vertex-id bases as kernels over NVLink peer memory, slab cuts
           balanced by measured load; strong scaling.  The N = 1 line also carries "scaling_base": the same c3
           workload timed on the single GPU, so that 1 -> N comparisons have a same-workload base.
Metric: Gvoxels/s = X*Y*Z / t / 1e9 (whole job).  `value` is timed with the field resident in HBM
(CUDA events, max over ranks); `e2e` goes through the public API from pinned HOST memory, H2D copy of
the field and D2H read of the mesh inside the timed region.  Inputs (537 MB / 34 GB) are larger than
the 126 MB L2, so no explicit L2 flush is needed between iterations.
`--impl reference` times the reference's own implementation of the path: GuangyanCai/isoext has NO CPU
extraction path, so that arm runs the UNMODIFIED reference CUDA sources (oracle/_ref, built from
/root/reference by oracle/Makefile) on the same GPU and config -- the baseline BASELINE.json names.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "Gvoxels/s and % HBM roofline for MC/DC at 1/2/4/8 B200 vs reference CUDA"


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_ burst)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a
    thread every few ms (the timed region of the headline workload is only tens of ms long, far below what an
    `nvidia-smi -lms` loop resolves); nvidia-smi is the fallback when the NVML binding is missing."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def _poll_nvml(self, nv, h):
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
            except Exception:
                pass
            time.sleep(0.001)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "nvml"
            self._thread = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self._thread.start()
        except Exception:
            self.source = "nvidia-smi"

    def stop(self):
        if self.source == "nvml":
            self._stop.set()
            self._thread.join(timeout=1.0)
        if not self.sm:   # fallback: one nvidia-smi query right after the timed region
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=20).stdout.strip().split(",")
                self.sm, self.max_mhz = [float(out[0])], float(out[1])
                for (bit, _), val in zip(self.REASONS, out[2:6]):
                    if val.strip().lower().startswith("active"):
                        self.mask |= bit
                self.source = "nvidia-smi (one query after the timed region)"
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for bit, name in self.REASONS if self.mask & bit), "samples": len(self.sm),
                "source": self.source}


def field_fn(name):
    import fields
    return {"torus": fields.torus(), "csg": fields.csg_box_minus_sphere(), "sphere": fields.sphere(0.5)}[name]


WORKLOADS = {
    "c1": dict(n=64, field="sphere", desc="64^3 sphere SDF UniformGrid marching_cubes (nagae)"),
    "c2": dict(n=512, field="torus", desc="512^3 dense torus SDF marching_cubes (nagae), 1 GPU"),
    "t1024": dict(n=1024, field="torus", desc="1024^3 dense torus SDF marching_cubes (nagae), 1 GPU"),
    "c3": dict(n=2048, field="csg", desc="2048^3 dense CSG box-minus-sphere marching_cubes, dim-0 slabs"),
}


def build_field_gpu(fn, n, x0, x1, device, slab=16):
    """(x1-x0, n, n) f32 field on the GPU from global indices (same formula as tests/fields.py)."""
    import torch
    import fields
    ax = fields.axis(n).to(device)
    out = torch.empty((x1 - x0, n, n), dtype=torch.float32, device=device)
    for a in range(x0, x1, slab):
        b = min(x1, a + slab)
        P = torch.stack(torch.meshgrid(ax[a:b], ax, ax, indexing="ij"), dim=-1)
        out[a - x0:b - x0] = fn(P)
        del P
    return out


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
def cpu_baseline_port(n_sample=256):
    """The oracle C port (single thread) on a bounded sample of the same field: the n_sample^3 grid of
    the torus SDF.  A reported baseline, not the target."""
    import fields
    import oracle
    vals = fields.eval_field(fields.torus(), (n_sample,) * 3).numpy()
    t0 = time.perf_counter()
    reps = 0
    while True:
        oracle.mc_dense(vals)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 64:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": n_sample ** 3 / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
            "sample": f"{n_sample}^3 torus SDF (same field family as the workload), oracle/oracle_c.c, "
                      f"{reps} repetitions in {dt * reps:.1f} s"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import isoext_b200 as iso
    from isoext_b200 import _lib

    wl_name = args.workload or ("c2" if world == 1 else "c3")
    wl = WORKLOADS[wl_name]
    n = wl["n"]
    fn = field_fn(wl["field"])
    lib = _lib.lib()
    sampler = ClockSampler(local_rank)
    peak, peak_src = load_peaks()

    cuts = None
    if world == 1:
        vals = build_field_gpu(fn, n, 0, n, dev)
        grid = iso.UniformGrid([n, n, n])
        grid.set_values(vals)

        def step():
            return iso.marching_cubes(grid)
    else:
        from isoext_b200 import dist as idist
        sg = idist.SlabGrid([n, n, n], group=dist.group.WORLD)
        x0, x1 = sg.owned_point_range()
        sg.set_owned_values(build_field_gpu(fn, n, x0, x1, dev))
        cuts = None
        if not args.even_slabs:
            # cut the slabs by measured load (setup, untimed): one extraction on even slabs gives the vertices per
            # cell layer; cost(layer) = time to stream one plane + surface-stage time per vertex
            v0, _ = idist.marching_cubes(sg)
            hist = idist.vertex_layer_histogram(v0, n, -1.0, 1.0, dist.group.WORLD)
            plane_s, vertex_s = 4.0 * n * n / (peak * 1e9), 0.43e-9
            cuts = idist.balanced_cuts((plane_s + vertex_s * hist).tolist(), world, ghost_cost=(vertex_s * hist).tolist())
            del v0
            sg.close()
            sg = idist.SlabGrid([n, n, n], group=dist.group.WORLD, cuts=cuts)
            x0, x1 = sg.owned_point_range()
            sg.set_owned_values(build_field_gpu(fn, n, x0, x1, dev))

        def step():
            return idist.marching_cubes(sg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        v, f = step()
    barrier()
    sampler.start()
    lib.isoext_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        v, f = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    sm_ms, sm_n, launches = C.c_double(), C.c_int64(), C.c_int64()
    lib.isoext_profile_end(C.byref(sm_ms), C.byref(sm_n), C.byref(launches))
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    voxels = float(n) ** 3
    value = voxels / (ms_step * 1e-3) / 1e9
    nV = 0 if v is None else int(v.shape[0])
    nT = 0 if f is None else int(f.shape[0])

    # ---- end to end through the public API with HOST buffers (N = 1: full field from pinned memory)
    e2e = None
    if world == 1:
        host = vals.cpu().pin_memory()
        grid2 = iso.UniformGrid([n, n, n])
        dbuf = torch.empty_like(vals)

        def e2e_step():
            dbuf.copy_(host, non_blocking=True)           # H2D of this step's input
            grid2.set_values(dbuf)
            vv, ff = iso.marching_cubes(grid2)
            return vv.cpu(), ff.cpu()                     # D2H of the step's result

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            hv, hf = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        e2e = {"value": voxels / dt / 1e9, "unit": "Gvoxels/s", "ms_per_step": dt * 1e3,
               "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": int(hv.numel() * 4 + hf.numel() * 4)}
    else:
        # multi-GPU: each rank uploads its own slab from pinned host memory and reads its mesh part back
        host = sg.owned_values().cpu().pin_memory()
        dbuf = torch.empty_like(sg.owned_values())

        def e2e_step():
            dbuf.copy_(host, non_blocking=True)
            sg.set_owned_values(dbuf)
            vv, ff = idist.marching_cubes(sg)
            return (vv.cpu() if vv is not None else None), (ff.cpu() if ff is not None else None)

        e2e_step()
        barrier()
        k = 3
        t0 = time.perf_counter()
        for _ in range(k):
            hv, hf = e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / k], device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        nb = torch.tensor([host.numel() * 4, (0 if hv is None else hv.numel() * 4) + (0 if hf is None else hf.numel() * 4)],
                          device=dev, dtype=torch.int64)
        dist.all_reduce(nb)
        e2e = {"value": voxels / float(dt.item()) / 1e9, "unit": "Gvoxels/s", "ms_per_step": float(dt.item()) * 1e3,
               "h2d_bytes_per_step": int(nb[0].item()), "d2h_bytes_per_step": int(nb[1].item())}

    if rank == 0:
        # roofline of the dominant kernel (k_signbits: the one volume-sized HBM stream); 4 B/voxel algorithmic
        # large grids are streamed as several x-chunks per step: bytes per launch = 4 B x voxels of that launch
        local_vox = voxels if world == 1 else float(sg.local_points())
        launches_per_step = max(1.0, sm_n.value / float(args.steps))
        local_vox = local_vox / launches_per_step
        k_ms = sm_ms.value / max(1, sm_n.value)
        achieved = 4.0 * local_vox / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        prof = ROOT / "profiles" / "k_signbits_traffic.json"
        if prof.exists():
            try:
                traffic = json.loads(prof.read_text()).get(wl_name)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 or wl_name == "c3" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "ours",
            "config": {"workload": f"{wl_name}: {wl['desc']}", "shape": [n, n, n], "level": 0.0, "method": "nagae",
                       "vertices": nV, "triangles": nT, "l2_policy": "inputs larger than L2 (no flush needed)",
                       "parallelism": "single GPU" if world == 1 else
                       f"dim-0 slabs x{world}; halo pull + vertex-id bases as kernels over NVLink peer memory"
                       + (f"; slab cuts balanced by measured load {cuts}" if cuts else "; even slabs")},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches.value),
            "roofline": {"bound": "hbm", "kernel": "k_signbits", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms, "kernel_launches_timed": int(sm_n.value), "kernel_launches_per_step": launches_per_step,
                         "algorithmic_bytes_per_launch": 4.0 * local_vox,
                         "whole_path_frac": (4.0 * voxels + 12.0 * nV + 12.0 * nT) / (ms_step * 1e-3) / 1e9 / peak / world},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port()
        if world == 1 and wl_name == "c2" and not args.no_scaling_base:
            # The N > 1 lines run c3 (2048^3 CSG, BASELINE.json configs[2], "sharded at 1/2/4/8 B200"): time the same
            # workload on this single GPU as well, so that a 1 -> N comparison has a same-workload base.
            try:
                del vals, grid, grid2, dbuf, host
                torch.cuda.empty_cache()
                n3 = WORKLOADS["c3"]["n"]
                g3 = iso.UniformGrid([n3] * 3)
                view = g3.values_view()
                fn3 = field_fn(WORKLOADS["c3"]["field"])
                for a in range(0, n3, 64):
                    view[a:a + 64] = build_field_gpu(fn3, n3, a, min(n3, a + 64), dev)
                for _ in range(3):
                    iso.marching_cubes(g3)
                torch.cuda.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record()
                for _ in range(5):
                    iso.marching_cubes(g3)
                s1.record()
                torch.cuda.synchronize()
                ms3 = s0.elapsed_time(s1) / 5
                line["scaling_base"] = {"workload": "c3: " + WORKLOADS["c3"]["desc"] + " on this single GPU (the N > 1 lines run c3)",
                                        "value": float(n3) ** 3 / (ms3 * 1e-3) / 1e9, "unit": "Gvoxels/s", "ms_per_step": ms3, "steps": 5}
            except Exception as exc:   # e.g. a GPU with less memory: the headline line must still be printed
                line["scaling_base"] = {"unavailable": repr(exc)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        sg.close()
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: UNMODIFIED reference CUDA sources (oracle/_ref) through their own entry points
    (UniformGrid ctor, set_values, mc::marching_cubes).  The reference has no CPU path and is single-GPU:
    under torchrun only rank 0 runs, on a bounded sample when the workload does not fit the reference
    (it is only correct below 2^29 cells and refuses more than INT_MAX points)."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    import torch
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libisoext_ref.so not built (needs /root/reference)"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    wl_name = args.workload or ("c2" if world == 1 else "c3")
    wl = WORKLOADS[wl_name]
    n_full = wl["n"]
    n = min(n_full, 512)     # bounded sample: the reference is invalid above 813 points/axis (uint overflow)
    fn = field_fn(wl["field"])
    if n == n_full:
        vals = build_field_gpu(fn, n, 0, n, dev)
        sample = f"full workload ({n}^3)"
    else:
        # central n^3 block of the n_full^3 field is not expressible in the reference (positions are tied to
        # the grid shape), so sample the same analytic field at n^3 over the same AABB.
        vals = build_field_gpu(fn, n, 0, n, dev)
        sample = f"{n}^3 sampling of the same field over the same AABB (reference cannot represent {n_full}^3)"
    grid = ref.UniformGrid([n, n, n])
    grid.set_values(vals)
    sampler = ClockSampler(local_rank)

    def step():
        v, f, nv, nf = ref.marching_cubes_timed_raw(grid, 0.0, "nagae")
        ref.free(v); ref.free(f)
        return nv, nf

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        nv, nf = step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    value = float(n) ** 3 / (ms_step * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{wl_name}: {wl['desc']}", "shape": [n, n, n], "level": 0.0, "method": "nagae",
                       "vertices": nv, "triangles": nf, "sample": sample,
                       "note": "reference = its own CUDA extension (thrust pipeline) on 1 GPU; it has no CPU path and no multi-GPU path"},
            "clocks": clocks,
            "cpu_baseline": {"value": value, "unit": "Gvoxels/s", "cores": 0, "kind": "reference",
                             "sample": sample + "; runs on the GPU: the reference has no CPU implementation of this path"},
            "e2e": {"value": value, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--even-slabs", action="store_true", help="N > 1: keep the even dim-0 split (default: cuts balanced by measured load)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true", help="N = 1: skip the extra c3 timing on one GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
