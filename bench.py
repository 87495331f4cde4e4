#!/usr/bin/env python3
"""bench.py -- headline benchmark of the iso-surface extraction hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path (marching_cubes, default LUT method "nagae", level 0) over one
synthetic analytic field resident in HBM.  The SAME workload at every N, so the driver's 1 -> N arithmetic
is meaningful:

  headline (every N): BASELINE.json configs[2], 2048^3 dense CSG box-minus-sphere ("c3", 34.4 GB: the largest
           configuration that fits one B200).  N = 1: one UniformGrid.  N > 1: dim-0 slabs, one process per
           GPU (torchrun), halo pull + vertex-id bases as kernels over NVLink peer memory; strong scaling.
           The headline value at N > 1 uses EVEN slabs (no foreknowledge of the surface); the same loop with
           slab cuts balanced by the measured load of a previous extraction is reported beside it
           ("balanced_cuts"), as is the first (cold: count + emit, capacity discovery) call ("first_call_ms").
  sub_records (N = 1): configs[1] 512^3 torus ("c2", the reference-representable single-GPU case) and the
           north-star target 1024^3 torus ("t1024"), each timed with the same protocol and with its own
           whole_path_frac; configs[3] 512^3 CSG dual_contouring ("c4") and configs[4] SparseGrid narrow band
           ("c5", 1024^3-equivalent sphere, marching_cubes and dual_contouring), counts verified.
Every line carries the global vertex / triangle totals and a 64-bit order-sensitive checksum of the mesh
(V bits and F ids weighted by their global index); they are checked against the values the single-GPU
path produced (EXPECTED below), so a wrong mesh at any N fails loudly instead of printing a number.

Metric: Gvoxels/s = X*Y*Z / t / 1e9 (whole job).  `value` is timed with the field resident in HBM (CUDA
events, max over ranks); `e2e` goes through the public API from pinned HOST memory, H2D copy of the field
and D2H read of the mesh inside the timed region.  Inputs (>= 537 MB) are larger than the 126 MB L2, so no
explicit L2 flush is needed between iterations.
`--impl reference` times the reference's own implementation of the path: GuangyanCai/isoext has NO CPU
extraction path, so that arm runs the UNMODIFIED reference CUDA sources (oracle/_ref, built from
/root/reference by oracle/Makefile) on one GPU.  The reference cannot represent 2048^3 (it refuses more than
INT_MAX points and is wrong above 2^29 cells), so each step is a bounded sample of the workload -- the same
field sampled at 512^3 over the same AABB -- and the line says so (`same_workload`); it also carries a
full-size, identical-workload record for c2 (512^3 torus).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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

WORKLOADS = {
    "c1": dict(n=64, field="sphere", desc="64^3 sphere SDF UniformGrid marching_cubes (nagae)"),
    "c2": dict(n=512, field="torus", desc="512^3 dense torus SDF marching_cubes (nagae), 1 GPU"),
    "t1024": dict(n=1024, field="torus", desc="1024^3 dense torus SDF marching_cubes (nagae), 1 GPU (north-star target size)"),
    "c3": dict(n=2048, field="csg", desc="2048^3 dense CSG box-minus-sphere marching_cubes (nagae)"),
}
# Mesh of each workload as produced by the single-GPU path (bit-identical to the reference build where the
# reference can represent the grid: c2; closed-manifold / slab-equality tested for the others).  Every run,
# at every N, must reproduce counts and checksum.
EXPECTED = {
    "c2": {"vertices": 372576, "triangles": 745152, "checksum": 0x1264ae17dfaf9555},
    "t1024": {"vertices": 1492552, "triangles": 2985104, "checksum": 0x305beda00021f936},
    "c3": {"vertices": 9047904, "triangles": 18095804, "checksum": 0xd3e1e9356034ecb5},
}


def workload_config(wl_name):
    """The workload-defining part of the JSON line: identical in the `ours` and `reference` arms."""
    wl = WORKLOADS[wl_name]
    n = wl["n"]
    return {"workload": f"{wl_name}: {wl['desc']}", "shape": [n, n, n], "field": wl["field"], "aabb": [-1.0, 1.0], "level": 0.0,
            "method": "nagae", "l2_policy": "inputs larger than L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_ burst)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a
    thread every ms (the timed regions here are tens to hundreds of ms, below what an `nvidia-smi -lms` loop
    resolves); nvidia-smi is the fallback when the NVML binding is missing."""
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


def build_field_gpu(fn, n, x0, x1, device, slab=16, out=None):
    """(x1-x0, n, n) f32 field on the GPU from global indices (same formula as tests/fields.py)."""
    import torch
    import fields
    ax = fields.axis(n).to(device)
    if out is None:
        out = torch.empty((x1 - x0, n, n), dtype=torch.float32, device=device)
    for a in range(x0, x1, slab):
        b = min(x1, a + slab)
        P = torch.stack(torch.meshgrid(ax[a:b], ax, ax, indexing="ij"), dim=-1)
        out[a - x0:b - x0] = fn(P)
        del P
    return out


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


_GOLD = -7046029254386353131     # 0x9E3779B97F4A7C15 as a signed 64-bit integer


def mesh_checksum(v, f, v_base=0, t_base=0):
    """Order-sensitive 64-bit checksum (wrap-around int64 arithmetic on the device): sum over the float BITS of V and
    the ids of F, each multiplied by an odd weight derived from its GLOBAL flat index.  Per-rank parts add up to the
    checksum of the concatenated mesh, so it is comparable across N."""
    import torch
    tot = 0
    for t, base, salt in ((v, 3 * v_base, 1), (f, 3 * t_base, 2)):
        if t is None or t.numel() == 0:
            continue
        bits = (t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t.contiguous()).reshape(-1).to(torch.int64)
        idx = torch.arange(bits.numel(), dtype=torch.int64, device=t.device) + int(base)
        w = ((idx + salt) * _GOLD) | 1
        tot = (tot + int((bits * w).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return tot


def verify_mesh(wl_name, nV, nT, checksum):
    exp = EXPECTED.get(wl_name)
    if not exp:
        return None
    ok = nV == exp["vertices"] and nT == exp["triangles"] and ("checksum" not in exp or checksum == exp["checksum"])
    if not ok:
        raise SystemExit(f"bench.py: WRONG MESH for {wl_name}: got {nV} V / {nT} T / checksum {checksum:016x}, expected {exp}")
    return True


# ------------------------------------------------------------------------------------------------
def cpu_baseline_port(field="csg", n_sample=256):
    """The oracle C port (single thread) on a bounded sample of the headline field: the n_sample^3 grid of
    the same SDF over the same AABB.  A reported baseline, not the target."""
    import fields
    import oracle
    vals = fields.eval_field(field_fn(field), (n_sample,) * 3).numpy()
    t0 = time.perf_counter()
    reps = 0
    while True:
        oracle.mc_dense(vals)
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 64:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": n_sample ** 3 / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
            "sample": f"{n_sample}^3 sampling of the workload's field ({field}) over the same AABB, oracle/oracle_c.c, "
                      f"{reps} repetitions in {dt * reps:.1f} s"}


def timed_loop(step, steps, warmup, barrier, sampler=None, lib=None):
    """W untimed steps, then exactly K steps between barriers + synchronize, CUDA events on the current stream."""
    import torch
    for _ in range(warmup):
        out = step()
    barrier()
    if sampler:
        sampler.start()
    if lib:
        lib.isoext_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        out = step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = None
    if lib:
        sm_ms, sm_n, launches = C.c_double(), C.c_int64(), C.c_int64()
        lib.isoext_profile_end(C.byref(sm_ms), C.byref(sm_n), C.byref(launches))
        prof = (sm_ms.value, sm_n.value, launches.value)
    clocks = sampler.stop() if sampler else None
    return ms_total, out, prof, clocks


def single_gpu_record(iso, lib, wl_name, dev, steps, warmup, peak, keep=False):
    """Build the workload on this GPU, time the first (cold) call and the steady-state loop."""
    import torch
    wl = WORKLOADS[wl_name]
    n = wl["n"]
    grid = iso.UniformGrid([n, n, n])
    view = grid.values_view()
    fn = field_fn(wl["field"])
    for a in range(0, n, 64):
        build_field_gpu(fn, n, a, min(n, a + 64), dev, out=view[a:a + 64])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    iso.marching_cubes(grid)          # first call: count + emit (two syncs), workspace allocation, capacity discovery
    torch.cuda.synchronize()
    first_ms = (time.perf_counter() - t0) * 1e3
    sampler = ClockSampler(dev.index or 0)
    ms_total, (v, f), prof, clocks = timed_loop(lambda: iso.marching_cubes(grid), steps, warmup, torch.cuda.synchronize, sampler, lib)
    ms_step = ms_total / steps
    nV, nT = int(v.shape[0]), int(f.shape[0])
    csum = mesh_checksum(v, f)
    voxels = float(n) ** 3
    rec = {"workload": f"{wl_name}: {wl['desc']}", "shape": [n, n, n], "value": voxels / (ms_step * 1e-3) / 1e9, "unit": "Gvoxels/s",
           "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "first_call_ms": first_ms,
           "vertices": nV, "triangles": nT, "mesh_checksum": f"{csum:016x}", "mesh_verified": verify_mesh(wl_name, nV, nT, csum),
           "whole_path_frac": (4.0 * voxels + 12.0 * nV + 12.0 * nT) / (ms_step * 1e-3) / 1e9 / peak,
           "whole_path_algorithmic_bytes": 4.0 * voxels + 12.0 * nV + 12.0 * nT, "clocks": clocks}
    if prof and prof[1]:
        k_ms = prof[0] / prof[1]
        rec["k_signbits_ms"] = k_ms
        rec["k_signbits_frac_of_peak"] = 4.0 * voxels / (prof[1] / steps) / (k_ms * 1e-3) / 1e9 / peak
    if keep:
        return rec, grid, prof
    del grid, view, v, f
    torch.cuda.empty_cache()
    return rec, None, prof


def other_config_records(iso, lib, dev, steps, warmup, peak):
    """BASELINE.json configs[3] and configs[4] as timed sub-records of the N = 1 line (same protocol: W warm-ups, K
    steps between synchronisations, CUDA events; field / band resident): c4 = 512^3 box-minus-sphere dual_contouring,
    c5 = SparseGrid narrow band of the 1024^3-equivalent sphere r = 0.7 (the reference's documented example size; the
    4096^3-equivalent band is in tools/bench_extra.py c5big), marching_cubes and dual_contouring."""
    import torch
    from isoext_b200 import sdf as S
    recs = []
    # ---- c4
    n = 512
    grid = iso.UniformGrid([n, n, n])
    view = grid.values_view()
    fn = field_fn("csg")
    for a in range(0, n, 64):
        build_field_gpu(fn, n, a, min(n, a + 64), dev, out=view[a:a + 64])
    ms_total, (v, f), _, clocks = timed_loop(lambda: iso.dual_contouring(grid), steps, warmup, torch.cuda.synchronize, ClockSampler(dev.index or 0))
    ms = ms_total / steps
    nV, nT = int(v.shape[0]), int(f.shape[0])
    recs.append({"workload": "c4: 512^3 dense CSG box-minus-sphere dual_contouring (reg=1e-2, svd_tol=1e-6, normals from the field), 1 GPU",
                 "shape": [n] * 3, "value": float(n) ** 3 / (ms * 1e-3) / 1e9, "unit": "Gvoxels/s", "ms_per_step": ms, "steps": steps,
                 "warmup": warmup, "vertices": nV, "triangles": nT, "mesh_verified": (nV, nT) == (561818, 1123632),
                 "whole_path_frac": (4.0 * n ** 3 + 12.0 * nV + 12.0 * nT) / (ms * 1e-3) / 1e9 / peak, "clocks": clocks})
    del grid, view, v, f
    # ---- c5 (1024^3-equivalent band, populated on the GPU from the analytic program: no 1024^3 field is allocated)
    n = 1024
    band = iso.SparseGrid([n, n, n])
    band.populate_from_dense(iso.ImplicitGrid([n, n, n], S.SphereSDF(0.7)))
    cells = band.get_num_cells()
    for op, run in (("marching_cubes", lambda: iso.marching_cubes(band)), ("dual_contouring", lambda: iso.dual_contouring(band))):
        ms_total, (v, f), _, clocks = timed_loop(run, steps, warmup, torch.cuda.synchronize, ClockSampler(dev.index or 0))
        ms = ms_total / steps
        nV, nT = int(v.shape[0]), int(f.shape[0])
        ok = (cells, nV) == (2416778, 2416776) if op == "marching_cubes" else (cells, nV) == (2416778, 2416778)
        recs.append({"workload": f"c5: SparseGrid narrow band, 1024^3-equivalent sphere r=0.7 ({cells} cells), {op}, 1 GPU",
                     "cells": cells, "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "equivalent_gvoxels_s": float(n) ** 3 / (ms * 1e-3) / 1e9,
                     "ms_per_step": ms, "steps": steps, "warmup": warmup, "vertices": nV, "triangles": nT, "mesh_verified": ok,
                     "whole_path_frac": (40.0 * cells + 12.0 * nV + 12.0 * nT) / (ms * 1e-3) / 1e9 / peak, "clocks": clocks})
    del band
    torch.cuda.empty_cache()
    return recs


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

    wl_name = args.workload or "c3"
    wl = WORKLOADS[wl_name]
    n = wl["n"]
    fn = field_fn(wl["field"])
    lib = _lib.lib()
    peak, peak_src = load_peaks()
    voxels = float(n) ** 3
    extra = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        rec, grid, prof = single_gpu_record(iso, lib, wl_name, dev, args.steps, args.warmup, peak, keep=True)
        ms_step, clocks, first_ms = rec["ms_per_step"], rec["clocks"], rec["first_call_ms"]
        nV, nT, csum = rec["vertices"], rec["triangles"], int(rec["mesh_checksum"], 16)
        local_vox = voxels
        parallelism = "single GPU"

        # ---- end to end through the public API with HOST buffers: full field from pinned memory every step
        host = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
        view = grid.values_view()
        for a in range(0, n, 64):
            host[a:a + 64].copy_(view[a:a + 64])
        hv_pin = torch.empty((nV + nV // 8 + 1024, 3), dtype=torch.float32, pin_memory=True)   # pinned landing buffers for the mesh
        hf_pin = torch.empty((nT + nT // 8 + 1024, 3), dtype=torch.int32, pin_memory=True)

        def e2e_step():
            view.copy_(host, non_blocking=True)           # H2D of this step's input, straight into the grid's storage (values_view())
            vv, ff = iso.marching_cubes(grid)
            hv, hf = hv_pin[: vv.shape[0]], hf_pin[: ff.shape[0]]
            hv.copy_(vv, non_blocking=True)               # D2H of the step's result
            hf.copy_(ff, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return hv, hf

        e2e_step()
        torch.cuda.synchronize()
        k = 3 if n >= 2048 else max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            hv, hf = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / k
        e2e = {"value": voxels / dt / 1e9, "unit": "Gvoxels/s", "ms_per_step": dt * 1e3, "steps": k,
               "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": int(hv.numel() * 4 + hf.numel() * 4)}
        del host, grid, view, hv, hf, hv_pin, hf_pin
        torch.cuda.empty_cache()
    else:
        from isoext_b200 import dist as idist

        def make_slabs(cuts=None):
            sg = idist.SlabGrid([n, n, n], group=dist.group.WORLD, cuts=cuts)
            x0, x1 = sg.owned_point_range()
            build_field_gpu(fn, n, x0, x1, dev, out=sg.owned_values())
            return sg

        def totals(v, f):
            """Global V / T totals and the checksum of the concatenated mesh (all_reduce of per-rank parts)."""
            mine = torch.tensor([v.shape[0], f.shape[0]], dtype=torch.int64, device=dev)
            allc = torch.empty(world * 2, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allc, mine)
            allc = allc.view(world, 2).cpu()
            c = mesh_checksum(v, f, int(allc[:rank, 0].sum()), int(allc[:rank, 1].sum()))
            ct = torch.tensor([c - (1 << 64) if c >= (1 << 63) else c], dtype=torch.int64, device=dev)
            dist.all_reduce(ct)
            return int(allc[:, 0].sum()), int(allc[:, 1].sum()), int(ct.item()) & 0xFFFFFFFFFFFFFFFF

        def max_over_ranks(x):
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # ---- headline: EVEN slabs (no foreknowledge)
        sg = make_slabs()
        barrier()
        t0 = time.perf_counter()
        v0, f0 = idist.marching_cubes(sg)
        barrier()
        first_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        sampler = ClockSampler(local_rank)
        ms_total, (v, f), prof, clocks = timed_loop(lambda: idist.marching_cubes(sg), args.steps, args.warmup, barrier, sampler, lib)
        ms_step = max_over_ranks(ms_total) / args.steps
        nV, nT, csum = totals(v, f)
        local_vox = float(sg.local_points())
        parallelism = f"dim-0 slabs x{world}, even cuts; halo pull + vertex-id bases as kernels over NVLink peer memory"

        # ---- end to end: each rank uploads its own slab from pinned host memory and reads its mesh part back
        host = sg.owned_values().cpu().pin_memory()
        dbuf = torch.empty_like(sg.owned_values())
        hv_pin = torch.empty((v.shape[0] + v.shape[0] // 8 + 1024, 3), dtype=torch.float32, pin_memory=True)
        hf_pin = torch.empty((f.shape[0] + f.shape[0] // 8 + 1024, 3), dtype=f.dtype, pin_memory=True)

        def e2e_step():
            dbuf.copy_(host, non_blocking=True)
            sg.set_owned_values(dbuf)                     # (guards the neighbours' halo pulls of the previous epoch)
            vv, ff = idist.marching_cubes(sg)
            hv, hf = hv_pin[: vv.shape[0]], hf_pin[: ff.shape[0]]
            hv.copy_(vv, non_blocking=True)
            hf.copy_(ff, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return hv, hf

        e2e_step()
        barrier()
        k = 3
        t0 = time.perf_counter()
        for _ in range(k):
            hv, hf = e2e_step()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / k)
        nb = torch.tensor([host.numel() * 4, hv.numel() * 4 + hf.numel() * 4], device=dev, dtype=torch.int64)
        dist.all_reduce(nb)
        e2e = {"value": voxels / dt / 1e9, "unit": "Gvoxels/s", "ms_per_step": dt * 1e3, "steps": k,
               "h2d_bytes_per_step": int(nb[0].item()), "d2h_bytes_per_step": int(nb[1].item())}
        del host, dbuf, hv_pin, hf_pin

        # ---- beside it: slab cuts balanced by the measured load of the previous extraction (time-series use)
        if not args.no_balanced:
            hist = idist.vertex_layer_histogram(v0, n, -1.0, 1.0, dist.group.WORLD)
            plane_s, vertex_s = 4.0 * n * n / (peak * 1e9), 0.43e-9
            cuts = idist.balanced_cuts((plane_s + vertex_s * hist).tolist(), world, ghost_cost=(vertex_s * hist).tolist())
            del v0, f0, v, f
            sg.close()
            sg = make_slabs(cuts)
            idist.marching_cubes(sg)
            ms_b, (vb, fb), _, _ = timed_loop(lambda: idist.marching_cubes(sg), args.steps, args.warmup, barrier)
            ms_b = max_over_ranks(ms_b) / args.steps
            bV, bT, bsum = totals(vb, fb)
            extra["balanced_cuts"] = {"ms_per_step": ms_b, "value": voxels / (ms_b * 1e-3) / 1e9, "unit": "Gvoxels/s", "cuts": cuts,
                                      "cost_model": {"plane_s": plane_s, "vertex_s": vertex_s},
                                      "vertices": bV, "triangles": bT, "mesh_checksum": f"{bsum:016x}",
                                      "mesh_verified": verify_mesh(wl_name, bV, bT, bsum),
                                      "note": "cuts from the per-layer vertex histogram of a PREVIOUS extraction of the same field"}
        sg.close()

    value = voxels / (ms_step * 1e-3) / 1e9
    if rank == 0:
        verified = verify_mesh(wl_name, nV, nT, csum)
        # roofline of the dominant kernel (k_signbits: the one volume-sized HBM stream); 4 B/voxel algorithmic
        sm_ms, sm_n, launches = prof
        launches_per_step = max(1.0, sm_n / float(args.steps))
        vox_per_launch = local_vox / launches_per_step
        k_ms = sm_ms / max(1, sm_n)
        achieved = 4.0 * vox_per_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        traffic = None
        tprof = ROOT / "profiles" / "k_signbits_traffic.json"
        if tprof.exists():
            try:
                traffic = json.loads(tprof.read_text()).get(wl_name)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
            "config": workload_config(wl_name),
            "parallelism": parallelism,
            "mesh": {"vertices": nV, "triangles": nT, "checksum": f"{csum:016x}", "verified": verified},
            "first_call_ms": first_ms,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_signbits", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": k_ms, "kernel_launches_timed": int(sm_n), "kernel_launches_per_step": launches_per_step,
                         "algorithmic_bytes_per_launch": 4.0 * vox_per_launch,
                         "whole_path_frac": (4.0 * voxels + 12.0 * nV + 12.0 * nT) / (ms_step * 1e-3) / 1e9 / peak / world},
        }
        line.update(extra)
        if world == 1 and not args.no_sub_records and wl_name == "c3":
            subs = []
            for name in ("c2", "t1024"):
                try:
                    r, _, _ = single_gpu_record(iso, lib, name, dev, args.steps, args.warmup, peak)
                    subs.append(r)
                except Exception as exc:    # the headline line must still be printed
                    subs.append({"workload": name, "unavailable": repr(exc)[:200]})
            try:
                subs.extend(other_config_records(iso, lib, dev, args.steps, args.warmup, peak))
            except Exception as exc:
                subs.append({"workload": "c4 / c5", "unavailable": repr(exc)[:200]})
            line["sub_records"] = subs
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(wl["field"])
        print(json.dumps(line), flush=True)
    if world > 1:
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
    wl_name = args.workload or "c3"

    def time_ref(name, n, steps, warmup, sampler=None):
        fn = field_fn(WORKLOADS[name]["field"])
        vals = build_field_gpu(fn, n, 0, n, dev)
        grid = ref.UniformGrid([n, n, n])
        grid.set_values(vals)

        def step():
            v, f, nv, nf = ref.marching_cubes_timed_raw(grid, 0.0, "nagae")
            ref.free(v); ref.free(f)
            return nv, nf

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            nv, nf = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        clocks = sampler.stop() if sampler else None
        del grid, vals
        torch.cuda.empty_cache()
        return ms, nv, nf, clocks

    n_full = WORKLOADS[wl_name]["n"]
    n = min(n_full, 512)     # bounded sample: the reference is invalid above 813 points/axis (uint overflow)
    sample = (f"full workload ({n}^3)" if n == n_full else
              f"{n}^3 sampling of the same field over the same AABB (the reference cannot represent {n_full}^3)")
    ms_step, nv, nf, clocks = time_ref(wl_name, n, args.steps, args.warmup, ClockSampler(local_rank))
    value = float(n) ** 3 / (ms_step * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "Gvoxels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": workload_config(wl_name),
            "same_workload": bool(n == n_full and world == 1),
            "sample": {"shape": [n, n, n], "vertices": nv, "triangles": nf, "what": sample,
                       "note": "reference = its own CUDA extension (thrust pipeline) on ONE GPU; it has no CPU path and no multi-GPU path"},
            "clocks": clocks,
            "cpu_baseline": {"value": value, "unit": "Gvoxels/s", "cores": 0, "kind": "reference",
                             "sample": sample + "; runs on the GPU: the reference has no CPU implementation of this path"},
            "e2e": {"value": value, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if wl_name == "c3" and not args.no_sub_records:
        ms2, nv2, nf2, _ = time_ref("c2", 512, max(3, min(args.steps, 10)), 1)
        line["sub_records"] = [{"workload": f"c2: {WORKLOADS['c2']['desc']}", "shape": [512] * 3, "same_workload": True,
                                "value": 512.0 ** 3 / (ms2 * 1e-3) / 1e9, "unit": "Gvoxels/s", "ms_per_step": ms2,
                                "vertices": nv2, "triangles": nf2}]
        try:
            line["sub_records"] += reference_other_configs(ref, dev)
        except Exception as exc:
            line["sub_records"].append({"workload": "c4 / c5", "unavailable": repr(exc)[:200]})
    print(json.dumps(line), flush=True)


def reference_other_configs(ref, dev, steps=3, warmup=1):
    """The reference CUDA build on BASELINE.json configs[3] (512^3 CSG dual_contouring) and configs[4] at the size its
    documentation uses (1024^3-equivalent sphere band: marching_cubes, dual_contouring) -- full size, identical inputs to
    the sub-records of our arm (the band is populated with our GPU population, the reference's recipe being a Python loop
    over 10,738 chunks)."""
    import torch
    import isoext_b200 as iso
    from isoext_b200 import sdf as S

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, out

    recs = []
    n = 512
    vals = build_field_gpu(field_fn("csg"), n, 0, n, dev)
    rg = ref.UniformGrid([n] * 3)
    rg.set_values(vals)
    ms, (v, f) = timed(lambda: ref.dual_contouring(rg))
    recs.append({"workload": "c4: 512^3 dense CSG box-minus-sphere dual_contouring", "same_workload": True, "ms_per_step": ms,
                 "value": float(n) ** 3 / (ms * 1e-3) / 1e9, "unit": "Gvoxels/s", "vertices": int(v.shape[0]), "triangles": int(f.shape[0])})
    del rg, vals, v, f
    n = 1024
    band = iso.SparseGrid([n] * 3)
    band.populate_from_dense(iso.ImplicitGrid([n] * 3, S.SphereSDF(0.7)))
    cells, vals8 = band.get_cell_indices(), band.get_values()
    rg = ref.SparseGrid([n] * 3)
    rg.add_cells(cells.to(torch.int32).contiguous())
    rg.set_values(vals8)
    for op, fn in (("marching_cubes", lambda: ref.marching_cubes(rg)), ("dual_contouring", lambda: ref.dual_contouring(rg))):
        ms, (v, f) = timed(fn)
        recs.append({"workload": f"c5: SparseGrid narrow band, 1024^3-equivalent sphere r=0.7 ({cells.numel()} cells), {op}",
                     "same_workload": True, "ms_per_step": ms, "value": cells.numel() / (ms * 1e-3) / 1e6, "unit": "Mcells/s",
                     "vertices": int(v.shape[0]), "triangles": int(f.shape[0])})
    return recs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--no-balanced", action="store_true", help="N > 1: skip the extra timing with load-balanced slab cuts")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="N = 1: skip the c2 / 1024^3 sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
