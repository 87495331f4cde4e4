"""``marching_cubes`` -- host-side mirror of the reference binding (src/isoext_ext.cu:95-109) over
the sm_100a kernels in csrc/mc_dense.cu (dense) and csrc/sparse.cu (sparse)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .grid import UniformGrid, _stream_ptr

METHODS = {"nagae": 0, "lorensen": 1}


def _method_id(method: str) -> int:
    if method not in METHODS:
        raise RuntimeError("Unknown method: " + str(method))  # src/mc/base.cu:23-25
    return METHODS[method]


def _initial_cap(shape) -> int:
    X, Y, Z = shape
    return int(min(X * Y * Z, max(1 << 16, 4 * (X * Y + Y * Z + X * Z))))


def mc_dense_raw(values: torch.Tensor, shape, aabb_min, aabb_max, level, method_id, ws, cap_hint=0, x_offset=0,
                 x_global=None, emit_range=None, x_thresholds=(-math.inf, math.inf), hints=None, halo=None, sdf_prog=None):
    """Run the dense pipeline on a (X,Y,Z) float32 CUDA tensor.

    Returns ``(V_own, F, n_lo, n_hi, cap_used)``: ``V_own`` the position-sorted welded vertices OWNED by
    the slab (x thresholds; everything on a single GPU), ``F`` (T,3) int32 ids in the slab's extended id
    space: ``[0,n_lo)`` vertices owned by the previous slab, ``[n_lo,n_hi)`` = rows of ``V_own``, above
    ``n_hi`` owned by the next slab (``n_lo = 0``, ``n_hi = len(V_own)`` on a single GPU).
    ``(None, None, 0, 0, cap)`` if empty.

    ``hints``: a dict owned by the caller (one per grid) remembering the sizes of the previous extraction.
    With hints the single-call fast path (``isoext_mc_dense_run``: one stream sync, no host round trip
    between the phases) is tried first; any capacity miss falls back to count + emit.

    ``sdf_prog``: device image of an analytic program (``sdf.SdfProgram.device``); ``values`` is then ``None`` and the
    kernels evaluate the program instead of loading values (``grid.ImplicitGrid``).

    ``halo``: ``(planes_lo, planes_hi, event)`` -- the first / last x planes of ``values`` are still being written on
    another stream that records ``event`` (a ``torch.cuda.Event``) when done; the fast path starts streaming the planes
    in between at once (slab halo pull overlapped with the volume stream), every other path waits for the event first.
    """
    lib = _lib.lib()
    X, Y, Z = shape
    xg = X if x_global is None else x_global
    lo, hi = (0, X - 1) if emit_range is None else emit_range
    # per-grid cache of what does not change between calls (ctypes arrays of the box, workspace sizes): the host side
    # of a 300 us extraction is worth keeping short
    cache = hints.setdefault("_cache", {}) if hints is not None else {}
    box_key = (tuple(float(v) for v in aabb_min), tuple(float(v) for v in aabb_max))
    if cache.get("box_key") != box_key:
        cache["box_key"], cache["box"] = box_key, (_lib.f3(aabb_min), _lib.f3(aabb_max))
    amin, amax = cache["box"]
    dev = values.device if values is not None else sdf_prog.device
    vptr = values.data_ptr() if values is not None else None
    sptr = sdf_prog.data_ptr() if sdf_prog is not None else None
    stream = _stream_ptr()
    counts = (C.c_int64 * 8)()
    cap = max(int(cap_hint), _initial_cap(shape))
    thr_lo, thr_hi = float(x_thresholds[0]), float(x_thresholds[1])

    if hints is not None and hints.get("Vc", 0) > 0 and hints.get("T", 0) > 0:
        cand_cap = hints["Vc"] + (hints["Vc"] >> 6) + 16
        tri_cap = hints["T"] + (hints["T"] >> 6) + 16
        nb_hint = hints.get("n_big", 0)
        big_cap = min(cand_cap, nb_hint + (nb_hint >> 6) + 16) if nb_hint > 0 else 0
        ws_key, sc_key = ("ws", X, Y, Z, cap), ("sc", cand_cap)
        if ws_key not in cache:
            cache[ws_key] = lib.isoext_mc_dense_workspace_bytes(X, Y, Z, cap)
        if sc_key not in cache:
            if len(cache) > 64:
                cache.clear()
                cache["box_key"], cache["box"] = box_key, (amin, amax)
                cache[ws_key] = lib.isoext_mc_dense_workspace_bytes(X, Y, Z, cap)
            cache[sc_key] = lib.isoext_mc_dense_scratch_bytes(cand_cap)
        wsbuf = ws.get("mc_ws", cache[ws_key], dev)
        scratch = ws.get("mc_scratch", cache[sc_key], dev)
        V = torch.empty((cand_cap, 3), dtype=torch.float32, device=dev)
        F = torch.empty((tri_cap, 3), dtype=torch.int32, device=dev)
        h_lo, h_hi, h_ev = (0, 0, None) if halo is None else (int(halo[0]), int(halo[1]), halo[2].cuda_event)
        halo = None      # consumed: whatever follows this call runs behind the event
        rc = lib.isoext_mc_dense_run(vptr, X, Y, Z, x_offset, xg, amin, amax, float(level), method_id, lo, hi,
                                     wsbuf.data_ptr(), wsbuf.numel(), cap, scratch.data_ptr(), scratch.numel(), cand_cap, tri_cap,
                                     big_cap, 1 if hints.get("radix") else 0, thr_lo, thr_hi, h_lo, h_hi, h_ev, V.data_ptr(),
                                     F.data_ptr(), sptr, stream, counts)
        hints["radix"] = int(counts[7]) > 0     # the sort's radix last resort is only enqueued when it was needed last time
        if rc == 0:
            S, T, Vc = int(counts[0]), int(counts[1]), int(counts[2])
            hints.update(Vc=Vc, T=T, n_big=int(counts[3]))
            if Vc == 0:
                return None, None, 0, 0, cap
            return V[int(counts[5]):int(counts[6])], F[:T], int(counts[5]), int(counts[6]), max(cap, S)
        if rc != 1:
            _lib.check(rc)
        if int(counts[0]) > cap:
            cap = int(counts[0]) + 1024
        # fall through to the two-phase path

    if halo is not None:
        torch.cuda.current_stream().wait_event(halo[2])
    while True:
        nbytes = lib.isoext_mc_dense_workspace_bytes(X, Y, Z, cap)
        if nbytes == 0:
            raise RuntimeError(_lib.last_error())
        wsbuf = ws.get("mc_ws", nbytes, dev)
        rc = lib.isoext_mc_dense_count(vptr, X, Y, Z, x_offset, xg, amin, amax, float(level), method_id,
                                       lo, hi, wsbuf.data_ptr(), wsbuf.numel(), cap, sptr, stream, counts)
        if rc == _lib.E_CAPACITY:
            cap = int(counts[0]) + 1024
            continue
        _lib.check(rc)
        break
    S, T, Vc, n_big = int(counts[0]), int(counts[1]), int(counts[2]), int(counts[3])
    if hints is not None:
        hints.update(Vc=Vc, T=T, n_big=n_big)
    if Vc == 0:   # no vertex at all (T == 0 alone is not enough: a slab may own vertices that only
        return None, None, 0, 0, cap   # its neighbour's faces reference)
    sbytes = lib.isoext_mc_dense_scratch_bytes(Vc)
    scratch = ws.get("mc_scratch", sbytes, dev)
    V = torch.empty((Vc, 3), dtype=torch.float32, device=dev)
    F = torch.empty((T, 3), dtype=torch.int32, device=dev)
    out = (C.c_int64 * 4)()
    _lib.check(lib.isoext_mc_dense_emit(vptr, X, Y, Z, x_offset, xg, amin, amax, float(level), method_id,
                                        lo, hi, wsbuf.data_ptr(), wsbuf.numel(), cap, scratch.data_ptr(), scratch.numel(),
                                        Vc, n_big, thr_lo, thr_hi, V.data_ptr(), F.data_ptr(), sptr,
                                        stream, out))
    n_lo, n_hi = int(out[1]), int(out[2])
    return V[n_lo:n_hi], F, n_lo, n_hi, max(cap, S)


def marching_cubes(grid, level: float = 0.0, method: str = "nagae"):
    """Extract the ``level`` iso-surface of ``grid`` (src/isoext_ext.cu:95-109).

    Returns ``(v, f)``: ``v`` (V, 3) float32 positions in the reference's order (lexicographic
    (x, y, z) of the welded positions, src/utils.cu:49-55), ``f`` (T, 3) int32 in ascending-cell x
    LUT order -- or ``(None, None)`` when no triangle exists (src/isoext_ext.cu:47-49).
    """
    mid = _method_id(method)
    from .grid import ImplicitGrid
    if isinstance(grid, (UniformGrid, ImplicitGrid)):
        args = (grid._values, grid.shape, grid.aabb_min, grid.aabb_max, level, mid, grid._ws)
        kw = dict(cap_hint=grid._cap_hint, hints=grid._hints, sdf_prog=getattr(grid, "_prog", None))
        if torch.cuda.current_device() == grid.device.index:       # (entering a device context costs several us)
            v, f, _, _, cap = mc_dense_raw(*args, **kw)
        else:
            with torch.cuda.device(grid.device):
                v, f, _, _, cap = mc_dense_raw(*args, **kw)
        grid._cap_hint = cap
        if f is None or f.shape[0] == 0:
            return None, None   # src/isoext_ext.cu:47-49: empty arrays become None
        return v, f
    from .sparse import SparseGrid, mc_sparse
    if isinstance(grid, SparseGrid):
        return mc_sparse(grid, level, mid)
    raise TypeError("marching_cubes: grid must be a UniformGrid or SparseGrid")
