"""``Intersection``, ``get_intersection`` and ``dual_contouring`` -- host-side mirror of the reference
bindings (src/isoext_ext.cu:305-378) over the sm_100a kernels in csrc/dc_dense.cu."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .grid import UniformGrid, _expect_cuda, _stream_ptr
from .mc import _initial_cap


class Intersection:
    """Edge/surface crossings of a grid (include/its.cuh:6-26).  Not constructible from Python, like the
    reference's; created by :func:`get_intersection`.

    ``get_points()`` / ``get_normals()`` are (I, 3) float32 with one row per (active cell, crossing edge)
    pair -- active cells ascending, edges 0..11 -- so an interior crossing appears once per incident
    cell (64^3 cube: 24,576 = 4 x 6,144).  Returned tensors are views of torch-owned storage (the
    reference hands ownership to the returned tensor and keeps a dangling view; not replicated).
    """

    def __init__(self, _token=None):
        if _token is not Intersection._TOKEN:
            raise TypeError("Intersection cannot be constructed directly; use get_intersection()")

    _TOKEN = object()

    @classmethod
    def _make(cls, **kw):
        self = cls(cls._TOKEN)
        self.__dict__.update(kw)
        return self

    def get_points(self) -> torch.Tensor:
        return self.points

    def get_normals(self) -> torch.Tensor:
        """Before normals exist the reference returns uninitialised memory (include/its.cuh:15-17);
        here the buffer is zero-filled."""
        return self.normals

    def has_normals(self) -> bool:
        return self._has_normals

    def set_normals(self, new_normals: torch.Tensor) -> None:
        _expect_cuda(new_normals, torch.float32, ndim=2, last=3, what="new_normals")
        if new_normals.shape[0] != self.points.shape[0]:
            raise RuntimeError("Cannot set values with different shapes")   # include/ndarray.cuh:79-84
        self.normals = new_normals.detach().clone()
        self._has_normals = True

    def _copy(self):
        """By-value copy, as the binding takes std::optional<Intersection> (src/isoext_ext.cu:347-351)."""
        d = dict(self.__dict__)
        d["normals"] = self.normals.clone()
        return Intersection._make(**d)


def _margin(n: int) -> int:
    return n + (n >> 6) + 16


def its_dense_raw(values, shape, aabb_min, aabb_max, level: float, compute_normals: bool, ws, cap_hint=0,
                  x_offset=0, x_global=None, sdf_prog=None, hints=None):
    """Intersections of a dense (X, Y, Z) float32 CUDA field (a whole grid, or the extended slab of a sharded grid:
    ``x_offset`` = global index of local plane 0, ``x_global`` = points along x of the whole grid).
    Returns ``(Intersection, entry capacity used)``."""
    lib = _lib.lib()
    X, Y, Z = shape
    xg = X if x_global is None else int(x_global)
    dev = values.device if values is not None else sdf_prog.device
    vptr = values.data_ptr() if values is not None else None
    sptr = sdf_prog.data_ptr() if sdf_prog is not None else None
    amin, amax = _lib.f3(aabb_min), _lib.f3(aabb_max)
    stream = _stream_ptr()
    counts = (C.c_int64 * 4)()
    cap = max(int(cap_hint), _initial_cap(shape))
    row_start = torch.empty(X * Y + 2, dtype=torch.int32, device=dev)
    # Outputs sized by the previous extraction of the grid (``hints``) are allocated BEFORE the counting call, so that
    # the host does nothing but slice between its synchronisation and the launch of the emit phase.
    pre = None
    if hints is not None and hints.get("its_n", 0) > 0:
        ci, cc = _margin(hints["its_n"]), _margin(hints["its_cells"])
        pre = (ci, cc, torch.empty((ci, 3), dtype=torch.float32, device=dev),
               (torch.empty if compute_normals else torch.zeros)((ci, 3), dtype=torch.float32, device=dev),
               torch.zeros(cc + 1, dtype=torch.int32, device=dev), torch.empty(cc, dtype=torch.int64, device=dev))
    isout_buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    while True:
        nbytes = lib.isoext_its_dense_workspace_bytes(X, Y, Z, cap)
        if nbytes == 0:
            raise RuntimeError(_lib.last_error())
        wsbuf = ws.get("its_ws", nbytes, dev)
        entries = torch.empty((cap + 1, 2), dtype=torch.int32, device=dev)
        cellslot = torch.empty(cap, dtype=torch.int32, device=dev)
        its_off = torch.empty(cap, dtype=torch.int32, device=dev)
        rc = lib.isoext_its_dense_count(vptr, X, Y, Z, x_offset, xg, amin, amax, float(level), wsbuf.data_ptr(),
                                        wsbuf.numel(), cap, entries.data_ptr(), row_start.data_ptr(), cellslot.data_ptr(),
                                        its_off.data_ptr(), sptr, stream, counts)
        if rc == _lib.E_CAPACITY:
            cap = int(counts[0]) + 1024
            continue
        _lib.check(rc)
        break
    S, n_cells, n_its = int(counts[0]), int(counts[1]), int(counts[2])
    entries, cellslot, its_off = entries[:S + 1], cellslot[:max(S, 1)], its_off[:max(S, 1)]
    if pre is not None and n_its <= pre[0] and n_cells <= pre[1]:
        points, normals, cell_offsets, cell_indices = pre[2][:n_its], pre[3][:n_its], pre[4][:n_cells + 1], pre[5][:n_cells]
    else:
        points = torch.empty((n_its, 3), dtype=torch.float32, device=dev)
        # every normal is written when they are computed; otherwise get_normals() must read zeros, not garbage
        normals = (torch.empty if compute_normals else torch.zeros)((n_its, 3), dtype=torch.float32, device=dev)
        cell_offsets = torch.zeros(n_cells + 1, dtype=torch.int32, device=dev)
        cell_indices = torch.empty(n_cells, dtype=torch.int64, device=dev)
    isout = isout_buf[:max(S, 1)] if isout_buf.numel() >= max(S, 1) else torch.empty(max(S, 1), dtype=torch.uint8, device=dev)
    if hints is not None:
        hints.update(its_n=n_its, its_cells=n_cells)
    _lib.check(lib.isoext_its_dense_emit(vptr, X, Y, Z, x_offset, xg, amin, amax, float(level),
                                         int(bool(compute_normals)), entries.data_ptr(), S, cellslot.data_ptr(),
                                         its_off.data_ptr(), n_cells, n_its, points.data_ptr(), normals.data_ptr(),
                                         isout.data_ptr(), cell_offsets.data_ptr(), cell_indices.data_ptr(), sptr, stream))
    its = Intersection._make(kind="dense", shape=tuple(shape), aabb_min=aabb_min, aabb_max=aabb_max, level=float(level),
                             entries=entries, row_start=row_start, cellslot=cellslot, its_off=its_off, isout=isout,
                             n_entries=S, n_cells=n_cells, points=points, normals=normals, cell_offsets=cell_offsets,
                             cell_indices=cell_indices, x_offset=int(x_offset), x_global=xg, _has_normals=bool(compute_normals))
    return its, max(cap, S)


def _its_dense(grid, level: float, compute_normals: bool) -> Intersection:
    its, cap = its_dense_raw(grid._values, grid.shape, grid.aabb_min, grid.aabb_max, level, compute_normals, grid._ws,
                             cap_hint=grid._cap_hint, sdf_prog=getattr(grid, "_prog", None), hints=getattr(grid, "_hints", None))
    grid._cap_hint = cap
    return its


def get_intersection(grid, level: float = 0.0, compute_normals: bool = False) -> Intersection:
    """Edge/iso-surface crossings of every active cell (src/isoext_ext.cu:329-343, src/its.cu:93-159)."""
    from .grid import ImplicitGrid
    if isinstance(grid, (UniformGrid, ImplicitGrid)):
        with torch.cuda.device(grid.device):
            return _its_dense(grid, level, compute_normals)
    from .sparse import SparseGrid, its_sparse
    if isinstance(grid, SparseGrid):
        return its_sparse(grid, level, compute_normals)
    raise TypeError("get_intersection: grid must be a UniformGrid or SparseGrid")


def _normals_dense(grid, its: Intersection) -> None:
    X, Y, Z = grid.shape
    prog = getattr(grid, "_prog", None)
    _lib.check(_lib.lib().isoext_its_dense_normals(grid._values.data_ptr() if grid._values is not None else None, X, Y, Z,
                                                   its.x_offset, its.x_global, _lib.f3(grid.aabb_min),
                                                   _lib.f3(grid.aabb_max), its.entries.data_ptr(), its.n_entries,
                                                   its.cellslot.data_ptr(), its.its_off.data_ptr(), its.points.data_ptr(),
                                                   its.normals.data_ptr(), prog.data_ptr() if prog is not None else None,
                                                   _stream_ptr()))
    its._has_normals = True


def dc_dense_raw(grid, its: Intersection, reg: float, svd_tol: float, want_quads: bool = False,
                 dual_v_in: torch.Tensor | None = None, emit_range=None, x_thresholds=(-math.inf, math.inf), with_counts=False):
    """Returns (v, f, dual_v, quads): welded mesh + per-active-cell dual vertices (+ oriented quads).
    ``grid`` only lends its device and workspace cache (a UniformGrid or a dist.SlabGrid); the geometry comes from
    ``its`` (whole grid or extended slab).  ``dual_v_in`` (n_cells, 3) replaces the solved dual vertices before the
    quad/split/weld stage (parity tests feed the reference's own dual vertices through it to compare everything
    downstream of the solve).  Slabs: ``emit_range`` = local point planes [lo, hi) whose edges emit quads,
    ``x_thresholds`` = ownership thresholds; ``with_counts`` appends (n_lo, n_hi) and then ``v`` holds ALL welded
    vertices of the slab."""
    lib = _lib.lib()
    X, Y, Z = its.shape
    dev = its.points.device
    amin, amax = _lib.f3(its.aabb_min), _lib.f3(its.aabb_max)
    xo, xg = its.x_offset, its.x_global
    lo, hi = (0, X) if emit_range is None else emit_range
    thr_lo, thr_hi = float(x_thresholds[0]), float(x_thresholds[1])
    stream = _stream_ptr()
    S, n_cells = its.n_entries, its.n_cells
    dual_v = torch.empty((n_cells, 3), dtype=torch.float32, device=dev)
    empty = (None, None, dual_v, None) + ((0, 0) if with_counts else ())
    if S == 0 or n_cells == 0:
        return empty
    ws = grid._ws.get("dc_ws", lib.isoext_dc_dense_workspace_bytes(S, n_cells), dev)
    hints = getattr(grid, "_hints", None)
    pre = None                  # outputs sized by the previous extraction, allocated before the counting call (see its_dense_raw)
    if hints is not None and hints.get("dc_Vc", 0) > 0:
        cv, cq = _margin(hints["dc_Vc"]), _margin(hints["dc_Q"])
        pre = (cv, cq, torch.empty((cv, 3), dtype=torch.float32, device=dev), torch.empty((2 * cq, 3), dtype=torch.int32, device=dev),
               torch.empty((cq, 4), dtype=torch.int32, device=dev) if want_quads else None)
    counts = (C.c_int64 * 4)()
    _lib.check(lib.isoext_dc_dense_count(X, Y, Z, xo, xg, amin, amax, lo, hi, its.entries.data_ptr(), S, its.row_start.data_ptr(),
                                         its.cellslot.data_ptr(), its.its_off.data_ptr(), n_cells, its.points.data_ptr(),
                                         its.normals.data_ptr(), float(reg), float(svd_tol), dual_v.data_ptr(), ws.data_ptr(),
                                         ws.numel(), stream, counts))
    Q, Vc = int(counts[0]), int(counts[1])
    if Vc == 0 or (Q == 0 and not with_counts):
        return empty
    if dual_v_in is not None:
        dual_v.copy_(dual_v_in)
    scratch = grid._ws.get("dc_scratch", lib.isoext_dc_dense_scratch_bytes(Vc), dev)
    if hints is not None:
        hints.update(dc_Vc=Vc, dc_Q=Q)
    if pre is not None and Vc <= pre[0] and Q <= pre[1]:
        V, F, quads = pre[2][:Vc], pre[3][:2 * Q], (pre[4][:Q] if want_quads else None)
    else:
        V = torch.empty((Vc, 3), dtype=torch.float32, device=dev)
        F = torch.empty((2 * Q, 3), dtype=torch.int32, device=dev)
        quads = torch.empty((Q, 4), dtype=torch.int32, device=dev) if want_quads else None
    out = (C.c_int64 * 4)()
    _lib.check(lib.isoext_dc_dense_emit(X, Y, Z, xo, xg, amin, amax, lo, hi, thr_lo, thr_hi, its.entries.data_ptr(), S,
                                        its.row_start.data_ptr(), its.cellslot.data_ptr(), its.isout.data_ptr(), n_cells,
                                        dual_v.data_ptr(), ws.data_ptr(), ws.numel(), scratch.data_ptr(), scratch.numel(), Vc,
                                        V.data_ptr(), F.data_ptr(), quads.data_ptr() if want_quads else None, stream, out))
    res = (V[:int(out[0])], F, dual_v, quads)
    return res + ((int(out[1]), int(out[2])) if with_counts else ())


def dual_contouring(grid, level: float = 0.0, intersection: Intersection | None = None, reg: float = 1e-2,
                    svd_tol: float = 1e-6):
    """Dual contouring (src/isoext_ext.cu:345-378, src/dc.cu:161-218): one vertex per active cell from a
    regularised QEF, one quad (two triangles, split along the shorter diagonal) per sign-change edge.

    Returns ``(v, f)``: (V, 3) float32 in lexicographic position order and (2Q, 3) int32, or
    ``(None, None)`` if there is no quad.  A passed ``intersection`` is not modified; if it carries no
    normals they are computed from the grid values (trilinear central differences).
    Deviations: (1) quads with a neighbour cell outside the grid are skipped on the upper faces as well (the
    reference only checks the lower faces, include/utils.cuh:128-135, and reads out of bounds).  (2) On a SparseGrid
    whose neighbouring cells carry INCONSISTENT values for a shared corner (``set_values`` allows it), the existence
    and orientation of the quad of an edge are taken from the cell whose corner 0 is the edge origin; the reference
    collects the sign-change edges of all cells around the edge and takes ``is_out`` from the first in sorted order
    (src/its.cu get_its_op + src/grid/sparse.cu:181-246).  With consistent values -- every field sampled from one
    function -- the two are identical (tested); marching cubes welds by position and has no such caveat."""
    from .grid import ImplicitGrid
    if isinstance(grid, (UniformGrid, ImplicitGrid)):
        with torch.cuda.device(grid.device):
            its = intersection._copy() if intersection is not None else _its_dense(grid, level, True)
            if its.kind != "dense" or tuple(its.shape) != tuple(grid.shape):
                raise RuntimeError("intersection does not belong to this grid")
            if not its.has_normals():
                _normals_dense(grid, its)
            v, f, _, _ = dc_dense_raw(grid, its, reg, svd_tol)
        return v, f
    from .sparse import SparseGrid, dc_sparse
    if isinstance(grid, SparseGrid):
        return dc_sparse(grid, level, intersection, reg, svd_tol)
    raise TypeError("dual_contouring: grid must be a UniformGrid or SparseGrid")
