"""``SparseGrid`` and the sparse marching-cubes / intersection / dual-contouring wrappers -- host-side
mirror of the reference's SparseGrid binding (src/isoext_ext.cu:170-303, src/grid/sparse.cu) over the
sm_100a kernels in csrc/sparse.cu.

Everything goes through the C-ABI kernels: list maintenance (sorted-unique insert, set difference, filter
compaction: csrc/setops.cu) as well as whatever touches corner values or positions (csrc/sparse.cu).
Extension: cell-index tensors may be int64 (needed beyond INT_MAX points, e.g. 4096^3-equivalent
narrow bands); int32 in -> int32 out, exactly like the reference, otherwise int64."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .grid import FLT_MAX, Grid, _expect_cuda, _stream_ptr, _Workspace


def _idx_tensor(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dim() != 1 or t.dtype not in (torch.int32, torch.int64):
        raise TypeError(f"{what}: expected a 1-D int32 (or int64) CUDA tensor")
    return t.contiguous()


def _ids64(t: torch.Tensor) -> torch.Tensor:
    """int32 ids are the reference's API type; it reinterprets them as uint32 (NDArray<int>::cast<uint>)."""
    return (t.to(torch.int64) & 0xFFFFFFFF) if t.dtype == torch.int32 else t


def _setop(fn_name, ws: _Workspace, device, n_ws, out_n, *args):
    """Run one csrc/setops.cu entry point: (args..., out, workspace, bytes, stream) -> first n_out items of out."""
    lib = _lib.lib()
    out = torch.empty(max(int(out_n), 1), dtype=torch.int64, device=device)
    wsbuf = ws.get("setops", lib.isoext_setops_workspace_bytes(max(int(n_ws), 1)), device)
    n_out = C.c_int64(0)
    with torch.cuda.device(device):
        _lib.check(getattr(lib, fn_name)(*args, out.data_ptr(), wsbuf.data_ptr(), wsbuf.numel(), _stream_ptr(), C.byref(n_out)))
    return out[:n_out.value]


class SparseGrid(Grid):
    """Only the cells that were added exist; each carries its own 8 corner values
    (src/grid/sparse.cu:9-19).  ``shape`` = points per axis of the enclosing uniform grid; the cell index
    is ``x*(Y-1)*(Z-1) + y*(Z-1) + z`` (include/utils.cuh:43-47)."""

    def __init__(self, shape, aabb_min=(-1.0, -1.0, -1.0), aabb_max=(1.0, 1.0, 1.0), default_value=FLT_MAX, device=None):
        shape = [int(s) for s in shape]
        if len(shape) != 3 or len(aabb_min) != 3 or len(aabb_max) != 3:
            raise TypeError("shape, aabb_min and aabb_max must have three elements")
        if min(shape) < 2:
            raise RuntimeError("Grid shape must be at least 2 points per axis")
        self.shape = tuple(shape)
        self.aabb_min = tuple(float(v) for v in aabb_min)
        self.aabb_max = tuple(float(v) for v in aabb_max)
        self.default_value = float(default_value)
        _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._cells = torch.empty(0, dtype=torch.int64, device=self.device)     # sorted, unique
        self._values = torch.empty((0, 8), dtype=torch.float32, device=self.device)
        self._ws = _Workspace()
        self._int32_api = True

    # -- sizes ----------------------------------------------------------------------------------
    def get_num_cells(self) -> int:
        return int(self._cells.numel())

    def get_num_points(self) -> int:
        return 8 * self.get_num_cells()

    def _out_idx(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(torch.int32) if self._int32_api else t

    # -- list maintenance (src/grid/sparse.cu:71-126) ---------------------------------------------
    def _reset_values(self):
        self._values = torch.full((self._cells.numel(), 8), self.default_value, dtype=torch.float32, device=self.device)

    def add_cells(self, new_cell_indices: torch.Tensor) -> None:
        """Insert cells; like the reference this resets ALL values to the default (sparse.cu:94-95)."""
        t = _idx_tensor(new_cell_indices, "new_cell_indices")
        self._int32_api = self._int32_api and t.dtype == torch.int32
        both = torch.cat([self._cells, _ids64(t)])
        self._cells = _setop("isoext_ids_sort_unique", self._ws, self.device, both.numel(), both.numel(),
                             both.data_ptr(), both.numel()).clone()
        self._reset_values()

    def remove_cells(self, cell_indices: torch.Tensor) -> None:
        """Remove cells (set difference); resets all values to the default (sparse.cu:119-120)."""
        t = _ids64(_idx_tensor(cell_indices, "new_cell_indices")).contiguous()
        n = self._cells.numel()
        self._cells = _setop("isoext_ids_difference", self._ws, self.device, max(n, t.numel()), n,
                             self._cells.data_ptr(), n, t.data_ptr(), t.numel()).clone()
        self._reset_values()

    def get_cell_indices(self) -> torch.Tensor:
        return self._out_idx(self._cells.clone())

    def get_potential_cell_indices(self, chunk_size: int) -> list:
        """Chunks of candidate ids 0 .. X*Y*Z-1 -- the reference enumerates the POINT count, overshooting
        the (X-1)(Y-1)(Z-1) valid cells (src/grid/sparse.cu:128-142); kept for drop-in behaviour."""
        X, Y, Z = self.shape
        total, chunk = X * Y * Z, int(chunk_size)
        dt = torch.int32 if total <= 2147483647 else torch.int64
        return [torch.arange(s, min(s + chunk, total), dtype=dt, device=self.device) for s in range(0, total, chunk)]

    # -- geometry / values --------------------------------------------------------------------------
    def _points_of(self, cells64: torch.Tensor) -> torch.Tensor:
        X, Y, Z = self.shape
        out = torch.empty((cells64.numel(), 8, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().isoext_sparse_points(X, Y, Z, _lib.f3(self.aabb_min), _lib.f3(self.aabb_max),
                                                       cells64.data_ptr(), cells64.numel(), out.data_ptr(), _stream_ptr()))
        return out

    def get_points(self) -> torch.Tensor:
        """(N, 8, 3) corner positions of the active cells (src/grid/sparse.cu:38-43)."""
        return self._points_of(self._cells)

    def get_points_by_cell_indices(self, cell_indices: torch.Tensor) -> torch.Tensor:
        return self._points_of(_idx_tensor(cell_indices, "cell_indices").to(torch.int64))

    def get_values(self) -> torch.Tensor:
        return self._values.clone()

    def set_values(self, new_values: torch.Tensor) -> None:
        _expect_cuda(new_values, torch.float32, ndim=2, last=8, what="new_values")
        if new_values.numel() != self.get_num_points():
            raise RuntimeError("New values size does not match number of points")   # src/grid/sparse.cu:53-56
        self._values.copy_(new_values)

    def get_cells(self) -> torch.Tensor:
        """(N, 8) iota, as the reference (src/grid/sparse.cu:61-69)."""
        from .grid import _as_uint32
        return _as_uint32(torch.arange(8 * self.get_num_cells(), device=self.device, dtype=torch.int64).view(-1, 8))

    def filter_cell_indices(self, cell_indices: torch.Tensor, values: torch.Tensor, level: float = 0.0) -> torch.Tensor:
        """Keep the cells whose 8 values straddle ``level`` (src/grid/sparse.cu:150-179)."""
        t = _idx_tensor(cell_indices, "cell_indices")
        _expect_cuda(values, torch.float32, ndim=2, last=8, what="values")
        if values.shape[0] != t.numel():
            raise RuntimeError("values must have one row per cell index")
        n = t.numel()
        if values.data_ptr() % 16:      # the kernel reads each row with two 128-bit loads
            values = values.clone()
        keep = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
        out = torch.empty(max(n, 1), dtype=t.dtype, device=self.device)
        lib = _lib.lib()
        wsbuf = self._ws.get("setops", lib.isoext_setops_workspace_bytes(max(n, 1)), self.device)
        n_out = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(lib.isoext_sparse_crossing(values.data_ptr(), n, float(level), keep.data_ptr(), _stream_ptr()))
            _lib.check(lib.isoext_compact_flagged(t.data_ptr(), t.element_size(), keep.data_ptr(), n, out.data_ptr(), wsbuf.data_ptr(),
                                                  wsbuf.numel(), _stream_ptr(), C.byref(n_out)))
        return out[:n_out.value]

    def _geom(self):
        X, Y, Z = self.shape
        return X, Y, Z, _lib.f3(self.aabb_min), _lib.f3(self.aabb_max)

    # -- population on the GPU (extension) ----------------------------------------------
    def populate_from_dense(self, values, level: float = 0.0, x_chunk: int | None = None) -> "SparseGrid":
        """Make this grid the narrow band of a dense field: every cell whose 8 corner values straddle ``level``
        (case not 0 / 255), with those values -- the result of the reference's population recipe
        (tests/conftest.py:39-61 of the reference: potential ids -> points -> sdf -> filter -> add_cells -> set_values)
        in one pass over the field instead of a Python loop over chunks.

        ``values``: a UniformGrid of the same shape and AABB, an (X, Y, Z) float32 CUDA tensor, or an ``ImplicitGrid``
        (analytic SDF evaluated in the kernels: nothing of size X*Y*Z floats is ever allocated, which is what makes a
        4096^3-equivalent band possible).  ``x_chunk``: planes per pass (default: all); chunks bound the sign-bit
        workspace (X*Y*Z/8 bytes unchunked) for very large grids."""
        from .dc import its_dense_raw
        from .grid import ImplicitGrid, UniformGrid
        prog = None
        if isinstance(values, ImplicitGrid):
            if values.shape != self.shape or values.aabb_min != self.aabb_min or values.aabb_max != self.aabb_max:
                raise RuntimeError("Cannot set values with different shapes")
            vals, prog = None, values.program.device(self.device)
        else:
            vals = values._values if isinstance(values, UniformGrid) else values
            _expect_cuda(vals, torch.float32, ndim=3, what="values")
            if tuple(vals.shape) != self.shape:
                raise RuntimeError("Cannot set values with different shapes")
        lib = _lib.lib()
        X, Y, Z = self.shape
        amin, amax = _lib.f3(self.aabb_min), _lib.f3(self.aabb_max)
        step = X if not x_chunk else max(2, int(x_chunk))
        cells, vals8 = [], []
        with torch.cuda.device(self.device):
            for x0 in range(0, X - 1, step - 1 if step < X else X):
                x1 = min(X, x0 + step)                      # planes [x0, x1): cell layers [x0, x1 - 1)
                sub = vals[x0:x1] if vals is not None else None
                if sub is not None and sub.data_ptr() % 32:      # the volume stream reads 256-bit words
                    sub = sub.clone()
                its, _ = its_dense_raw(sub, (x1 - x0, Y, Z), self.aabb_min, self.aabb_max, level, False, self._ws, x_offset=x0,
                                       x_global=X, sdf_prog=prog)
                n = its.n_cells
                ci = torch.empty(n, dtype=torch.int64, device=self.device)
                v8 = torch.empty((n, 8), dtype=torch.float32, device=self.device)
                _lib.check(lib.isoext_band_from_dense_emit(sub.data_ptr() if sub is not None else None, x1 - x0, Y, Z, x0, X, amin, amax,
                                                           its.entries.data_ptr(), its.n_entries, its.cellslot.data_ptr(), ci.data_ptr(),
                                                           v8.data_ptr(), prog.data_ptr() if prog is not None else None, _stream_ptr()))
                cells.append(ci); vals8.append(v8)
                if x1 >= X:
                    break
        self._cells = torch.cat(cells) if len(cells) > 1 else cells[0]
        self._values = torch.cat(vals8) if len(vals8) > 1 else vals8[0]
        self._int32_api = X * Y * Z <= 2147483647
        return self


def mc_sparse(grid: SparseGrid, level: float, method_id: int, emit_range=None, x_thresholds=(-math.inf, math.inf),
              with_counts: bool = False):
    """Marching cubes over the cell list.  Slabs (dist.SparseSlab): only the cells ``emit_range`` = [begin, end) of the
    list emit triangles, the vertices of all cells are welded, and with ``with_counts`` the result is
    ``(V_all, F, n_lo, n_hi)`` -- all welded vertices plus the ownership counts against ``x_thresholds``."""
    lib = _lib.lib()
    n = grid.get_num_cells()
    empty = (None, None, 0, 0) if with_counts else (None, None)
    if n == 0:
        return empty
    X, Y, Z, amin, amax = grid._geom()
    dev = grid.device
    e0, e1 = (0, n) if emit_range is None else emit_range
    with torch.cuda.device(dev):
        stream = _stream_ptr()
        ws = grid._ws.get("mc_ws", lib.isoext_mc_sparse_workspace_bytes(n), dev)
        counts = (C.c_int64 * 4)()
        _lib.check(lib.isoext_mc_sparse_count(grid._values.data_ptr(), grid._cells.data_ptr(), n, X, Y, Z, amin, amax, float(level),
                                              method_id, e0, e1, ws.data_ptr(), ws.numel(), stream, counts))
        T, Vc = int(counts[0]), int(counts[1])
        if Vc == 0 or (T == 0 and not with_counts):
            return empty
        scratch = grid._ws.get("scratch", lib.isoext_sparse_scratch_bytes(Vc, X, Y), dev)
        V = torch.empty((Vc, 3), dtype=torch.float32, device=dev)
        F = torch.empty((T, 3), dtype=torch.int32, device=dev)
        out = (C.c_int64 * 4)()
        _lib.check(lib.isoext_mc_sparse_emit(grid._values.data_ptr(), grid._cells.data_ptr(), n, X, Y, Z, amin, amax, float(level),
                                             method_id, float(x_thresholds[0]), float(x_thresholds[1]), ws.data_ptr(), ws.numel(),
                                             scratch.data_ptr(), scratch.numel(), Vc, V.data_ptr(), F.data_ptr(), stream, out))
    if with_counts:
        return V[:int(out[0])], F, int(out[1]), int(out[2])
    return V[:int(out[0])], F


def its_sparse(grid: SparseGrid, level: float, compute_normals: bool):
    from .dc import Intersection
    lib = _lib.lib()
    n = grid.get_num_cells()
    X, Y, Z, amin, amax = grid._geom()
    dev = grid.device
    with torch.cuda.device(dev):
        stream = _stream_ptr()
        cinfo = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        cellslot = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        its_off = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        counts = (C.c_int64 * 4)()
        ws = grid._ws.get("mc_ws", lib.isoext_mc_sparse_workspace_bytes(n), dev)
        _lib.check(lib.isoext_its_sparse_count(grid._values.data_ptr(), n, float(level), cinfo.data_ptr(), cellslot.data_ptr(),
                                               its_off.data_ptr(), ws.data_ptr(), ws.numel(), stream, counts))
        n_cells, n_its = int(counts[0]), int(counts[1])
        points = torch.empty((n_its, 3), dtype=torch.float32, device=dev)
        normals = torch.zeros((n_its, 3), dtype=torch.float32, device=dev)
        cell_offsets = torch.zeros(n_cells + 1, dtype=torch.int32, device=dev)
        cell_indices = torch.empty(n_cells, dtype=torch.int64, device=dev)
        _lib.check(lib.isoext_its_sparse_emit(grid._values.data_ptr(), grid._cells.data_ptr(), n, X, Y, Z, amin, amax, float(level),
                                              1 if compute_normals else 0, cinfo.data_ptr(), cellslot.data_ptr(), its_off.data_ptr(),
                                              n_cells, n_its, points.data_ptr(), normals.data_ptr(), cell_offsets.data_ptr(),
                                              cell_indices.data_ptr(), stream))
    return Intersection._make(kind="sparse", shape=grid.shape, aabb_min=grid.aabb_min, aabb_max=grid.aabb_max, level=float(level),
                              cinfo=cinfo, cellslot=cellslot, its_off=its_off, n_sparse=n, n_cells=n_cells, cells=grid._cells,
                              points=points, normals=normals, cell_offsets=cell_offsets, cell_indices=cell_indices,
                              _has_normals=bool(compute_normals))


def dc_sparse_raw(grid: SparseGrid, its, reg: float, svd_tol: float, want_quads: bool = False, dual_v_in=None, emit_range=None,
                  x_thresholds=(-math.inf, math.inf), with_counts: bool = False):
    """(v, f, dual_v, quads) like dc.dc_dense_raw; ``dual_v_in`` replaces the solved dual vertices (parity tests).

    Slabs (``dist.SparseSlab(dc=True)``): only the cells ``emit_range = (begin, end)`` of the list emit quads, the welded
    vertices are classified by x against ``x_thresholds``; with ``with_counts`` the result is
    ``(v_all, f, dual_v, quads, n_lo, n_hi)`` -- ``v_all[n_lo:n_hi]`` are the vertices this slab owns, ``f`` holds ids in
    the slab's extended id space -- and a slab that emits no quad still welds (its vertices may be referenced by a
    neighbour's quads)."""
    lib = _lib.lib()
    n = grid.get_num_cells()
    X, Y, Z, amin, amax = grid._geom()
    dev = grid.device
    dual_v = torch.empty((its.n_cells, 3), dtype=torch.float32, device=dev)
    empty = (None, None, dual_v, None, 0, 0) if with_counts else (None, None, dual_v, None)
    if n == 0 or its.n_cells == 0:
        return empty
    e0, e1 = (0, n) if emit_range is None else (int(emit_range[0]), int(emit_range[1]))
    with torch.cuda.device(dev):
        stream = _stream_ptr()
        ws = grid._ws.get("dc_ws", lib.isoext_dc_sparse_workspace_bytes(n, X), dev)
        counts = (C.c_int64 * 4)()
        _lib.check(lib.isoext_dc_sparse_count(grid._values.data_ptr(), grid._cells.data_ptr(), n, X, Y, Z, amin, amax,
                                              its.cinfo.data_ptr(), its.cellslot.data_ptr(), its.its_off.data_ptr(),
                                              its.points.data_ptr(), its.normals.data_ptr(), float(reg), float(svd_tol), e0, e1,
                                              dual_v.data_ptr(), ws.data_ptr(), ws.numel(), stream, counts))
        Q, Vc = int(counts[0]), int(counts[1])
        if Vc == 0 or (Q == 0 and not with_counts):
            return empty
        if dual_v_in is not None:
            dual_v.copy_(dual_v_in)
        scratch = grid._ws.get("scratch", lib.isoext_sparse_scratch_bytes(Vc, X, Y), dev)
        V = torch.empty((Vc, 3), dtype=torch.float32, device=dev)
        F = torch.empty((2 * Q, 3), dtype=torch.int32, device=dev)
        quads = torch.empty((Q, 4), dtype=torch.int32, device=dev) if want_quads else None
        out = (C.c_int64 * 4)()
        _lib.check(lib.isoext_dc_sparse_emit(grid._cells.data_ptr(), n, X, Y, Z, amin, amax, its.cinfo.data_ptr(),
                                             its.cellslot.data_ptr(), dual_v.data_ptr(), ws.data_ptr(), ws.numel(),
                                             scratch.data_ptr(), scratch.numel(), Vc, float(x_thresholds[0]), float(x_thresholds[1]),
                                             V.data_ptr(), F.data_ptr() if Q else None, quads.data_ptr() if want_quads and Q else None,
                                             stream, out))
    if with_counts:
        return V[:int(out[0])], F, dual_v, quads, int(out[1]), int(out[2])
    return V[:int(out[0])], F, dual_v, quads


def dc_sparse(grid: SparseGrid, level, intersection, reg, svd_tol):
    lib = _lib.lib()
    its = intersection._copy() if intersection is not None else its_sparse(grid, level, True)
    if its.kind != "sparse" or its.n_sparse != grid.get_num_cells():
        raise RuntimeError("intersection does not belong to this grid")
    if not its.has_normals():
        X, Y, Z, amin, amax = grid._geom()
        with torch.cuda.device(grid.device):
            _lib.check(lib.isoext_its_sparse_emit(grid._values.data_ptr(), grid._cells.data_ptr(), its.n_sparse, X, Y, Z, amin, amax,
                                                  its.level, 2, its.cinfo.data_ptr(), its.cellslot.data_ptr(), its.its_off.data_ptr(),
                                                  its.n_cells, its.points.shape[0], its.points.data_ptr(), its.normals.data_ptr(),
                                                  None, None, _stream_ptr()))
        its._has_normals = True
    v, f, _, _ = dc_sparse_raw(grid, its, reg, svd_tol)
    return v, f
