"""Grid classes of the public API -- host-side mirror of the reference's ``UniformGrid`` /
``SparseGrid`` bindings (src/isoext_ext.cu:111-303, src/grid/uniform.cu, src/grid/sparse.cu).

Storage is torch-owned CUDA memory; kernels are reached through the C-ABI (``_lib``).  Unlike the
reference nothing is materialised per call: no ``points`` (12 B/pt) and no ``cells`` (32 B/cell).
"""
from __future__ import annotations

import torch

from . import _lib

FLT_MAX = 3.4028234663852886e38
INT_MAX = 2147483647


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _expect_cuda(t, dtype, ndim=None, last=None, what="tensor"):
    """nanobind rejects wrong dtype/device/rank at argument matching with TypeError
    (src/isoext_ext.cu:21-29); mirror that."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise TypeError(f"{what}: expected a CUDA tensor")
    if t.dtype != dtype:
        raise TypeError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise TypeError(f"{what}: expected {ndim} dimensions, got {t.dim()}")
    if last is not None and t.shape[-1] != last:
        raise TypeError(f"{what}: expected last dimension {last}, got {t.shape[-1]}")
    if not t.is_contiguous():
        raise TypeError(f"{what}: expected a contiguous tensor")
    return t


class Grid:
    """Common base (include/grid/grid.cuh:10-35).  Not constructible, like the reference's."""

    def get_num_cells(self) -> int:
        raise NotImplementedError

    def get_num_points(self) -> int:
        raise NotImplementedError


class _Workspace:
    """Grow-only device buffers reused across calls on the same grid (avoids the reference's
    >= 12 cudaMalloc/cudaFree pairs per extraction)."""

    def __init__(self):
        self.buf = {}

    def get(self, name: str, nbytes: int, device) -> torch.Tensor:
        t = self.buf.get(name)
        if t is None or t.numel() < nbytes or t.device != device:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self.buf[name] = t
        return t


def _as_uint32(ids: torch.Tensor) -> torch.Tensor:
    """The reference hands point ids out as uint32 (NDArray<uint>, src/isoext_ext.cu:90); keep that dtype while the ids
    fit (and the installed torch can convert on this device), int64 beyond -- the reference overflows there."""
    if ids.numel() and int(ids.max()) > 0xFFFFFFFF:
        return ids
    try:
        return ids.to(torch.uint32)
    except (RuntimeError, TypeError):
        return ids


class UniformGrid(Grid):
    """Dense scalar field on a regular lattice.  ``shape`` is the number of POINTS per axis
    (src/grid/uniform.cu:8-20; 64^3 -> 250,047 cells / 262,144 points).

    Extension over the reference: grids with more than INT_MAX points are accepted (the reference
    throws, src/grid/uniform.cu:13-17).  Limits of one extraction call: Z <= 65,535 points along the last axis,
    X*Y < 2^31 rows, X*Y*ceil(Z/32) < 2^31, fewer than 2^29 surface entries and vertex candidates; the field
    storage must be 32-byte aligned (torch allocations are).  Larger grids: ``isoext_b200.dist.SlabGrid``.
    """

    def __init__(self, shape, aabb_min=(-1.0, -1.0, -1.0), aabb_max=(1.0, 1.0, 1.0), default_value=FLT_MAX,
                 device=None):
        shape = [int(s) for s in shape]
        if len(shape) != 3 or len(aabb_min) != 3 or len(aabb_max) != 3:
            raise TypeError("shape, aabb_min and aabb_max must have three elements")
        if min(shape) < 1:
            raise RuntimeError("Grid shape must be positive")
        self.shape = tuple(shape)
        self.aabb_min = tuple(float(v) for v in aabb_min)
        self.aabb_max = tuple(float(v) for v in aabb_max)
        self.default_value = float(default_value)
        _lib.lib()  # fail loudly before allocating anything if the CUDA library is missing
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._values = torch.full(self.shape, self.default_value, dtype=torch.float32, device=self.device)
        self._ws = _Workspace()
        self._cap_hint = 0
        self._hints = {}   # sizes of the previous extraction (single-call fast path of marching_cubes)

    # -- sizes ------------------------------------------------------------------------------
    def get_num_cells(self) -> int:
        X, Y, Z = self.shape
        return (X - 1) * (Y - 1) * (Z - 1)

    def get_num_points(self) -> int:
        X, Y, Z = self.shape
        return X * Y * Z

    # -- data -------------------------------------------------------------------------------
    def get_points(self) -> torch.Tensor:
        """(X, Y, Z, 3) float32 positions, bit-identical to the reference's get_vtx_pos_op
        (include/utils.cuh:62-80)."""
        X, Y, Z = self.shape
        out = torch.empty((X, Y, Z, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().isoext_grid_points_dense(X, Y, Z, 0, X, _lib.f3(self.aabb_min), _lib.f3(self.aabb_max),
                                                           out.data_ptr(), _stream_ptr()))
        return out

    def get_values(self) -> torch.Tensor:
        """A copy of the (X, Y, Z) field (the reference deep-copies too, src/grid/uniform.cu:32-35)."""
        return self._values.clone()

    def set_values(self, new_values: torch.Tensor) -> None:
        _expect_cuda(new_values, torch.float32, ndim=3, what="new_values")
        if tuple(new_values.shape) != self.shape:
            raise RuntimeError("Cannot set values with different shapes")  # include/ndarray.cuh:79-84
        self._values.copy_(new_values)

    def values_view(self) -> torch.Tensor:
        """Zero-copy view of the grid's own storage (extension; lets callers fill it in place)."""
        return self._values

    def get_cells(self) -> torch.Tensor:
        """(X-1, Y-1, Z-1, 8) point ids per cell in Morton corner order (src/grid/uniform.cu:42-51,
        include/utils.cuh:32-60): uint32 like the reference while the ids fit, int64 beyond 2^32 points (where the
        reference overflows).  Only for inspection: no kernel of this package consumes it."""
        X, Y, Z = self.shape
        wide = X * Y * Z > 0xFFFFFFFF + 1
        try:
            out = torch.empty((X - 1, Y - 1, Z - 1, 8), dtype=torch.int64 if wide else torch.uint32, device=self.device)
        except (RuntimeError, TypeError):          # a torch build without uint32 CUDA tensors
            wide, out = True, torch.empty((X - 1, Y - 1, Z - 1, 8), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().isoext_grid_cells_dense(X, Y, Z, int(wide), out.data_ptr(), _stream_ptr()))
        return out


class ImplicitGrid(Grid):
    """A uniform lattice whose values are an analytic SDF evaluated INSIDE the extraction kernels (extension,
    SURVEY.md 8f-1): ``marching_cubes`` / ``get_intersection`` / ``dual_contouring`` accept it like a ``UniformGrid``,
    but no (X, Y, Z) field is ever allocated, written or read -- where a kernel would load a value it evaluates the
    compiled program (csrc/sdfprog.cuh) at the position ``get_points()`` reports for that point, and the volume pass
    skips every 32-point word that the 1-Lipschitz bound proves to be far from the surface.

    ``sdf``: a tree of the built-in classes of ``isoext_b200.sdf`` (``compile_sdf``).  ``materialize()`` returns the
    equivalent ``UniformGrid`` (field written by the same device function): extracting that grid gives the same mesh
    bit for bit -- the parity contract of the fused path."""

    def __init__(self, shape, sdf, aabb_min=(-1.0, -1.0, -1.0), aabb_max=(1.0, 1.0, 1.0), device=None):
        from .sdf import SdfProgram, compile_sdf
        shape = [int(s) for s in shape]
        if len(shape) != 3 or len(aabb_min) != 3 or len(aabb_max) != 3:
            raise TypeError("shape, aabb_min and aabb_max must have three elements")
        if min(shape) < 1:
            raise RuntimeError("Grid shape must be positive")
        self.shape = tuple(shape)
        self.aabb_min = tuple(float(v) for v in aabb_min)
        self.aabb_max = tuple(float(v) for v in aabb_max)
        _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.program = sdf if isinstance(sdf, SdfProgram) else compile_sdf(sdf)
        self._prog = self.program.device(self.device)
        self._values = None
        self._ws = _Workspace()
        self._cap_hint = 0
        self._hints = {}

    def get_num_cells(self) -> int:
        X, Y, Z = self.shape
        return (X - 1) * (Y - 1) * (Z - 1)

    def get_num_points(self) -> int:
        X, Y, Z = self.shape
        return X * Y * Z

    get_points = UniformGrid.get_points

    def eval_slab(self, x0: int, x1: int) -> torch.Tensor:
        """(x1 - x0, Y, Z) float32 values of point planes [x0, x1) (kernel evaluation of the program)."""
        X, Y, Z = self.shape
        out = torch.empty((x1 - x0, Y, Z), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().isoext_sdf_eval_dense(self._prog.data_ptr(), x1 - x0, Y, Z, x0, X, _lib.f3(self.aabb_min),
                                                        _lib.f3(self.aabb_max), out.data_ptr(), _stream_ptr()))
        return out

    def get_values(self) -> torch.Tensor:
        return self.eval_slab(0, self.shape[0])

    def materialize(self) -> UniformGrid:
        g = UniformGrid(list(self.shape), self.aabb_min, self.aabb_max, device=self.device)
        g._values = self.get_values()
        return g
