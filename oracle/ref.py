"""ctypes binding of oracle/_ref/libisoext_ref.so -- the UNMODIFIED reference CUDA sources behind
the C shim oracle/ref_shim.cu (TEST INFRASTRUCTURE; needs a GPU).  Mirrors the reference's Python
API (src/isoext_ext.cu:93-384) closely enough for parity tests and the ``--impl reference`` bench
arm.  torch is used only to hold device memory.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libisoext_ref.so"
_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference is mounted")
        h = C.CDLL(str(LIB_PATH))
        h.ref_last_error.restype = C.c_char_p
        for n in ("ref_grid_num_cells", "ref_grid_num_points"):
            getattr(h, n).restype = C.c_uint
        for n in ("ref_its_num_points", "ref_its_num_cells"):
            getattr(h, n).restype = C.c_size_t
        for n in ("ref_its_points", "ref_its_normals", "ref_its_edges", "ref_its_cell_indices", "ref_its_cell_offsets",
                  "ref_its_is_out"):
            getattr(h, n).restype = C.c_void_p
        _lib = h
    return _lib


def _ck(rc):
    if rc != 0:
        raise RuntimeError(lib().ref_last_error().decode())


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _adopt(ptr, shape, dtype):
    """Copy a reference-owned device buffer into a torch tensor and free the original."""
    n = 1
    for s in shape:
        n *= s
    t = torch.empty(shape, dtype=dtype, device="cuda")
    if n:
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr),
                                                          C.c_size_t(n * t.element_size()), C.c_int(3))
    lib().ref_free(C.c_void_p(ptr))
    return t


def _borrow(ptr, shape, dtype):
    n = 1
    for s in shape:
        n *= s
    t = torch.empty(shape, dtype=dtype, device="cuda")
    if n:
        C.cdll.LoadLibrary("libcudart.so.12").cudaMemcpy(C.c_void_p(t.data_ptr()), C.c_void_p(ptr),
                                                          C.c_size_t(n * t.element_size()), C.c_int(3))
    return t


class _GridBase:
    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_grid_delete(self.h)
            self.h = None

    def get_num_cells(self):
        return int(lib().ref_grid_num_cells(self.h))

    def get_num_points(self):
        return int(lib().ref_grid_num_points(self.h))


class UniformGrid(_GridBase):
    def __init__(self, shape, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1), default_value=3.4028234663852886e38):
        self.shape = tuple(int(s) for s in shape)
        self.h = C.c_void_p()
        _ck(lib().ref_uniform_new(*[C.c_uint(s) for s in self.shape], _f3(aabb_min), _f3(aabb_max),
                                  C.c_float(default_value), C.byref(self.h)))

    def set_values(self, t: torch.Tensor):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        _ck(lib().ref_uniform_set_values(self.h, C.c_void_p(t.data_ptr()), *[C.c_size_t(s) for s in t.shape]))

    def get_points(self):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().ref_grid_points(self.h, C.byref(p), C.byref(n)))
        return _adopt(p.value, (*self.shape, 3), torch.float32)


    def get_cells(self):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().ref_grid_cells(self.h, C.byref(p), C.byref(n)))
        X, Y, Z = self.shape
        return _adopt(p.value, (X - 1, Y - 1, Z - 1, 8), torch.int32)      # uint32 bits


class SparseGrid(_GridBase):
    def __init__(self, shape, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1), default_value=3.4028234663852886e38):
        self.shape = tuple(int(s) for s in shape)
        self.h = C.c_void_p()
        _ck(lib().ref_sparse_new(*[C.c_uint(s) for s in self.shape], _f3(aabb_min), _f3(aabb_max),
                                 C.c_float(default_value), C.byref(self.h)))

    def add_cells(self, idx: torch.Tensor):
        _ck(lib().ref_sparse_add_cells(self.h, C.c_void_p(idx.data_ptr()), C.c_size_t(idx.numel())))

    def remove_cells(self, idx: torch.Tensor):
        _ck(lib().ref_sparse_remove_cells(self.h, C.c_void_p(idx.data_ptr()), C.c_size_t(idx.numel())))

    def set_values(self, t: torch.Tensor):
        _ck(lib().ref_sparse_set_values(self.h, C.c_void_p(t.data_ptr()), C.c_size_t(t.shape[0])))

    def get_cell_indices(self):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().ref_sparse_cell_indices(self.h, C.byref(p), C.byref(n)))
        return _adopt(p.value, (n.value,), torch.int32)

    def get_points(self):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().ref_grid_points(self.h, C.byref(p), C.byref(n)))
        return _adopt(p.value, (n.value // 8, 8, 3), torch.float32)

    def get_points_by_cell_indices(self, idx: torch.Tensor):
        p = C.c_void_p()
        _ck(lib().ref_sparse_points_by_cell_indices(self.h, C.c_void_p(idx.data_ptr()), C.c_size_t(idx.numel()), C.byref(p)))
        return _adopt(p.value, (idx.numel(), 8, 3), torch.float32)

    def filter_cell_indices(self, idx: torch.Tensor, values: torch.Tensor, level=0.0):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().ref_sparse_filter_cell_indices(self.h, C.c_void_p(idx.data_ptr()), C.c_void_p(values.data_ptr()),
                                                 C.c_size_t(idx.numel()), C.c_float(level), C.byref(p), C.byref(n)))
        return _adopt(p.value, (n.value,), torch.int32)


def marching_cubes(grid, level=0.0, method="nagae"):
    v, f, nv, nf = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    _ck(lib().ref_marching_cubes(grid.h, C.c_float(level), method.encode(), C.byref(v), C.byref(nv), C.byref(f), C.byref(nf)))
    if nv.value == 0:
        return None, None
    return _adopt(v.value, (nv.value, 3), torch.float32), _adopt(f.value, (nf.value, 3), torch.int32)


def marching_cubes_timed_raw(grid, level, method):
    """One reference call with results freed, no copies (for the bench arm)."""
    v, f, nv, nf = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    _ck(lib().ref_marching_cubes(grid.h, C.c_float(level), method.encode(), C.byref(v), C.byref(nv), C.byref(f), C.byref(nf)))
    return v, f, nv.value, nf.value


class Intersection:
    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_its_delete(self.h)
            self.h = None

    def num_points(self):
        return int(lib().ref_its_num_points(self.h))

    def num_cells(self):
        return int(lib().ref_its_num_cells(self.h))

    def has_normals(self):
        return bool(lib().ref_its_has_normals(self.h))

    def get_points(self):
        return _borrow(lib().ref_its_points(self.h), (self.num_points(), 3), torch.float32)

    def get_normals(self):
        return _borrow(lib().ref_its_normals(self.h), (self.num_points(), 3), torch.float32)

    def get_edges(self):
        return _borrow(lib().ref_its_edges(self.h), (self.num_points(), 2), torch.int32)

    def get_is_out(self):
        return _borrow(lib().ref_its_is_out(self.h), (self.num_points(),), torch.bool)

    def get_cell_indices(self):
        return _borrow(lib().ref_its_cell_indices(self.h), (self.num_cells(),), torch.int32)

    def get_cell_offsets(self):
        return _borrow(lib().ref_its_cell_offsets(self.h), (self.num_cells() + 1,), torch.int32)

    def set_normals(self, n: torch.Tensor):
        _ck(lib().ref_its_set_normals(self.h, C.c_void_p(n.data_ptr()), C.c_size_t(n.shape[0])))


def get_intersection(grid, level=0.0, compute_normals=False):
    h = C.c_void_p()
    _ck(lib().ref_get_intersection(grid.h, C.c_float(level), C.c_int(int(compute_normals)), C.byref(h)))
    return Intersection(h)


def dual_contouring(grid, level=0.0, intersection=None, reg=1e-2, svd_tol=1e-6):
    v, f, nv, nf = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
    ih = intersection.h if intersection is not None else None
    _ck(lib().ref_dual_contouring(grid.h, C.c_float(level), ih, C.c_float(reg), C.c_float(svd_tol),
                                  C.byref(v), C.byref(nv), C.byref(f), C.byref(nf)))
    if nv.value == 0:
        return None, None
    return _adopt(v.value, (nv.value, 3), torch.float32), _adopt(f.value, (nf.value, 3), torch.int32)


def dc_dual_vertices(grid, intersection, reg=1e-2, svd_tol=1e-6):
    """Per ACTIVE CELL (order of ``intersection.get_cell_indices()``), from the reference's own compiled code
    (oracle/ref_shim_dc.cu -> src/dc.cu:166-182): dict(dual_v (S,3) clipped, raw (S,3) before the clip, ATA (S,3,3),
    ATb (S,3) float32 QEF, info (S,) cuSOLVER gesvdj status).  ``intersection`` must carry normals."""
    h = lib()
    h.ref_dc_last_error.restype = C.c_char_p
    p = [C.c_void_p() for _ in range(5)]
    n = C.c_size_t()
    rc = h.ref_dc_dual_vertices(grid.h, intersection.h, C.c_float(reg), C.c_float(svd_tol),
                                *[C.byref(x) for x in p], C.byref(n))
    if rc != 0:
        raise RuntimeError(h.ref_dc_last_error().decode())
    S = n.value
    return dict(dual_v=_adopt(p[0].value, (S, 3), torch.float32), raw=_adopt(p[1].value, (S, 3), torch.float32),
                ATA=_adopt(p[2].value, (S, 3, 3), torch.float32), ATb=_adopt(p[3].value, (S, 3), torch.float32),
                info=_adopt(p[4].value, (S,), torch.int32))


def free(ptr):
    lib().ref_free(ptr)
