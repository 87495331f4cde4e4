"""ctypes/numpy binding of oracle/oracle_c.c (CPU restatement of the reference; test infrastructure).

Every function cites the reference file:line it restates in oracle_c.c.  Inputs/outputs are numpy
arrays; nothing here touches CUDA or the product package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "liboracle_c.so"
_lib = None

METHODS = {"nagae": 0, "lorensen": 1}


class _Mesh(C.Structure):
    _fields_ = [("v", C.POINTER(C.c_float)), ("f", C.POINTER(C.c_int32)),
                ("nv", C.c_int64), ("nf", C.c_int64), ("n_active", C.c_int64)]


class _Its(C.Structure):
    _fields_ = [("points", C.POINTER(C.c_float)), ("normals", C.POINTER(C.c_float)),
                ("edges", C.POINTER(C.c_uint64)), ("is_out", C.POINTER(C.c_uint8)),
                ("cell_indices", C.POINTER(C.c_int64)), ("cell_coords", C.POINTER(C.c_int64)),
                ("cell_offsets", C.POINTER(C.c_int64)), ("n_points", C.c_int64), ("n_cells", C.c_int64)]


class _DC(C.Structure):
    _fields_ = [("mesh", _Mesh), ("dual_v", C.POINTER(C.c_double)), ("quads", C.POINTER(C.c_int64)),
                ("quad_edge", C.POINTER(C.c_uint64)), ("n_quads", C.c_int64), ("n_skipped", C.c_int64)]


def build_c(force: bool = False) -> Path:
    """Compile oracle_c.c -> liboracle_c.so (gcc, no FMA contraction)."""
    src = _HERE / "oracle_c.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(
            src.stat().st_mtime, (_HERE / "oracle_luts.h").stat().st_mtime):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
                        "-o", str(_LIB_PATH), str(src), "-lm"], check=True, cwd=_HERE)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_c()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.orc_case_histogram.restype = C.c_int64
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _aabb(aabb_min, aabb_max):
    return _f32(aabb_min), _f32(aabb_max)


def _take_mesh(m: _Mesh):
    if m.nf == 0:
        v, f = np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32)
    else:
        v = np.ctypeslib.as_array(m.v, shape=(m.nv, 3)).copy()
        f = np.ctypeslib.as_array(m.f, shape=(m.nf, 3)).copy()
    return v, f, int(m.n_active)


def points_dense(shape, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)) -> np.ndarray:
    """(X,Y,Z,3) grid point positions -- src/grid/uniform.cu:22-30, include/utils.cuh:62-80."""
    X, Y, Z = (int(s) for s in shape)
    lo, hi = _aabb(aabb_min, aabb_max)
    out = np.empty((X, Y, Z, 3), np.float32)
    _load().orc_points_dense(C.c_int64(X), C.c_int64(Y), C.c_int64(Z), _fp(lo), _fp(hi), _fp(out))
    return out


def points_cells(shape, cell_idx, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)) -> np.ndarray:
    """(K,8,3) corner positions of the given cells -- src/grid/sparse.cu:21-36."""
    X, Y, Z = (int(s) for s in shape)
    lo, hi = _aabb(aabb_min, aabb_max)
    idx = np.ascontiguousarray(cell_idx, dtype=np.int64)
    out = np.empty((len(idx), 8, 3), np.float32)
    _load().orc_points_cells(C.c_int64(X), C.c_int64(Y), C.c_int64(Z), _fp(lo), _fp(hi),
                             idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(idx)), _fp(out))
    return out


def cells_dense(shape) -> np.ndarray:
    """(X-1,Y-1,Z-1,8) corner point ids of every cell -- UniformGrid::get_cells, src/grid/uniform.cu:42-51 with
    idx_to_cell_op, include/utils.cuh:32-60 (integer index arithmetic, restated in numpy; int64, no overflow)."""
    X, Y, Z = (int(s) for s in shape)
    i = np.arange((X - 1) * (Y - 1) * (Z - 1), dtype=np.int64)
    z = i % (Z - 1); i = i // (Z - 1)
    y = i % (Y - 1)
    x = i // (Y - 1)
    yz = Y * Z
    c = np.empty((len(z), 8), np.int64)
    c[:, 0] = x * yz + y * Z + z
    c[:, 1] = c[:, 0] + 1
    c[:, 2] = c[:, 0] + Z
    c[:, 3] = c[:, 1] + Z
    c[:, 4] = c[:, 0] + yz
    c[:, 5] = c[:, 1] + yz
    c[:, 6] = c[:, 2] + yz
    c[:, 7] = c[:, 3] + yz
    return c.reshape(X - 1, Y - 1, Z - 1, 8)


def mc_dense(values, level=0.0, method="nagae", aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1), x_range=None):
    """Marching cubes on a dense (X,Y,Z) field -- src/mc/mc.cu:17-68.  Returns (v, f, n_active)."""
    vals = _f32(values)
    X, Y, Z = vals.shape
    lo, hi = _aabb(aabb_min, aabb_max)
    x0, x1 = (0, X - 1) if x_range is None else x_range
    m = _Mesh()
    _load().orc_mc_dense(_fp(vals), C.c_int64(X), C.c_int64(Y), C.c_int64(Z), _fp(lo), _fp(hi),
                         C.c_float(level), C.c_int(METHODS[method]), C.c_int64(x0), C.c_int64(x1), C.byref(m))
    out = _take_mesh(m)
    _lib.orc_mesh_free(C.byref(m))
    return out


def mc_sparse(values8, cell_idx, shape, level=0.0, method="nagae", aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)):
    """Marching cubes on a sparse grid ((N,8) values + N sorted cell indices) -- src/mc/mc.cu:17-68."""
    vals = _f32(values8).reshape(-1, 8)
    idx = np.ascontiguousarray(cell_idx, dtype=np.int64)
    X, Y, Z = (int(s) for s in shape)
    lo, hi = _aabb(aabb_min, aabb_max)
    m = _Mesh()
    _load().orc_mc_sparse(_fp(vals), idx.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(idx)),
                          C.c_int64(X), C.c_int64(Y), C.c_int64(Z), _fp(lo), _fp(hi),
                          C.c_float(level), C.c_int(METHODS[method]), C.byref(m))
    out = _take_mesh(m)
    _lib.orc_mesh_free(C.byref(m))
    return out


def case_histogram(values, level=0.0):
    """(hist[256], n_active) of the MC case index -- include/utils.cuh:82-102."""
    vals = _f32(values)
    X, Y, Z = vals.shape
    hist = np.zeros(256, np.int64)
    n = _load().orc_case_histogram(_fp(vals), C.c_int64(X), C.c_int64(Y), C.c_int64(Z), C.c_float(level),
                                   hist.ctypes.data_as(C.POINTER(C.c_int64)))
    return hist, int(n)


class Intersection:
    """Host copy of the reference's Intersection struct (include/its.cuh:6-26)."""

    def __init__(self, raw: _Its, has_normals: bool):
        S, I = int(raw.n_cells), int(raw.n_points)
        arr = np.ctypeslib.as_array
        self.points = arr(raw.points, shape=(max(I, 1), 3))[:I].copy()
        self.normals = arr(raw.normals, shape=(max(I, 1), 3))[:I].copy()
        self.edges = arr(raw.edges, shape=(max(I, 1), 2))[:I].copy()
        self.is_out = arr(raw.is_out, shape=(max(I, 1),))[:I].copy().astype(bool)
        self.cell_indices = arr(raw.cell_indices, shape=(max(S, 1),))[:S].copy()
        self.cell_coords = arr(raw.cell_coords, shape=(max(S, 1), 3))[:S].copy()
        self.cell_offsets = arr(raw.cell_offsets, shape=(S + 1,)).copy()
        self.has_normals = has_normals


def get_intersection(values, shape=None, cell_idx=None, level=0.0, compute_normals=False,
                     aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)) -> Intersection:
    """src/its.cu:93-159 (+ :170-284 for the normals).  Dense if cell_idx is None."""
    vals = _f32(values)
    lo, hi = _aabb(aabb_min, aabb_max)
    if cell_idx is None:
        X, Y, Z = vals.shape
        idx_p, n = None, 0
    else:
        X, Y, Z = (int(s) for s in shape)
        idx = np.ascontiguousarray(cell_idx, dtype=np.int64)
        idx_p, n = idx.ctypes.data_as(C.POINTER(C.c_int64)), len(idx)
    raw = _Its()
    _load().orc_get_intersection(_fp(vals), idx_p, C.c_int64(n), C.c_int64(X), C.c_int64(Y), C.c_int64(Z),
                                 _fp(lo), _fp(hi), C.c_float(level), C.c_int(int(compute_normals)), C.byref(raw))
    out = Intersection(raw, bool(compute_normals))
    _lib.orc_its_free(C.byref(raw))
    return out


def dual_contouring(its: Intersection, shape, reg=1e-2, svd_tol=1e-6, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1)):
    """src/dc.cu:161-218 with a float64 solve.  Returns dict(v, f, dual_v, quads, quad_edge, n_skipped)."""
    X, Y, Z = (int(s) for s in shape)
    lo, hi = _aabb(aabb_min, aabb_max)
    raw = _Its()
    keep = []

    def put(name, a, ctype, dtype):
        a = np.ascontiguousarray(a, dtype=dtype)
        keep.append(a)
        setattr(raw, name, a.ctypes.data_as(C.POINTER(ctype)))

    put("points", its.points, C.c_float, np.float32)
    put("normals", its.normals, C.c_float, np.float32)
    put("edges", its.edges, C.c_uint64, np.uint64)
    put("is_out", its.is_out.astype(np.uint8), C.c_uint8, np.uint8)
    put("cell_indices", its.cell_indices, C.c_int64, np.int64)
    put("cell_coords", its.cell_coords, C.c_int64, np.int64)
    put("cell_offsets", its.cell_offsets, C.c_int64, np.int64)
    raw.n_points, raw.n_cells = len(its.points), len(its.cell_indices)
    d = _DC()
    _load().orc_dual_contouring(C.byref(raw), C.c_int64(X), C.c_int64(Y), C.c_int64(Z), _fp(lo), _fp(hi),
                                C.c_float(reg), C.c_float(svd_tol), C.byref(d))
    v, f, _ = _take_mesh(d.mesh)
    S, Q = raw.n_cells, int(d.n_quads)
    arr = np.ctypeslib.as_array
    out = dict(v=v, f=f,
               dual_v=arr(d.dual_v, shape=(max(S, 1), 3))[:S].copy(),
               quads=arr(d.quads, shape=(max(Q, 1), 4))[:Q].copy(),
               quad_edge=arr(d.quad_edge, shape=(max(Q, 1), 2))[:Q].copy(),
               n_skipped=int(d.n_skipped))
    _lib.orc_dc_free(C.byref(d))
    return out
