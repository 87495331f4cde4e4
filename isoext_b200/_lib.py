"""ctypes loader of the C-ABI library ``libisoext_b200.so`` (declared in include/isoext_b200.h).

There is no CPU fallback and no alternative backend: if the CUDA library is missing or a call
fails, this module raises.  The test-only checker package is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libisoext_b200.so"

E_CAPACITY = -4
E_METHOD = -5

_i64, _f32, _sz, _vp, _int = C.c_int64, C.c_float, C.c_size_t, C.c_void_p, C.c_int
_f3 = C.POINTER(C.c_float)
_pi64 = C.POINTER(C.c_int64)

# name -> (restype, argtypes); mirrors include/isoext_b200.h one to one
SIGNATURES = {
    "isoext_last_error": (C.c_char_p, []),
    "isoext_build_info": (C.c_char_p, []),
    "isoext_abi_version": (_int, []),
    "isoext_axis_position": (_f32, [_i64, _i64, _f32, _f32]),
    "isoext_debug_set_signbits_variant": (_int, [_int]),
    "isoext_debug_set_tuning": (_int, [_int, _int]),
    "isoext_debug_detail_enable": (_int, [_int]),
    "isoext_debug_detail_report": (_int, [C.c_char_p, _int]),
    "isoext_debug_detail_timeline": (_int, [C.c_char_p, _int]),
    "isoext_profile_begin": (_int, []),
    "isoext_profile_end": (_int, [C.POINTER(C.c_double), _pi64, _pi64]),
    "isoext_grid_points_dense": (_int, [_i64, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _vp]),
    "isoext_mc_dense_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "isoext_mc_dense_scratch_bytes": (_sz, [_i64]),
    "isoext_mc_dense_count": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _i64, _i64,
                                     _vp, _sz, _i64, _vp, _vp, _pi64]),
    "isoext_mc_dense_emit": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _i64, _i64,
                                    _vp, _sz, _i64, _vp, _sz, _i64, _i64, _f32, _f32, _vp, _vp, _vp, _vp, _pi64]),
    "isoext_its_dense_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "isoext_its_dense_count": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _vp, _sz, _i64, _vp, _vp, _vp, _vp,
                                      _vp, _vp, _pi64]),
    "isoext_its_dense_emit": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _vp, _i64, _vp, _vp, _i64, _i64,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "isoext_its_dense_normals": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "isoext_sdf_program_bytes": (_sz, []),
    "isoext_sdf_eval_dense": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _vp]),
    "isoext_band_from_dense_emit": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "isoext_dc_dense_workspace_bytes": (_sz, [_i64, _i64]),
    "isoext_dc_dense_scratch_bytes": (_sz, [_i64]),
    "isoext_dc_dense_count": (_int, [_i64, _i64, _i64, _i64, _i64, _f3, _f3, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _f32,
                                     _f32, _vp, _vp, _sz, _vp, _pi64]),
    "isoext_dc_dense_emit": (_int, [_i64, _i64, _i64, _i64, _i64, _f3, _f3, _i64, _i64, _f32, _f32, _vp, _i64, _vp, _vp, _vp, _i64, _vp,
                                    _vp, _sz, _vp, _sz, _i64, _vp, _vp, _vp, _vp, _pi64]),
    "isoext_sparse_points": (_int, [_i64, _i64, _i64, _f3, _f3, _vp, _i64, _vp, _vp]),
    "isoext_sparse_crossing": (_int, [_vp, _i64, _f32, _vp, _vp]),
    "isoext_setops_workspace_bytes": (_sz, [_i64]),
    "isoext_ids_sort_unique": (_int, [_vp, _i64, _vp, _vp, _sz, _vp, _pi64]),
    "isoext_ids_difference": (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _sz, _vp, _pi64]),
    "isoext_compact_flagged": (_int, [_vp, _int, _vp, _i64, _vp, _vp, _sz, _vp, _pi64]),
    "isoext_grid_cells_dense": (_int, [_i64, _i64, _i64, _int, _vp, _vp]),
    "isoext_vertex_layer_histogram": (_int, [_vp, _i64, _f32, _f32, _i64, _vp, _vp]),
    "isoext_mc_sparse_workspace_bytes": (_sz, [_i64]),
    "isoext_sparse_scratch_bytes": (_sz, [_i64, _i64, _i64]),
    "isoext_mc_sparse_count": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _i64, _i64, _vp, _sz, _vp, _pi64]),
    "isoext_mc_sparse_emit": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _f32, _f32, _vp, _sz, _vp, _sz, _i64,
                                     _vp, _vp, _vp, _pi64]),
    "isoext_its_sparse_count": (_int, [_vp, _i64, _f32, _vp, _vp, _vp, _vp, _sz, _vp, _pi64]),
    "isoext_its_sparse_emit": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _vp, _vp, _vp, _i64, _i64, _vp, _vp,
                                      _vp, _vp, _vp]),
    "isoext_dc_sparse_workspace_bytes": (_sz, [_i64, _i64]),
    "isoext_dc_sparse_count": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _i64, _i64, _vp,
                                      _vp, _sz, _vp, _pi64]),
    "isoext_dc_sparse_emit": (_int, [_vp, _i64, _i64, _i64, _i64, _f3, _f3, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _i64, _f32, _f32, _vp,
                                     _vp, _vp, _vp, _pi64]),
    "isoext_mc_dense_run": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _f3, _f3, _f32, _int, _i64, _i64,
                                   _vp, _sz, _i64, _vp, _sz, _i64, _i64, _i64, _int, _f32, _f32, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _pi64]),
    "isoext_relabel_faces": (_int, [_vp, _i64, _i64, _i64, _i64, _i64, _vp]),
    "isoext_gaussian_smooth_separable": (_int, [_vp, _i64, _i64, _i64, _f3, _int, _vp, _vp, _vp]),
    "isoext_write_obj": (_int, [C.c_char_p, _vp, _i64, _vp, _i64]),
    "isoext_write_ply": (_int, [C.c_char_p, _vp, _i64, _vp, _i64]),
    "isoext_peer_sync_words": (_int, []),
    "isoext_peer_alloc": (_int, [_sz, C.POINTER(_vp), C.POINTER(C.c_ubyte)]),
    "isoext_peer_free": (_int, [_vp]),
    "isoext_peer_open": (_int, [C.POINTER(C.c_ubyte), C.POINTER(_vp)]),
    "isoext_peer_close": (_int, [_vp]),
    "isoext_peer_publish": (_int, [_vp, C.c_uint64, _vp]),
    "isoext_peer_wait": (_int, [_vp, _vp, C.c_uint64, _vp, _vp]),
    "isoext_peer_halo_pull": (_int, [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, C.c_uint64, _vp, _vp]),
    "isoext_peer_publish_counts": (_int, [_vp, C.c_uint64, _i64, _i64, _vp]),
    "isoext_relabel_faces_peer": (_int, [_vp, _i64, _i64, _i64, _i64, C.POINTER(_vp), _int, C.c_uint64, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """The loaded library; raises ImportError with build instructions if it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: the sm_100a CUDA library has not been built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C isoext_b200/csrc`). "
                "isoext_b200 has no CPU fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
        # development knobs (isoext_debug_set_tuning), e.g. ISOEXT_B200_TUNING="3=101,6=1" for A/B runs of the tools
        for item in filter(None, os.environ.get("ISOEXT_B200_TUNING", "").split(",")):
            key, _, val = item.partition("=")
            handle.isoext_debug_set_tuning(int(key), int(val))
    return _lib


def last_error() -> str:
    return lib().isoext_last_error().decode()


def check(rc: int) -> None:
    """Raise RuntimeError (the class the reference's std::runtime_error maps to) on failure."""
    if rc != 0:
        raise RuntimeError(last_error() or f"isoext_b200 call failed with status {rc}")


def f3(values):
    a = (C.c_float * 3)(*[float(v) for v in values])
    return a
