"""Utilities of the public API: ``gaussian_smooth``, ``write_obj``, ``make_grid`` -- host-side mirror of the
reference's src/isoext/utils.py:5-83 (same names, arguments, defaults and results) -- plus ``write_ply``."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

from . import _lib

__all__ = ["gaussian_smooth", "write_obj", "write_ply", "make_grid"]


def gaussian_smooth(field: torch.Tensor, sigma: float = 1.0, kernel_size: int | None = None, separable: bool | None = None) -> torch.Tensor:
    """Gaussian-blur a (X, Y, Z) scalar field with replicate padding (src/isoext/utils.py:5-39).

    ``kernel_size`` defaults to ``int(6 * sigma) | 1`` (odd).  The result has the input's shape.
    float32 CUDA fields go through the separable kernel of the native library (three 1-D passes, csrc/smooth.cu:
    3k instead of k^3 multiply-adds per voxel, no padded copy); it agrees with the reference's dense conv3d to
    float32 rounding.  ``separable=False`` forces the dense k^3 convolution written exactly like the reference's
    (also used for CPU tensors, other dtypes and kernels longer than 127 taps)."""
    import ctypes as C
    k = (int(6 * sigma) | 1) if kernel_size is None else kernel_size
    taps = torch.arange(k, device=field.device, dtype=field.dtype) - k // 2
    g = torch.exp(-0.5 * (taps / sigma) ** 2)
    g = g / g.sum()
    native_ok = field.is_cuda and field.dtype == torch.float32 and field.dim() == 3 and (k & 1) and 1 <= k <= 127
    if separable is None:
        separable = native_ok
    if separable:
        if not native_ok:
            raise RuntimeError("separable gaussian_smooth needs a 3-D float32 CUDA field and an odd kernel size <= 127")
        src = field.contiguous()
        out, tmp = torch.empty_like(src), torch.empty_like(src)
        w = (C.c_float * k)(*g.tolist())
        X, Y, Z = src.shape
        with torch.cuda.device(src.device):
            _lib.check(_lib.lib().isoext_gaussian_smooth_separable(src.data_ptr(), X, Y, Z, w, k, tmp.data_ptr(), out.data_ptr(),
                                                                   torch.cuda.current_stream().cuda_stream))
        return out
    # Dense k^3 kernel as the outer product of the 1-D taps, exactly like the reference.
    g3 = (g[:, None, None] * g[None, :, None] * g[None, None, :]).view(1, 1, k, k, k)
    pad = k // 2
    vol = F.pad(field.view(1, 1, *field.shape), [pad] * 6, mode="replicate")
    return F.conv3d(vol, g3).reshape(field.shape)


def _host_mesh(v, f):
    v = v.detach().to("cpu", torch.float32).contiguous()
    f = f.detach().to("cpu", torch.int32).contiguous()
    if v.dim() != 2 or v.shape[1] != 3 or f.dim() != 2 or f.shape[1] != 3:
        raise ValueError("v must be (N, 3) and f (M, 3)")
    return v, f


def write_obj(obj_path, v: torch.Tensor | None, f: torch.Tensor | None) -> None:
    """Write a triangle mesh as Wavefront OBJ (src/isoext/utils.py:42-63).

    ``v`` is (N, 3), ``f`` is (M, 3) zero-based.  An empty or ``None`` mesh leaves an empty file, which is what
    ``marching_cubes`` returning ``(None, None)`` leads to in the reference.  The file is byte-identical to the
    reference's (coordinates as Python prints ``repr(float(x))``, one-based ids); it is formatted by the native
    library on all host cores instead of a Python loop (csrc/meshio.cu)."""
    path = os.fsencode(obj_path)
    if v is None or f is None or v.numel() == 0 or f.numel() == 0:
        _lib.check(_lib.lib().isoext_write_obj(path, None, 0, None, 0))
        return
    v, f = _host_mesh(v, f)
    _lib.check(_lib.lib().isoext_write_obj(path, v.data_ptr(), v.shape[0], f.data_ptr(), f.shape[0]))


def write_ply(ply_path, v: torch.Tensor | None, f: torch.Tensor | None) -> None:
    """Binary little-endian PLY (extension; 12 B per vertex + 13 B per face instead of ~40 / ~25 B of OBJ text)."""
    path = os.fsencode(ply_path)
    if v is None or f is None or v.numel() == 0 or f.numel() == 0:
        _lib.check(_lib.lib().isoext_write_ply(path, None, 0, None, 0))
        return
    v, f = _host_mesh(v, f)
    _lib.check(_lib.lib().isoext_write_ply(path, v.data_ptr(), v.shape[0], f.data_ptr(), f.shape[0]))


def make_grid(aabb, res, device: str = "cuda") -> torch.Tensor:
    """(x_res, y_res, z_res, 3) lattice of positions spanning ``aabb`` = [x0, y0, z0, x1, y1, z1]
    (src/isoext/utils.py:66-83).  ``res`` is an int or three ints (points per axis)."""
    if isinstance(res, int):
        res = [res] * 3
    axes = [torch.linspace(aabb[a], aabb[a + 3], res[a], device=device) for a in range(3)]
    return torch.stack(torch.meshgrid(axes, indexing="ij"), dim=-1)
