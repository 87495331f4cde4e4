"""Pure-torch utilities of the public API (no kernel): ``gaussian_smooth``, ``write_obj``,
``make_grid`` -- host-side mirror of the reference's src/isoext/utils.py:5-83 (same names,
arguments, defaults and results)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

__all__ = ["gaussian_smooth", "write_obj", "make_grid"]


def gaussian_smooth(field: torch.Tensor, sigma: float = 1.0, kernel_size: int | None = None) -> torch.Tensor:
    """Gaussian-blur a (X, Y, Z) scalar field with replicate padding (src/isoext/utils.py:5-39).

    ``kernel_size`` defaults to ``int(6 * sigma) | 1`` (odd).  The result has the input's shape.
    """
    k = (int(6 * sigma) | 1) if kernel_size is None else kernel_size
    taps = torch.arange(k, device=field.device, dtype=field.dtype) - k // 2
    g = torch.exp(-0.5 * (taps / sigma) ** 2)
    g = g / g.sum()
    # Dense k^3 kernel as the outer product of the 1-D taps, exactly like the reference, so the
    # floating-point summation order (and hence every bit of the result) is the same.
    g3 = (g[:, None, None] * g[None, :, None] * g[None, None, :]).view(1, 1, k, k, k)
    pad = k // 2
    vol = F.pad(field.view(1, 1, *field.shape), [pad] * 6, mode="replicate")
    return F.conv3d(vol, g3).reshape(field.shape)


def write_obj(obj_path, v: torch.Tensor | None, f: torch.Tensor | None) -> None:
    """Write a triangle mesh as Wavefront OBJ (src/isoext/utils.py:42-63).

    ``v`` is (N, 3), ``f`` is (M, 3) zero-based.  An empty or ``None`` mesh leaves an empty file,
    which is what ``marching_cubes`` returning ``(None, None)`` leads to in the reference.
    """
    with open(obj_path, "w") as out:
        if v is None or f is None or v.numel() == 0 or f.numel() == 0:
            return
        rows = [f"v {x} {y} {z}\n" for x, y, z in v.tolist()]
        rows += [f"f {a} {b} {c}\n" for a, b, c in (f + 1).tolist()]
        out.writelines(rows)


def make_grid(aabb, res, device: str = "cuda") -> torch.Tensor:
    """(x_res, y_res, z_res, 3) lattice of positions spanning ``aabb`` = [x0, y0, z0, x1, y1, z1]
    (src/isoext/utils.py:66-83).  ``res`` is an int or three ints (points per axis)."""
    if isinstance(res, int):
        res = [res] * 3
    axes = [torch.linspace(aabb[a], aabb[a + 3], res[a], device=device) for a in range(3)]
    return torch.stack(torch.meshgrid(axes, indexing="ij"), dim=-1)
