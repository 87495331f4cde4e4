"""Synthetic analytic fields used by the parity tests and the bench (CPU torch).

The field is ``SDF(P)`` with ``P[i,j,k] = (a_x[i], a_y[j], a_z[k])``, ``a = arange(n)/(n-1)*size+min``
in float32, evaluated in x-slabs to bound host RAM.  For the default AABB [-1,1]^3 this is
bit-identical to the reference's ``grid.get_points()`` (include/utils.cuh:62-80) because the
multiplication by 2 is exact.
"""
from __future__ import annotations

import numpy as np
import torch

from isoext_b200 import sdf as S


def axis(n, lo=-1.0, hi=1.0):
    q = torch.arange(n, dtype=torch.float32) / float(n - 1)
    return q * (np.float32(hi) - np.float32(lo)) + np.float32(lo)


def eval_field(fn, shape, aabb_min=(-1, -1, -1), aabb_max=(1, 1, 1), slab=32) -> torch.Tensor:
    X, Y, Z = shape
    ax, ay, az = (axis(n, lo, hi) for n, lo, hi in zip(shape, aabb_min, aabb_max))
    out = torch.empty((X, Y, Z), dtype=torch.float32)
    for x0 in range(0, X, slab):
        x1 = min(X, x0 + slab)
        P = torch.stack(torch.meshgrid(ax[x0:x1], ay, az, indexing="ij"), dim=-1)
        out[x0:x1] = fn(P)
    return out


def sphere(r=0.5):
    return S.SphereSDF(r)


def torus(R=0.5, r=0.2):
    return S.TorusSDF(R, r)


def csg_box_minus_sphere():
    """BASELINE.json c3/c4: box(1.2) minus sphere(0.5) centred at (.6,.6,.6)."""
    return S.IntersectionOp([S.CuboidSDF([1.2] * 3),
                             S.NegationOp(S.TranslationOp(S.SphereSDF(0.5), [0.6, 0.6, 0.6]))])


def quickstart():
    """doc/quickstart.ipynb: sphere(.75) minus three tori (R=.75, r=.15) about the three axes."""
    t = S.TorusSDF(0.75, 0.15)
    tori = S.UnionOp([t, S.RotationOp(t, [1, 0, 0], 90), S.RotationOp(t, [0, 1, 0], 90)])
    return S.IntersectionOp([S.SphereSDF(0.75), S.NegationOp(tori)])


def gyroid(scale=6.0):
    def fn(p):
        x, y, z = (scale * p[..., i] for i in range(3))
        return torch.sin(x) * torch.cos(y) + torch.sin(y) * torch.cos(z) + torch.sin(z) * torch.cos(x)
    return fn


def noise(shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g, dtype=torch.float32)
