"""Signed-distance helpers (pure torch; no kernel) -- host-side mirror of the reference's
``isoext.sdf`` module (src/isoext/sdf.py:1-221): same class names, constructor fields and torch
expressions, so a field built with this module is bit-identical to one built with the reference.
They feed the extraction path (they are the input generator of every benchmark) but are not part
of the accelerated hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Sequence

import torch
import torch.nn.functional as F

__all__ = [
    "SDF", "SDFProtocol", "SphereSDF", "TorusSDF", "CuboidSDF", "UnionOp", "SmoothUnionOp",
    "IntersectionOp", "NegationOp", "TranslationOp", "RotationOp", "get_sdf_grad", "get_sdf_normal",
]

# Anything callable on a (..., 3) tensor returning (...) distances (src/isoext/sdf.py:8-21).
SDFProtocol = Callable[[torch.Tensor], torch.Tensor]


def get_sdf_grad(sdf: SDFProtocol, p: torch.Tensor) -> torch.Tensor:
    """d sdf / d p at the points ``p`` (..., 3) via autograd (src/isoext/sdf.py:24-37).

    Like the reference this flips ``requires_grad`` on ``p`` in place.
    """
    p = p.requires_grad_()
    d = sdf(p)
    (g,) = torch.autograd.grad(d, p, grad_outputs=torch.ones_like(d))
    return g


def get_sdf_normal(sdf: SDFProtocol, p: torch.Tensor) -> torch.Tensor:
    """Unit-length gradient of ``sdf`` at ``p`` (src/isoext/sdf.py:40-51)."""
    return F.normalize(get_sdf_grad(sdf, p), dim=-1)


class SDF:
    """Base class: subclasses implement ``__call__(p: (...,3)) -> (...)`` (src/isoext/sdf.py:54-68)."""

    def __call__(self, p: torch.Tensor) -> torch.Tensor:  # pragma: no cover - interface
        raise NotImplementedError


def _each(parts: Sequence[SDFProtocol], p: torch.Tensor) -> torch.Tensor:
    return torch.stack([s(p) for s in parts], dim=-1)


# ---- primitives -------------------------------------------------------------------------------
@dataclass
class SphereSDF(SDF):
    """Origin-centred sphere (src/isoext/sdf.py:71-78)."""
    radius: float

    def __call__(self, p):
        return p.norm(dim=-1) - self.radius


@dataclass
class TorusSDF(SDF):
    """Torus around the z axis: major radius ``R``, tube radius ``r`` (src/isoext/sdf.py:81-95)."""
    R: float
    r: float

    def __call__(self, p):
        ring = p[..., [0, 1]].norm(dim=-1) - self.R
        return torch.stack([ring, p[..., 2]], dim=-1).norm(dim=-1) - self.r


@dataclass
class CuboidSDF(SDF):
    """Axis-aligned box with full edge lengths ``size`` (src/isoext/sdf.py:98-116)."""
    size: list

    def __call__(self, p):
        half = torch.tensor(self.size, device=p.device, dtype=p.dtype) / 2
        q = torch.abs(p) - half
        inside = q.max(dim=-1).values
        outside = torch.norm(torch.maximum(q, torch.zeros_like(q)), dim=-1)
        return outside + torch.minimum(inside, torch.zeros_like(inside))


# ---- combinators --------------------------------------------------------------------------------
@dataclass
class UnionOp(SDF):
    """min over the children (src/isoext/sdf.py:119-127)."""
    sdf_list: list

    def __call__(self, p):
        return _each(self.sdf_list, p).min(dim=-1).values


@dataclass
class IntersectionOp(SDF):
    """max over the children (src/isoext/sdf.py:154-162)."""
    sdf_list: list

    def __call__(self, p):
        return _each(self.sdf_list, p).max(dim=-1).values


@dataclass
class SmoothUnionOp(SDF):
    """Pairwise log-sum-exp blend with sharpness ``k`` (src/isoext/sdf.py:130-151)."""
    sdf_list: list
    k: float

    def __call__(self, p):
        dists = [s(p) for s in self.sdf_list]
        acc = dists[0]
        for other in dists[1:]:
            acc = -self.k * torch.log(torch.exp(-acc / self.k) + torch.exp(-other / self.k))
        return acc


@dataclass
class NegationOp(SDF):
    """Complement (src/isoext/sdf.py:165-172)."""
    sdf: SDF

    def __call__(self, p):
        return -self.sdf(p)


@dataclass
class TranslationOp(SDF):
    """Child moved by ``offset`` (src/isoext/sdf.py:175-183)."""
    sdf: SDF
    offset: list

    def __call__(self, p):
        return self.sdf(p - torch.tensor(self.offset).to(p))


@dataclass
class RotationOp(SDF):
    """Child rotated by ``angle`` about ``axis`` (Rodrigues; src/isoext/sdf.py:186-221)."""
    sdf: SDF
    axis: list
    angle: float
    use_degree: bool = True
    R: torch.Tensor = field(init=False, repr=False, compare=False)

    def __post_init__(self):
        u = F.normalize(torch.tensor(self.axis).float(), dim=0).reshape(3, 1)
        theta = torch.tensor(self.angle).float()
        if self.use_degree:
            theta = torch.deg2rad(theta)
        s, c = torch.sin(theta), torch.cos(theta)
        skew = torch.zeros((3, 3))
        skew[0, 1], skew[0, 2] = -u[2], u[1]
        skew[1, 0], skew[1, 2] = u[2], -u[0]
        skew[2, 0], skew[2, 1] = -u[1], u[0]
        self.R = c * torch.eye(3) + s * skew + (1 - c) * (u @ u.T)

    def __call__(self, p):
        return self.sdf(p @ self.R.to(p))
