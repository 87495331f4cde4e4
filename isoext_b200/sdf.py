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
    "compile_sdf", "SdfProgram",
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


# ---- analytic programs for the fused path (extension; csrc/sdfprog.cuh) ----------------------------------------
_OP = dict(END=0, RESET=1, TRANSLATE=2, ROTATE=3, SPHERE=4, TORUS=5, CUBOID=6, UNION=7, INTER=8, SMOOTH=9, NEG=10)
_MAX_OPS, _MAX_CONSTS, _MAX_STACK = 240, 256, 12


def _f32(x) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


class SdfProgram:
    """A tree of the built-in SDF classes compiled to the postfix program the extraction kernels interpret
    (csrc/sdfprog.cuh).  ``blob`` is the host image of ``struct SdfProg``; ``device(dev)`` uploads it once per device.

    Every leaf carries the chain of point transforms above it (outermost first, exactly the order in which the
    torch expressions of this module apply them), combinators work on a small value stack.  Constants are the
    float32 values the torch expressions use (e.g. ``size / 2`` of a cuboid and the rotation matrix are computed
    with torch, in float32)."""

    def __init__(self, ops: list, consts: list):
        import struct
        if len(ops) > _MAX_OPS or len(consts) > _MAX_CONSTS:
            raise RuntimeError("SDF tree too large for a device program")
        self.ops, self.consts = list(ops), list(consts)
        self.blob = (struct.pack("<IIfI", len(ops), len(consts), 1.0 + 1e-3, 0) + bytes(ops) + bytes(_MAX_OPS - len(ops))
                     + struct.pack(f"<{_MAX_CONSTS}f", *(list(consts) + [0.0] * (_MAX_CONSTS - len(consts)))))
        self._dev = {}

    def device(self, device) -> torch.Tensor:
        device = torch.device(device)
        t = self._dev.get(device)
        if t is None:
            from . import _lib
            assert len(self.blob) == _lib.lib().isoext_sdf_program_bytes(), "SdfProg layout mismatch"
            t = torch.frombuffer(bytearray(self.blob), dtype=torch.uint8).to(device)
            self._dev[device] = t
        return t


def compile_sdf(sdf) -> SdfProgram:
    """Compile a tree of SphereSDF / TorusSDF / CuboidSDF / UnionOp / SmoothUnionOp / IntersectionOp / NegationOp /
    TranslationOp / RotationOp to an :class:`SdfProgram`.  Raises ``TypeError`` for anything else (arbitrary callables
    cannot run inside a kernel: evaluate them with torch and use ``set_values``)."""
    ops, consts = [], []
    depth = [0, 0]   # current / maximum value-stack depth

    def push():
        depth[0] += 1
        depth[1] = max(depth[1], depth[0])

    def leaf_prefix(chain):
        ops.append(_OP["RESET"])
        for kind, data in chain:
            ops.append(_OP[kind])
            consts.extend(data)

    def walk(node, chain):
        if isinstance(node, SphereSDF):
            leaf_prefix(chain); ops.append(_OP["SPHERE"]); consts.append(_f32(node.radius)); push()
        elif isinstance(node, TorusSDF):
            leaf_prefix(chain); ops.append(_OP["TORUS"]); consts.extend([_f32(node.R), _f32(node.r)]); push()
        elif isinstance(node, CuboidSDF):
            half = (torch.tensor(node.size, dtype=torch.float32) / 2).tolist()
            if len(half) != 3:
                raise TypeError("CuboidSDF.size must have three elements")
            leaf_prefix(chain); ops.append(_OP["CUBOID"]); consts.extend(half); push()
        elif isinstance(node, (UnionOp, IntersectionOp, SmoothUnionOp)):
            kids = list(node.sdf_list)
            if not 1 <= len(kids) <= _MAX_STACK - 2:
                raise TypeError("combinators take between 1 and 10 children in a device program")
            for k in kids:
                walk(k, chain)
            if isinstance(node, SmoothUnionOp):
                ops.extend([_OP["SMOOTH"], len(kids)]); consts.append(_f32(node.k))
            else:
                ops.extend([_OP["UNION" if isinstance(node, UnionOp) else "INTER"], len(kids)])
            depth[0] -= len(kids) - 1
        elif isinstance(node, NegationOp):
            walk(node.sdf, chain); ops.append(_OP["NEG"])
        elif isinstance(node, TranslationOp):
            off = torch.tensor(node.offset).to(torch.float32).tolist()
            walk(node.sdf, chain + [("TRANSLATE", off)])
        elif isinstance(node, RotationOp):
            walk(node.sdf, chain + [("ROTATE", node.R.to(torch.float32).reshape(-1).tolist())])
        else:
            raise TypeError(f"compile_sdf: {type(node).__name__} is not one of the built-in SDF classes")

    walk(sdf, [])
    if depth[1] > _MAX_STACK:
        raise RuntimeError("SDF tree too deep for a device program")
    return SdfProgram(ops, consts)
