"""Oracle checks at FULL size on windows of x layers: the crossing cells of the layers [a, b) of a
resident field go through the CPU oracle's sparse marching cubes (global grid geometry, so the positions carry the bits
of the full-size grid) and must equal, triangle for triangle and bit for bit, the part of the full-size mesh that
lies in those layers."""
import numpy as np
import torch

import oracle


def check_layers_vs_oracle(lib, vals: torch.Tensor, v: torch.Tensor, f: torch.Tensor, a: int, b: int, level: float = 0.0,
                           method: str = "nagae", aabb=((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))) -> int:
    X, Y, Z = vals.shape
    sub = vals[a:b + 1]
    neg = (sub - level) < 0                                  # the reference's sign test: v - level < 0 in float32
    corners = [neg[dx:dx + (b - a), dy:dy + Y - 1, dz:dz + Z - 1] for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)]
    any_, all_ = corners[0].clone(), corners[0].clone()
    for c in corners[1:]:
        any_ |= c
        all_ &= c
    lx, y, z = torch.nonzero(any_ & ~all_, as_tuple=True)     # lexicographic = ascending cell id
    del corners, any_, all_, neg
    ids = (lx + a) * ((Y - 1) * (Z - 1)) + y * (Z - 1) + z
    # corner k: bit 0 = z, bit 1 = y, bit 2 = x (include/utils.cuh:32-60 of the reference)
    v8 = torch.stack([sub[lx + ((k >> 2) & 1), y + ((k >> 1) & 1), z + (k & 1)] for k in range(8)], dim=1)
    ov, of, _ = oracle.mc_sparse(v8.cpu().numpy(), ids.cpu().numpy(), (X, Y, Z), level, method, aabb[0], aabb[1])
    theirs = ov[of].reshape(-1, 9)
    xa = float(lib.isoext_axis_position(a, X, aabb[0][0], aabb[1][0]))
    xb = float(lib.isoext_axis_position(b, X, aabb[0][0], aabb[1][0]))
    # our triangles of those layers: all three corners inside [px[a], px[b]] (faces are in ascending-cell order, so the
    # selection keeps the oracle's order)
    # (a vertex on an edge inside plane a or b carries x = fma(px, 1-t, t*px) = px +- 1 ulp: widen the window by 2 ulp)
    for _ in range(2):
        xa = float(np.nextafter(np.float32(xa), np.float32(-np.inf)))
        xb = float(np.nextafter(np.float32(xb), np.float32(np.inf)))
    vx = v[:, 0]
    inside = (vx >= xa) & (vx <= xb)
    sel = inside[f.long()].all(dim=1)
    ours = v[f[sel].long()].reshape(-1, 9).cpu().numpy()
    assert ours.shape == theirs.shape, (a, b, ours.shape, theirs.shape)
    assert np.array_equal(ours.view(np.uint32), theirs.view(np.uint32)), f"layers [{a},{b}): triangle bits differ from the oracle"
    return int(theirs.shape[0])
