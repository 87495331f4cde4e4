"""GPU parity tests of dense marching cubes: product (CUDA, through the C-ABI) vs the CPU oracle,
vs the committed reference fixtures, and -- where oracle/_ref is present -- vs the reference's own
CUDA build.  Bar: faces and vertex bits identical (north_star asks for 1e-6 relative on positions;
we assert bit equality, which is stronger)."""
from pathlib import Path

import numpy as np
import pytest
import torch

import fields
import oracle
from isoext_b200 import sdf as S

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
BOX = ((-1, -1, -1), (1, 1, 1))


def run_ours(iso, vals, level=0.0, method="nagae", aabb=BOX):
    g = iso.UniformGrid(list(vals.shape), aabb[0], aabb[1])
    g.set_values(vals.cuda())
    return iso.marching_cubes(g, level, method)


def assert_same_mesh(v, f, ov, of):
    if len(of) == 0:
        assert v is None and f is None
        return
    assert v.dtype == torch.float32 and f.dtype == torch.int32
    v, f = v.cpu().numpy(), f.cpu().numpy()
    assert v.shape == ov.shape and f.shape == of.shape, (v.shape, ov.shape, f.shape, of.shape)
    assert np.array_equal(f, of), "face connectivity differs"
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32)), "vertex bits differ"


CASES = {
    "c1_sphere64": lambda: (fields.eval_field(S.SphereSDF(0.5), (64, 64, 64)), 0.0, BOX),
    "sphere64_lvl+0.1": lambda: (fields.eval_field(S.SphereSDF(0.5), (64, 64, 64)), 0.1, BOX),
    "sphere64_lvl-0.1": lambda: (fields.eval_field(S.SphereSDF(0.5), (64, 64, 64)), -0.1, BOX),
    "aniso_16x32x48": lambda: (fields.eval_field(S.SphereSDF(0.5), (16, 32, 48)), 0.0, BOX),
    "aniso_8x64x16": lambda: (fields.eval_field(S.SphereSDF(0.5), (8, 64, 16)), 0.0, BOX),
    "odd_33x35x37_torus": lambda: (fields.eval_field(fields.torus(), (33, 35, 37)), 0.0, BOX),
    "cuboid65_exact_level_hits": lambda: (fields.eval_field(S.CuboidSDF([1, 1, 1]), (65, 65, 65)), 0.0, BOX),
    "gyroid64": lambda: (fields.eval_field(fields.gyroid(6.0), (64, 64, 64)), 0.0, BOX),
    "noise40": lambda: (fields.noise((40, 40, 40)), 0.0, BOX),
    "noise_37x41x131": lambda: (fields.noise((37, 41, 131), 1), 0.0, BOX),
    "noise_20x24x128_spanpath": lambda: (fields.noise((20, 24, 128), 3), 0.0, BOX),
    "noise_7x5x256_spanpath": lambda: (fields.noise((7, 5, 256), 4), 0.0, BOX),
    "sphere_33x40x128_spanpath": lambda: (fields.eval_field(S.SphereSDF(0.8), (33, 40, 128)), 0.0, BOX),
    # span-summary compaction (dense.cuh): 3 / 4 / 8 / 16 spans per row (1, 4 and 8 spans per summary load), rows of
    # more than 32 spans (thread-per-row fallback), and rows lying in a z-aligned face (heavy-row fill)
    "noise_5x6x384_spans3": lambda: (fields.noise((5, 6, 384), 5), 0.0, BOX),
    "sphere_9x300x512_spans4": lambda: (fields.eval_field(S.SphereSDF(0.8), (9, 300, 512)), 0.0, BOX),
    "torus_6x40x1024_spans8": lambda: (fields.eval_field(fields.torus(), (6, 40, 1024)), 0.0, BOX),
    "noise_3x4x2048_spans16": lambda: (fields.noise((3, 4, 2048), 6), 0.0, BOX),
    "sphere_3x5x4224_spans33_serial": lambda: (fields.eval_field(S.SphereSDF(0.8), (3, 5, 4224)), 0.0, BOX),
    "yface_4x9x640_heavy_rows": lambda: (_plane((4, 9, 640), (0.0, 1.0, 0.0), 0.13), 0.0, BOX),
    "surface_crosses_boundary": lambda: (fields.eval_field(S.SphereSDF(1.2), (48, 48, 48)), 0.0, BOX),
    "shifted_aabb": lambda: (fields.eval_field(S.SphereSDF(0.5), (40, 40, 40)), 0.0, ((0, -2, 5), (3, 1, 6))),
    "tiny_2x2x2": lambda: (torch.tensor([[[-1., 1], [1, 1]], [[1, 1], [1, -1]]]), 0.0, BOX),
    "thin_3x2x200": lambda: (fields.noise((3, 2, 200), 2), 0.0, BOX),
    "occupancy64_level0.5": lambda: ((fields.eval_field(S.SphereSDF(0.7), (64, 64, 64)) < 0).float(), 0.5, BOX),
    "all_positive_empty": lambda: (torch.ones(8, 8, 8), 0.0, BOX),
    "all_negative_empty": lambda: (-torch.ones(9, 8, 7), 0.0, BOX),
    "single_exact_zero_point": lambda: (_single_zero(), 0.0, BOX),
    # oversized sort buckets (> 4096 vertices in one x layer) and how they are cut a second time (segsort.cuh):
    "xface_6x200x200_second_level_by_y": lambda: (_plane((6, 200, 200), (1.0, 0.0, 0.0), 0.13), 0.0, BOX),
    "xface_5x3x9000_radix_last_resort": lambda: (_plane((5, 3, 9000), (1.0, 0.0, 0.0), 0.13), 0.0, BOX),
    "tilted_7x160x160_second_level_by_x": lambda: (_plane((7, 160, 160), (1.0, 0.013, 0.007), 0.05), 0.0, BOX),
    "xface_on_plane_5x96x96_exact_hits": lambda: (_plane((5, 96, 96), (1.0, 0.0, 0.0), 0.5), 0.0, BOX),
    "box_4x150x150_two_xfaces": lambda: (fields.eval_field(S.CuboidSDF([0.9, 1.5, 1.5]), (4, 150, 150)), 0.0, BOX),
}


def _plane(shape, normal, offset):
    """n . p - offset on the [-1,1]^3 grid: a face (nearly) perpendicular to x puts all its vertices in one x layer."""
    ax = [fields.axis(n) for n in shape]
    X, Y, Z = torch.meshgrid(*ax, indexing="ij")
    return (normal[0] * X + normal[1] * Y + normal[2] * Z - offset).contiguous()


def _single_zero():
    t = -torch.ones(5, 5, 5)
    t[2, 2, 2] = 0.0   # every incident triangle is degenerate: reference output is empty
    return t


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_parity_vs_oracle(iso, name, method):
    vals, level, aabb = CASES[name]()
    v, f = run_ours(iso, vals, level, method, aabb)
    ov, of, _ = oracle.mc_dense(vals.numpy(), level, method, aabb[0], aabb[1])
    assert_same_mesh(v, f, ov, of)


@pytest.mark.parametrize("name,fn,counts", [
    ("torus256", fields.torus, (91850, 183700)),
    ("csg256", fields.csg_box_minus_sphere, (138624, 277244)),
    ("quickstart256", fields.quickstart, (196176, 392348)),
])
def test_parity_256(iso, name, fn, counts):
    vals = fields.eval_field(fn(), (256, 256, 256))
    v, f = run_ours(iso, vals)
    assert (len(v), len(f)) == counts
    ov, of, _ = oracle.mc_dense(vals.numpy())
    assert_same_mesh(v, f, ov, of)


@pytest.mark.parametrize("path", sorted(GOLDEN.glob("mc_*.npz")), ids=lambda p: p.stem)
def test_parity_vs_reference_fixtures(iso, path):
    g = np.load(path)
    v, f = run_ours(iso, torch.from_numpy(g["values"]), float(g["level"]), str(g["method"]),
                    (tuple(g["aabb_min"]), tuple(g["aabb_max"])))
    assert_same_mesh(v, f, g["v"], g["f"])


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("name", ["c1_sphere64", "cuboid65_exact_level_hits", "noise40", "noise_20x24x128_spanpath",
                                  "surface_crosses_boundary", "shifted_aabb"])
def test_parity_vs_reference_cuda_build(iso, ref, name, method):
    vals, level, aabb = CASES[name]()
    v, f = run_ours(iso, vals, level, method, aabb)
    rg = ref.UniformGrid(list(vals.shape), aabb[0], aabb[1])
    rg.set_values(vals.cuda())
    rv, rf = ref.marching_cubes(rg, level, method)
    assert torch.equal(rf, f)
    assert torch.equal(rv.view(torch.int32), v.view(torch.int32))


def test_c2_512_torus_vs_reference_and_topology(iso, ref):
    """BASELINE.json configs[1] at full size: bit parity with the reference CUDA build (the oracle
    would take ~10 s), plus size-independent properties: closed 2-manifold, Euler characteristic 0."""
    vals = fields.eval_field(fields.torus(), (512, 512, 512)).cuda()
    g = iso.UniformGrid([512] * 3)
    g.set_values(vals)
    v, f = iso.marching_cubes(g)
    assert (len(v), len(f)) == (372576, 745152)
    rg = ref.UniformGrid([512] * 3)
    rg.set_values(vals)
    rv, rf = ref.marching_cubes(rg)
    assert torch.equal(rf, f) and torch.equal(rv.view(torch.int32), v.view(torch.int32))
    check_closed_manifold(v, f, euler=0)


def check_closed_manifold(v, f, euler):
    f64 = f.long()
    e = torch.cat([f64[:, [0, 1]], f64[:, [1, 2]], f64[:, [2, 0]]])
    key = torch.minimum(e[:, 0], e[:, 1]) * len(v) + torch.maximum(e[:, 0], e[:, 1])
    uniq, cnt = torch.unique(key, return_counts=True)
    assert bool((cnt == 2).all()), "every edge of a closed surface is shared by exactly two triangles"
    assert len(v) - len(uniq) + len(f) == euler
    # V is strictly increasing in lexicographic (x,y,z) order (the reference's sort + unique)
    a, b = v[:-1], v[1:]
    lt = (a[:, 0] < b[:, 0]) | ((a[:, 0] == b[:, 0]) & ((a[:, 1] < b[:, 1]) | ((a[:, 1] == b[:, 1]) & (a[:, 2] < b[:, 2]))))
    assert bool(lt.all())
    # every vertex is referenced
    assert len(torch.unique(f64)) == len(v)


def test_1024_properties(iso):
    """North-star size (the reference is invalid here: 32-bit cell_idx*8 overflow, include/utils.cuh:48,95).
    Checked through size-independent properties + linearity of the level set."""
    n = 1024
    g = iso.UniformGrid([n] * 3)
    ax = fields.axis(n).cuda()
    vals = g.values_view()
    for x0 in range(0, n, 64):   # sphere r=0.7 built on the GPU slab by slab
        P = torch.stack(torch.meshgrid(ax[x0:x0 + 64], ax, ax, indexing="ij"), dim=-1)
        vals[x0:x0 + 64] = P.norm(dim=-1) - 0.7
        del P
    v, f = iso.marching_cubes(g)
    assert len(v) == 2416776   # doc/grids.ipynb:307-309 records this count for the same field
    check_closed_manifold(v, f, euler=2)
    # the CPU oracle on windows of x layers of the full-size grid: where the sphere is tangent to the planes, at its
    # equator and in between (bit-equal triangles; the reference itself is invalid at this size)
    import fullsize
    from isoext_b200 import _lib
    checked = sum(fullsize.check_layers_vs_oracle(_lib.lib(), vals, v, f, a, b) for a, b in ((152, 156), (300, 302), (510, 513), (868, 871)))
    assert checked > 50000
    r = v.double().norm(dim=-1)
    assert float((r - 0.7).abs().max()) < 2e-6 * n / 64   # linear interpolation error of a sphere SDF
    # the same surface through level shift: marching_cubes(values, L) == marching_cubes(values - L', L - L') topologically
    v2, f2 = iso.marching_cubes(g, level=0.05)
    check_closed_manifold(v2, f2, euler=2)
    assert len(v2) > len(v)   # level 0.05 of |p|-0.7 is the sphere of radius 0.75


# ---- API behaviour mirrored from the reference binding (src/isoext_ext.cu:95-168) ----------------
def test_unknown_method_raises_runtime_error(iso):
    g = iso.UniformGrid([8, 8, 8])
    with pytest.raises(RuntimeError, match="Unknown method: foo"):
        iso.marching_cubes(g, 0.0, "foo")


def test_default_grid_is_empty_and_returns_none(iso):
    g = iso.UniformGrid([8, 8, 8])   # filled with FLT_MAX
    assert iso.marching_cubes(g) == (None, None)


def test_set_values_contract(iso):
    g = iso.UniformGrid([4, 5, 6])
    assert g.get_num_points() == 120 and g.get_num_cells() == 60
    with pytest.raises(RuntimeError, match="different shapes"):
        g.set_values(torch.zeros(4, 5, 7, device="cuda"))
    with pytest.raises(TypeError):
        g.set_values(torch.zeros(4, 5, 6))                       # CPU tensor
    with pytest.raises(TypeError):
        g.set_values(torch.zeros(4, 5, 6, device="cuda", dtype=torch.float64))
    t = torch.rand(4, 5, 6, device="cuda")
    g.set_values(t)
    got = g.get_values()
    assert torch.equal(got, t) and got.data_ptr() != g.values_view().data_ptr()   # a copy, like the reference


def test_get_points_bit_exact(iso):
    shape, lo, hi = (31, 17, 53), (-1, 0, 2), (1, 5, 2.5)
    g = iso.UniformGrid(list(shape), lo, hi)
    p = g.get_points()
    assert p.shape == (*shape, 3)
    assert np.array_equal(p.cpu().numpy().view(np.uint32), oracle.points_dense(shape, lo, hi).view(np.uint32))


def test_get_points_vs_reference_build(iso, ref):
    shape, lo, hi = (31, 17, 53), (-1, 0, 2), (1, 5, 2.5)
    g = iso.UniformGrid(list(shape), lo, hi)
    rg = ref.UniformGrid(list(shape), lo, hi)
    assert torch.equal(rg.get_points().view(torch.int32), g.get_points().view(torch.int32))


def test_repeated_calls_reuse_workspace_and_are_deterministic(iso):
    vals = fields.eval_field(fields.torus(), (96, 96, 128)).cuda()
    g = iso.UniformGrid([96, 96, 128])
    g.set_values(vals)
    v0, f0 = iso.marching_cubes(g)
    for _ in range(3):
        v, f = iso.marching_cubes(g)
        assert torch.equal(v, v0) and torch.equal(f, f0)


@pytest.mark.parametrize("name", ["xface_6x200x200_second_level_by_y", "xface_5x3x9000_radix_last_resort",
                                  "tilted_7x160x160_second_level_by_x", "box_4x150x150_two_xfaces", "cuboid65_exact_level_hits"])
def test_fast_path_with_oversized_sort_buckets(iso, name):
    """Second and later extractions of a grid take the single-sync path (isoext_mc_dense_run), whose second bucket
    level and radix last resort are only enqueued when the previous call reported them as needed: every call must
    give the oracle's mesh, also right after the field changed from one with oversized buckets to one without."""
    vals, level, aabb = CASES[name]()
    ov, of, _ = oracle.mc_dense(vals.numpy(), level, "nagae", aabb[0], aabb[1])
    g = iso.UniformGrid(list(vals.shape), aabb[0], aabb[1])
    g.set_values(vals.cuda())
    for _ in range(4):
        v, f = iso.marching_cubes(g, level)
        assert_same_mesh(v, f, ov, of)
    other = torch.from_numpy(np.random.default_rng(5).standard_normal(tuple(vals.shape)).astype(np.float32))
    oov, oof, _ = oracle.mc_dense(other.numpy(), 0.0, "nagae", aabb[0], aabb[1])
    g.set_values(other.cuda())
    for _ in range(2):
        v, f = iso.marching_cubes(g, 0.0)
        assert_same_mesh(v, f, oov, oof)
    g.set_values(vals.cuda())
    for _ in range(2):
        v, f = iso.marching_cubes(g, level)
        assert_same_mesh(v, f, ov, of)


def test_capacity_retry_path(iso):
    """A noise field has far more active cells than the initial capacity guess."""
    vals = fields.noise((96, 96, 96), 5)
    v, f = run_ours(iso, vals)
    ov, of, _ = oracle.mc_dense(vals.numpy())
    assert_same_mesh(v, f, ov, of)
