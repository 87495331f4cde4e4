"""GPU parity tests of SparseGrid: list maintenance, corner positions, filter, marching cubes,
intersections and dual contouring over the active-cell list -- vs the CPU oracle, vs the dense path
(a narrow band that holds every crossing cell must give the dense mesh bit for bit) and vs the
reference's own CUDA build."""
import numpy as np
import pytest
import torch

import fields
import oracle
from isoext_b200 import sdf as S

pytestmark = pytest.mark.gpu


def populate(grid, sdf, level=0.0, chunk=None):
    """The reference's population recipe (tests/conftest.py:39-61 of the reference)."""
    X, Y, Z = grid.shape
    for c in grid.get_potential_cell_indices(chunk or X * Y * Z):
        valid = c[c < (X - 1) * (Y - 1) * (Z - 1)]     # the reference enumerates X*Y*Z ids; ids past the cell range alias
        if len(valid) == 0:
            continue
        pts = grid.get_points_by_cell_indices(valid)
        keep = grid.filter_cell_indices(valid, sdf(pts), level=level)
        if len(keep):
            grid.add_cells(keep)
    if grid.get_num_cells():
        grid.set_values(sdf(grid.get_points()))
    return grid


SDFS = {
    "sphere": (lambda: S.SphereSDF(0.5), (32, 32, 32)),
    "torus_aniso": (lambda: fields.torus(), (24, 40, 32)),
    "csg": (lambda: fields.csg_box_minus_sphere(), (40, 40, 40)),
    "cuboid_exact_hits": (lambda: S.CuboidSDF([1, 1, 1]), (33, 33, 33)),
}


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("name", sorted(SDFS))
def test_sparse_mc_equals_oracle_and_dense(iso, name, method):
    mk, shape = SDFS[name]
    sdf = mk()
    g = populate(iso.SparseGrid(list(shape)), sdf)
    v, f = iso.marching_cubes(g, 0.0, method)
    cells = g.get_cell_indices()
    assert cells.dtype == torch.int32 and bool((cells[1:] > cells[:-1]).all())
    ov, of, _ = oracle.mc_sparse(g.get_values().cpu().numpy(), cells.cpu().numpy(), shape, 0.0, method)
    assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(v.cpu().numpy().view(np.uint32), ov.view(np.uint32))
    # the band holds every crossing cell -> identical to the dense extraction of the same field
    d = iso.UniformGrid(list(shape))
    d.set_values(sdf(d.get_points()))
    dv, df = iso.marching_cubes(d, 0.0, method)
    assert torch.equal(dv.view(torch.int32), v.view(torch.int32)) and torch.equal(df, f)


def test_sparse_mc_vs_reference_cuda_build(iso, ref):
    shape = (32, 32, 32)
    sdf = S.SphereSDF(0.5)
    g = populate(iso.SparseGrid(list(shape)), sdf)
    rg = ref.SparseGrid(list(shape))
    rg.add_cells(g.get_cell_indices())
    assert torch.equal(rg.get_cell_indices(), g.get_cell_indices())
    assert torch.equal(rg.get_points().view(torch.int32), g.get_points().view(torch.int32))
    rg.set_values(g.get_values())
    for method in ("nagae", "lorensen"):
        rv, rf = ref.marching_cubes(rg, 0.0, method)
        v, f = iso.marching_cubes(g, 0.0, method)
        assert torch.equal(rf, f) and torch.equal(rv.view(torch.int32), v.view(torch.int32))
    idx = g.get_cell_indices()[:100].contiguous()
    vals = g.get_values()[:100].contiguous()
    assert torch.equal(rg.filter_cell_indices(idx, vals, 0.0), g.filter_cell_indices(idx, vals, 0.0))
    assert torch.equal(rg.get_points_by_cell_indices(idx).view(torch.int32), g.get_points_by_cell_indices(idx).view(torch.int32))


def test_sparse_with_inconsistent_corner_values_welds_by_position(iso):
    """Neighbouring sparse cells may disagree on shared corner values: welding is positional."""
    shape = (12, 12, 12)
    g = iso.SparseGrid(list(shape))
    gen = torch.Generator().manual_seed(5)
    cells = torch.sort(torch.randperm(11 ** 3, generator=gen)[:400]).values.to(torch.int32).cuda()
    g.add_cells(cells)
    vals = torch.randn((400, 8), generator=gen).cuda()
    g.set_values(vals)
    for method in ("nagae", "lorensen"):
        v, f = iso.marching_cubes(g, 0.0, method)
        ov, of, _ = oracle.mc_sparse(vals.cpu().numpy(), cells.cpu().numpy(), shape, 0.0, method)
        assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(v.cpu().numpy().view(np.uint32), ov.view(np.uint32))


@pytest.mark.parametrize("name", ["sphere", "cuboid_exact_hits"])
def test_sparse_partially_inconsistent_values(iso, name):
    """A band with consistent values in which some cells carry perturbed corners: the z-direction candidate
    de-duplication (csrc/sparse.cu k_sp_mc_dedupe) applies to some shared edges and must not apply to others."""
    make, shape = SDFS[name]
    g = populate(iso.SparseGrid(list(shape)), make())
    cells = g.get_cell_indices()
    vals = g.get_values().clone()
    gen = torch.Generator().manual_seed(11)
    pick = torch.rand(vals.shape[0], generator=gen) < 0.3
    corner = torch.randint(0, 8, (vals.shape[0],), generator=gen)
    delta = (torch.rand(vals.shape[0], generator=gen) - 0.5) * 2e-2
    rows = torch.nonzero(pick).flatten()
    vals[rows.cuda(), corner[rows].cuda()] += delta[rows].cuda()
    vals[rows[::5].cuda(), 0] = 0.0                  # exact level hits on a shared corner: degenerate triangles on one side only
    g.set_values(vals)
    for method in ("nagae", "lorensen"):
        v, f = iso.marching_cubes(g, 0.0, method)
        ov, of, _ = oracle.mc_sparse(vals.cpu().numpy(), cells.cpu().numpy(), shape, 0.0, method)
        assert np.array_equal(f.cpu().numpy(), of) and np.array_equal(v.cpu().numpy().view(np.uint32), ov.view(np.uint32))


@pytest.mark.parametrize("name", sorted(SDFS))
def test_sparse_intersection_and_dc_vs_oracle(iso, name):
    from isoext_b200.sparse import dc_sparse_raw
    mk, shape = SDFS[name]
    g = populate(iso.SparseGrid(list(shape)), mk())
    its = iso.get_intersection(g, 0.0, compute_normals=True)
    cells = g.get_cell_indices().cpu().numpy()
    o = oracle.get_intersection(g.get_values().cpu().numpy(), shape=shape, cell_idx=cells, level=0.0, compute_normals=True)
    assert np.array_equal(its.get_points().cpu().numpy().view(np.uint32), o.points.view(np.uint32))
    assert np.array_equal(its.get_normals().cpu().numpy().view(np.uint32), o.normals.view(np.uint32))
    assert np.array_equal(its.cell_offsets.cpu().numpy().astype(np.int64), o.cell_offsets)
    assert np.array_equal(its.cell_indices.cpu().numpy(), o.cell_indices)
    v, f, dual_v, quads = dc_sparse_raw(g, its, 1e-2, 1e-6, want_quads=True)
    od = oracle.dual_contouring(o, shape)
    cell = 2.0 / (max(shape) - 1)
    assert np.abs(dual_v.cpu().numpy().astype(np.float64) - od["dual_v"]).max() < 1e-4 * cell
    assert np.array_equal(quads.cpu().numpy().astype(np.int64), od["quads"])
    assert f.shape == (2 * len(od["quads"]), 3) and int(f.max()) == len(v) - 1
    # same mesh as the dense dual contouring of the same field
    d = iso.UniformGrid(list(shape))
    d.set_values(mk()(d.get_points()))
    dv, df = iso.dual_contouring(d)
    sv, sf = iso.dual_contouring(g)
    assert torch.equal(dv, sv) and torch.equal(df, sf)


def test_sparse_dc_vs_reference_cuda_build(iso, ref):
    shape = (32, 32, 32)
    g = populate(iso.SparseGrid(list(shape)), S.SphereSDF(0.5))
    rg = ref.SparseGrid(list(shape))
    rg.add_cells(g.get_cell_indices())
    rg.set_values(g.get_values())
    rits = ref.get_intersection(rg, 0.0, True)
    its = iso.get_intersection(g, 0.0, True)
    assert torch.equal(rits.get_points().view(torch.int32), its.get_points().view(torch.int32))
    assert torch.equal(rits.get_normals().view(torch.int32), its.get_normals().view(torch.int32))
    rv, rf = ref.dual_contouring(rg, 0.0)
    v, f = iso.dual_contouring(g, 0.0)
    assert rf.shape == f.shape
    d = torch.cdist(rv.double(), v.double()).min(dim=1).values.max()
    assert float(d) < 5e-4


def test_list_maintenance_contract(iso):
    g = iso.SparseGrid([8, 8, 8])
    assert g.get_num_cells() == 0 and g.get_num_points() == 0
    g.add_cells(torch.tensor([5, 3, 3, 9], dtype=torch.int32, device="cuda"))
    assert g.get_cell_indices().tolist() == [3, 5, 9] and g.get_num_points() == 24
    assert g.get_points().shape == (3, 8, 3) and g.get_values().shape == (3, 8)
    assert bool((g.get_values() == torch.finfo(torch.float32).max).all())
    g.set_values(torch.zeros((3, 8), device="cuda"))
    g.add_cells(torch.tensor([1], dtype=torch.int32, device="cuda"))          # resets ALL values (sparse.cu:94-95)
    assert g.get_cell_indices().tolist() == [1, 3, 5, 9]
    assert bool((g.get_values() == torch.finfo(torch.float32).max).all())
    g.remove_cells(torch.tensor([3, 100], dtype=torch.int32, device="cuda"))
    assert g.get_cell_indices().tolist() == [1, 5, 9]
    with pytest.raises(RuntimeError, match="does not match"):
        g.set_values(torch.zeros((2, 8), device="cuda"))
    with pytest.raises(TypeError):
        g.set_values(torch.zeros((3, 7), device="cuda"))
    with pytest.raises(TypeError):
        g.add_cells(torch.tensor([1.0], device="cuda"))
    chunks = g.get_potential_cell_indices(200)
    assert sum(len(c) for c in chunks) == 512 and chunks[0].dtype == torch.int32 and len(chunks) == 3
    assert g.get_cells().shape == (3, 8)
    # cell index -> corner 0 position: idx = x*(Y-1)*(Z-1) + y*(Z-1) + z
    p = g.get_points_by_cell_indices(torch.tensor([1 * 49 + 2 * 7 + 3], dtype=torch.int32, device="cuda"))
    assert torch.allclose(p[0, 0], torch.tensor([-1 + 2 / 7, -1 + 4 / 7, -1 + 6 / 7], device="cuda"))
    assert torch.allclose(p[0, 7] - p[0, 0], torch.full((3,), 2 / 7, device="cuda"))
    assert iso.marching_cubes(g) == (None, None)


def test_sparse_levels_and_empty(iso):
    sdf = S.SphereSDF(0.5)
    for level in (-0.1, 0.0, 0.1):
        g = populate(iso.SparseGrid([32, 32, 32]), sdf, level=level)
        v, f = iso.marching_cubes(g, level=level)
        assert float((v.norm(dim=-1) - (0.5 + level)).abs().max()) < 5e-3
    e = iso.SparseGrid([16, 16, 16])
    assert iso.marching_cubes(e) == (None, None) and iso.dual_contouring(e) == (None, None)


def test_sparse_1024_sphere_golden_counts(iso):
    """doc/grids.ipynb:307-309: 1024^3 SparseGrid of |p|-0.7 -> 2,416,778 active cells, 2,416,776 vertices."""
    g = populate(iso.SparseGrid([1024] * 3), S.SphereSDF(0.7), chunk=1 << 24)
    assert g.get_num_cells() == 2416778
    v, f = iso.marching_cubes(g)
    assert len(v) == 2416776
    assert float((v.double().norm(dim=-1) - 0.7).abs().max()) < 1e-5


def test_int64_indices_beyond_int_max(iso):
    """Extension: a 4096^3-equivalent grid (6.9e10 points) cannot exist in the reference at all."""
    n = 4096
    g = iso.SparseGrid([n] * 3)
    # cells around the +x pole of the sphere r=0.7: x index near (0.7+1)/2*(n-1)
    xs = torch.arange(3478, 3484, dtype=torch.int64, device="cuda")
    ys = torch.arange(2040, 2056, dtype=torch.int64, device="cuda")
    cx, cy, cz = torch.meshgrid(xs, ys, ys, indexing="ij")
    cand = ((cx * (n - 1) + cy) * (n - 1) + cz).flatten()
    assert int(cand.max()) > 2 ** 31
    sdf = S.SphereSDF(0.7)
    keep = g.filter_cell_indices(cand, sdf(g.get_points_by_cell_indices(cand)))
    g.add_cells(keep)
    g.set_values(sdf(g.get_points()))
    assert g.get_cell_indices().dtype == torch.int64
    v, f = iso.marching_cubes(g)
    assert len(v) > 0 and float((v.double().norm(dim=-1) - 0.7).abs().max()) < 1e-5
    dv, df = iso.dual_contouring(g)
    assert dv is not None and len(df) % 2 == 0


@pytest.mark.parametrize("name", sorted(SDFS))
def test_populate_from_dense_equals_the_reference_recipe(iso, name):
    """SparseGrid.populate_from_dense (one pass of kernels over the field) == the reference's population recipe
    (tests/conftest.py:39-61 of the reference, a Python loop over chunks): same cell list, same (N, 8) values,
    hence the same meshes; also in x-chunks."""
    mk, shape = SDFS[name]
    sdf = mk()
    want = populate(iso.SparseGrid(list(shape)), sdf, chunk=4096)
    d = iso.UniformGrid(list(shape))
    d.set_values(sdf(d.get_points()))
    for chunk in (None, 9, 2):
        g = iso.SparseGrid(list(shape)).populate_from_dense(d, 0.0, x_chunk=chunk)
        assert g.get_cell_indices().dtype == torch.int32
        assert torch.equal(g.get_cell_indices(), want.get_cell_indices())
        assert torch.equal(g.get_values().view(torch.int32), want.get_values().view(torch.int32))
    v, f = iso.marching_cubes(g)
    dv, df = iso.marching_cubes(d)
    assert torch.equal(v.view(torch.int32), dv.view(torch.int32)) and torch.equal(f, df)
    e = iso.SparseGrid(list(shape)).populate_from_dense(torch.ones(shape, device="cuda"))
    assert e.get_num_cells() == 0 and iso.marching_cubes(e) == (None, None)
    with pytest.raises(RuntimeError):
        iso.SparseGrid([8, 8, 8]).populate_from_dense(d)


def test_populate_from_dense_1024_sphere_golden_count(iso):
    """doc/grids.ipynb:307-309 of the reference: 2,416,778 active cells at 1024^3 -- here in one call."""
    n = 1024
    d = iso.UniformGrid([n] * 3)
    ax = fields.axis(n).cuda()
    view = d.values_view()
    sdf = S.SphereSDF(0.7)
    for a in range(0, n, 32):
        view[a:a + 32] = sdf(torch.stack(torch.meshgrid(ax[a:a + 32], ax, ax, indexing="ij"), dim=-1))
    g = iso.SparseGrid([n] * 3).populate_from_dense(d)
    assert g.get_num_cells() == 2416778
    v, f = iso.marching_cubes(g)
    assert len(v) == 2416776
