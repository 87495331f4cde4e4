"""GPU tests of csrc/setops.cu: SparseGrid list maintenance as kernels (add_cells = sort + unique + union,
remove_cells = set difference, filter compaction; src/grid/sparse.cu:71-126,150-179 of the reference),
UniformGrid.get_cells (src/grid/uniform.cu:42-51) and the per-layer vertex histogram of the slab balancer."""
import numpy as np
import pytest
import torch

import fields
import oracle
from isoext_b200 import sdf as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(5, 4, 3), (17, 9, 33), (2, 2, 2), (64, 3, 130)])
def test_get_cells_equals_oracle_and_reference(iso, ref, shape):
    g = iso.UniformGrid(list(shape))
    c = g.get_cells()
    assert tuple(c.shape) == (shape[0] - 1, shape[1] - 1, shape[2] - 1, 8)
    assert c.dtype in (torch.uint32, torch.int64)
    got = c.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, oracle.cells_dense(shape))
    rc = ref.UniformGrid(list(shape)).get_cells().cpu().numpy().view(np.uint32).astype(np.int64)
    assert np.array_equal(got, rc)


def test_get_cells_int64_beyond_2_32_points(iso):
    """X*Y*Z > 2^32: the reference overflows (uint); here the ids widen to int64.  Checked on the last cells."""
    from isoext_b200 import _lib
    from isoext_b200.grid import _stream_ptr
    X, Y, Z = 2, 2, (1 << 32) // 4 + 8
    out = torch.empty((X - 1, Y - 1, Z - 1, 8), dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().isoext_grid_cells_dense(X, Y, Z, 1, out.data_ptr(), _stream_ptr()))
    last = out[0, 0, -1].cpu().tolist()
    z = Z - 2
    assert last == [z, z + 1, Z + z, Z + z + 1, 2 * Z + z, 2 * Z + z + 1, 3 * Z + z, 3 * Z + z + 1]
    assert last[-1] > 2 ** 32
    with pytest.raises(RuntimeError):
        _lib.check(_lib.lib().isoext_grid_cells_dense(X, Y, Z, 0, out.data_ptr(), _stream_ptr()))


@pytest.mark.parametrize("n_old,n_new,hi", [(0, 1, 10), (0, 1000, 50), (5000, 7000, 3000), (100000, 300000, 1 << 20),
                                              (1000, 4000, 1 << 40), (3, 0, 10)])
def test_add_and_remove_cells_equal_torch_set_ops(iso, n_old, n_new, hi):
    gen = torch.Generator().manual_seed(n_old + n_new)
    big = hi > 2 ** 31
    dt = torch.int64 if big else torch.int32
    a = torch.randint(0, hi, (n_old,), generator=gen, dtype=torch.int64).to(dt).cuda()
    b = torch.randint(0, hi, (n_new,), generator=gen, dtype=torch.int64).to(dt).cuda()
    g = iso.SparseGrid([4096, 4096, 4096] if big else [1200, 1200, 1200])
    if n_old:
        g.add_cells(a)
        assert torch.equal(g.get_cell_indices().long(), torch.unique(a.long()))
    g.add_cells(b)
    want = torch.unique(torch.cat([a.long(), b.long()]))
    got = g.get_cell_indices()
    assert got.dtype == dt and torch.equal(got.long(), want)
    assert g.get_values().shape == (len(want), 8)
    # remove: random subset + ids that are not in the list + duplicates
    rm = torch.cat([want[torch.randperm(len(want), generator=gen)[:len(want) // 3].cuda()] if len(want) else want,
                    torch.randint(0, hi, (17,), generator=gen, dtype=torch.int64).cuda()])
    rm = torch.cat([rm, rm[:5]]).to(dt)
    g.remove_cells(rm)
    want2 = want[~torch.isin(want, rm.long())]
    assert torch.equal(g.get_cell_indices().long(), want2)
    g.remove_cells(g.get_cell_indices())
    assert g.get_num_cells() == 0 and g.get_values().shape == (0, 8)


def test_filter_compaction_is_stable_and_typed(iso):
    g = iso.SparseGrid([64, 64, 64])
    gen = torch.Generator().manual_seed(3)
    for dt in (torch.int32, torch.int64):
        for n in (0, 1, 257, 100001):
            idx = torch.randperm(max(n, 1), generator=gen)[:n].to(dt).cuda()      # NOT sorted: order must be preserved
            vals = torch.randn((n, 8), generator=gen).cuda()
            vals[::3] = vals[::3].abs() + 0.1                                     # a third of the rows cannot cross
            out = g.filter_cell_indices(idx, vals, 0.0)
            neg = vals < 0
            want = idx[neg.any(1) & ~neg.all(1)]
            assert out.dtype == dt and torch.equal(out, want)
    # misaligned (odd storage offset) value rows are accepted (the reference takes any contiguous tensor)
    buf = torch.randn(1 + 8 * 100, generator=gen).cuda()
    v = buf[1:].view(100, 8)
    idx = torch.arange(100, dtype=torch.int32, device="cuda")
    neg = v < 0
    assert torch.equal(g.filter_cell_indices(idx, v, 0.0), idx[neg.any(1) & ~neg.all(1)])


def test_reference_population_recipe_unfiltered(iso, ref):
    """tests/conftest.py:39-61 of the reference: X*Y*Z candidate ids (past the cell range) straight through
    get_points_by_cell_indices / filter_cell_indices / add_cells -- same list as the reference build."""
    shape = (24, 20, 28)
    sdf = S.SphereSDF(0.5)
    g, rg = iso.SparseGrid(list(shape)), ref.SparseGrid(list(shape))
    for chunk in g.get_potential_cell_indices(5000):
        pts = g.get_points_by_cell_indices(chunk)
        assert torch.equal(pts.view(torch.int32), rg.get_points_by_cell_indices(chunk).view(torch.int32))
        vals = sdf(pts)
        keep = g.filter_cell_indices(chunk, vals, level=0.0)
        assert torch.equal(keep, rg.filter_cell_indices(chunk, vals, 0.0))
        if len(keep):
            g.add_cells(keep)
            rg.add_cells(keep)
    assert g.get_num_cells() > 0 and torch.equal(g.get_cell_indices(), rg.get_cell_indices())


def test_vertex_layer_histogram_kernel_equals_torch(iso):
    from isoext_b200 import dist as idist
    n = 96
    g = iso.UniformGrid([n] * 3)
    g.set_values(fields.eval_field(fields.csg_box_minus_sphere(), (n, n, n)).cuda())
    v, _ = iso.marching_cubes(g)
    h = idist.vertex_layer_histogram(v, n, -1.0, 1.0)
    t = (v[:, 0].double() + 1.0) / 2.0 * (n - 1)
    want = torch.bincount(t.floor().clamp_(0, n - 2).long(), minlength=n - 1).to(torch.float64)
    assert torch.equal(h, want) and float(h.sum()) == len(v)
    assert torch.equal(idist.vertex_layer_histogram(v.cpu(), n, -1.0, 1.0), want.cpu())
