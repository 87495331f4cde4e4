"""GPU: drop-in behaviour of the public package surface (names of src/isoext/__init__.py:16-26,
utilities of src/isoext/utils.py and src/isoext/sdf.py) -- the scenarios of the reference's own
tests/test_grid_utils.py, test_sdf_*.py and test_integration.py, with result checks added."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_import_alias_and_names(iso):
    import isoext
    for name in ("UniformGrid", "marching_cubes", "make_grid", "write_obj", "gaussian_smooth"):
        assert hasattr(isoext, name)
    from isoext.sdf import SphereSDF  # noqa: F401
    from isoext.utils import make_grid  # noqa: F401


def test_make_grid_shapes(iso):
    g = iso.make_grid([-1, -1, -1, 1, 1, 1], 16)
    assert g.shape == (16, 16, 16, 3) and g.is_cuda
    g = iso.make_grid([-1, -1, -1, 1, 1, 1], [8, 16, 4], device="cpu")
    assert g.shape == (8, 16, 4, 3)
    assert torch.allclose(g[0, 0, 0], torch.tensor([-1.0, -1, -1])) and torch.allclose(g[-1, -1, -1], torch.ones(3))


def test_write_obj_roundtrip(iso, tmp_path):
    from isoext_b200.sdf import SphereSDF
    grid = iso.UniformGrid([24, 24, 24])
    grid.set_values(SphereSDF(0.5)(grid.get_points()))
    v, f = iso.marching_cubes(grid)
    path = tmp_path / "m.obj"
    iso.write_obj(path, v, f)
    lines = path.read_text().splitlines()
    assert sum(l.startswith("v ") for l in lines) == len(v) and sum(l.startswith("f ") for l in lines) == len(f)
    assert lines[len(v)].split()[1:] == [str(int(i) + 1) for i in f[0].tolist()]
    iso.write_obj(tmp_path / "e.obj", None, None)
    assert (tmp_path / "e.obj").read_text() == ""


def test_make_grid_field_equals_get_points_field(iso):
    from isoext_b200.sdf import SphereSDF
    res = 16
    grid = iso.UniformGrid([res] * 3)
    grid.set_values(SphereSDF(0.5)(iso.make_grid([-1, -1, -1, 1, 1, 1], res)))
    v, f = iso.marching_cubes(grid)
    assert len(v) > 0 and f.shape[1] == 3


def test_csg_integration(iso):
    from isoext_b200 import sdf as S
    shape = S.IntersectionOp([S.SphereSDF(0.75), S.NegationOp(S.UnionOp([
        S.TorusSDF(0.75, 0.15), S.RotationOp(S.TorusSDF(0.75, 0.15), [1, 0, 0], 90)]))])
    grid = iso.UniformGrid([64] * 3)
    grid.set_values(shape(grid.get_points()))
    for method in ("nagae", "lorensen"):
        v, f = iso.marching_cubes(grid, method=method)
        assert v.shape[1] == 3 and f.shape[1] == 3 and int(f.max()) == len(v) - 1 and int(f.min()) == 0


def test_gaussian_smooth_occupancy_workflow(iso):
    """doc/occupancy_grids.ipynb: 64^3 binary occupancy, level 0.5: 9,168 V / 18,332 F; sigma=5: 8,232 V / 16,460 F."""
    from isoext_b200.sdf import SphereSDF
    grid = iso.UniformGrid([64] * 3)
    occ = (SphereSDF(0.7)(grid.get_points()) < 0).float()
    grid.set_values(occ)
    v, f = iso.marching_cubes(grid, level=0.5)
    assert (len(v), len(f)) == (9168, 18332)
    grid.set_values(iso.gaussian_smooth(occ, sigma=5.0).contiguous())
    v, f = iso.marching_cubes(grid, level=0.5)
    assert (len(v), len(f)) == (8232, 16460)


def test_sdf_primitives_and_ops(iso):
    from isoext_b200 import sdf as S
    p = torch.tensor([[0.0, 0, 0], [1.0, 0, 0], [0.25, 0.25, 0]], device="cuda")
    assert torch.allclose(S.SphereSDF(0.5)(p), torch.tensor([-0.5, 0.5, 0.25 * 2 ** 0.5 - 0.5], device="cuda"))
    assert S.TorusSDF(0.5, 0.2)(p)[0] > 0 and S.CuboidSDF([1, 1, 1])(p)[0] < 0
    u = S.UnionOp([S.SphereSDF(0.3), S.TranslationOp(S.SphereSDF(0.3), [1.0, 0, 0])])
    assert u(p)[0] < 0 and u(p)[1] < 0
    n = S.get_sdf_normal(S.SphereSDF(0.5), p[1:].clone())
    assert torch.allclose(n.norm(dim=-1), torch.ones(2, device="cuda"), atol=1e-6)
    sm = S.SmoothUnionOp([S.SphereSDF(0.3), S.SphereSDF(0.4)], k=0.1)
    assert float(sm(p)[0]) <= -0.4 + 1e-6


@pytest.mark.parametrize("shape,sigma,k", [((40, 33, 57), 1.0, None), ((16, 16, 16), 2.5, None), ((9, 70, 5), 0.8, 9), ((64, 64, 64), 1.5, 5)])
def test_gaussian_smooth_separable_kernel_equals_dense_conv3d(iso, shape, sigma, k):
    """csrc/smooth.cu (three 1-D passes, clamped indices) vs the reference's dense k^3 conv3d with replicate padding
    (src/isoext/utils.py:5-39): float tolerance 1e-6 of the field's range (the summation order differs)."""
    from isoext_b200 import utils as U
    gen = torch.Generator().manual_seed(7)
    field = torch.randn(shape, generator=gen).cuda() * 3.0 + 1.0
    a = U.gaussian_smooth(field, sigma, k)                       # native separable path (default on CUDA float32)
    b = U.gaussian_smooth(field, sigma, k, separable=False)      # the reference's expression
    assert a.shape == field.shape and a.dtype == torch.float32
    tol = 1e-6 * float(field.max() - field.min())
    assert float((a - b).abs().max()) <= tol
    # a constant field stays constant; the input is not modified
    c = torch.full(shape, 2.5, device="cuda")
    assert float((U.gaussian_smooth(c, sigma, k) - 2.5).abs().max()) < 1e-6
    assert torch.equal(field, torch.randn(shape, generator=torch.Generator().manual_seed(7)).cuda() * 3.0 + 1.0)
