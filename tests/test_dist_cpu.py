"""CPU tests of the multi-GPU slab logic (isoext_b200/dist.py): partitioning, the gloo halo exchange at
world_size 2 and 3, and -- using the CPU oracle in place of the CUDA kernels -- the vertex-ownership /
global-id scheme itself: per-rank parts must concatenate to exactly the single-device mesh."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fields
import oracle
from isoext_b200 import _lib
from isoext_b200 import dist as idist
from isoext_b200 import sdf as S


def test_partition_covers_all_layers():
    for X, world in [(64, 1), (64, 2), (65, 3), (2048, 8), (17, 8)]:
        c = idist.partition_cells(X, world)
        assert c[0] == 0 and c[-1] == X - 1 and all(b - a >= 2 for a, b in zip(c, c[1:]))
    with pytest.raises(RuntimeError):
        idist.partition_cells(8, 8)


def test_slab_plan_halo_geometry():
    X, world = 64, 4
    for depth in (1, 2):        # 1: marching cubes, 2: slabs that also serve dual contouring
        plans = [idist.slab_plan(X, r, world, halo_below=depth) for r in range(world)]
        assert plans[0]["halo_below"] == [] and plans[-1]["halo_above"] == []
        for r in range(world - 1):
            assert plans[r]["halo_above"] == [plans[r + 1]["own_lo"], plans[r + 1]["own_lo"] + 1]
            assert plans[r + 1]["halo_below"] == list(range(plans[r]["own_hi"] - depth, plans[r]["own_hi"]))
        assert sum(p["own_hi"] - p["own_lo"] for p in plans) == X
        assert [p["emit_hi"] - p["emit_lo"] for p in plans] == [p["c_hi"] - p["c_lo"] for p in plans]


def _halo_worker(rank, world, port, shape):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = fields.noise(shape, 11)
        for depth in (1, 2):
            plan = idist.slab_plan(shape[0], rank, world, halo_below=depth)
            plan["depth_up"] = depth
            ext = torch.full((plan["n_ext"], shape[1], shape[2]), float("nan"))
            ext[plan["own_lo"] - plan["ext_lo"]:plan["own_hi"] - plan["ext_lo"]] = full[plan["own_lo"]:plan["own_hi"]]
            idist.exchange_halos(ext, plan, rank, world)
            assert torch.equal(ext, full[plan["ext_lo"]:plan["ext_hi"] + 1]), f"rank {rank}: halo mismatch (depth {depth})"
        vb, fb, totals, allc = idist.global_bases(10 + rank, 100 + rank, torch.device("cpu"))
        assert vb == sum(10 + r for r in range(rank)) and fb == sum(100 + r for r in range(rank))
        assert totals == (sum(10 + r for r in range(world)), sum(100 + r for r in range(world)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_count_allgather_gloo(world):
    mp.spawn(_halo_worker, args=(world, 29500 + world, (13, 5, 6)), nprocs=world, join=True)


def px(i, X, lo=-1.0, hi=1.0):
    return float(_lib.lib().isoext_axis_position(i, X, lo, hi))


def simulate_rank(vals, rank, world, method, cuts=None):
    """What one rank computes, with the oracle standing in for the CUDA pipeline."""
    X = vals.shape[0]
    p = idist.slab_plan(X, rank, world, cuts)
    v_ext, f_ext, _ = oracle.mc_dense(vals, 0.0, method, x_range=(p["ext_lo"], p["ext_hi"]))
    n_below = len(oracle.mc_dense(vals, 0.0, method, x_range=(p["ext_lo"], p["c_lo"]))[1])
    n_own = len(oracle.mc_dense(vals, 0.0, method, x_range=(p["c_lo"], p["c_hi"]))[1])
    f_own = f_ext[n_below:n_below + n_own]
    b_lo = px(p["c_lo"], X) if rank > 0 else -np.inf
    b_hi = px(p["c_hi"], X) if rank < world - 1 else np.inf
    n_lo = int((v_ext[:, 0] < np.float32(b_lo)).sum()) if len(v_ext) else 0
    n_hi = int((v_ext[:, 0] < np.float32(b_hi)).sum()) if len(v_ext) else 0
    return v_ext[n_lo:n_hi], f_own, n_lo, n_hi


FIELDS = {
    "cuboid33_faces_on_slab_planes": lambda: fields.eval_field(S.CuboidSDF([1, 1, 1]), (33, 33, 33)),
    "sphere40": lambda: fields.eval_field(S.SphereSDF(0.6), (40, 24, 28)),
    "noise24": lambda: fields.noise((24, 10, 12), 3),
    "csg36": lambda: fields.eval_field(fields.csg_box_minus_sphere(), (36, 36, 36)),
    "planes_exactly_at_level": lambda: _plane_field(),
}


def _plane_field():
    # f = x - px[k] for several k: whole grid planes sit exactly on the level set, including slab boundaries
    X = 25
    ax = fields.axis(X)
    f = (ax[:, None, None] - ax[12]).expand(X, 9, 11).clone()
    f[:, 4:, :] = (ax[:, None, None] - ax[8]).expand(X, 5, 11)
    return f


def _concat_check(vals, world, method, cuts=None):
    gv, gf, _ = oracle.mc_dense(vals, 0.0, method)
    parts = [simulate_rank(vals, r, world, method, cuts) for r in range(world)]
    owned = [len(p[0]) for p in parts]
    bases = np.concatenate([[0], np.cumsum(owned)])
    vs, fs = [], []
    for r, (v_own, f_own, n_lo, n_hi) in enumerate(parts):
        f = idist.relabel_ids_torch(torch.from_numpy(f_own), n_lo, n_hi, int(bases[r]), int(bases[r + 1])).numpy()
        vs.append(v_own); fs.append(f)
    v = np.concatenate(vs) if vs else np.zeros((0, 3), np.float32)
    f = np.concatenate(fs)
    assert v.shape == gv.shape and np.array_equal(v.view(np.uint32), gv.view(np.uint32))
    assert np.array_equal(f, gf)


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_per_rank_parts_concatenate_to_the_single_device_mesh(name, world, method):
    _concat_check(FIELDS[name]().numpy(), world, method)


@pytest.mark.parametrize("name,cuts", [("cuboid33_faces_on_slab_planes", [0, 2, 16, 18, 32]), ("csg36", [0, 9, 11, 35]),
                                       ("planes_exactly_at_level", [0, 8, 12, 14, 24]), ("noise24", [0, 2, 4, 6, 21, 23])])
def test_uneven_slab_boundaries_give_the_same_mesh(name, cuts):
    """Explicit (load-balancing) cuts: the mesh must not depend on where the slabs are cut, including cuts right
    on planes that carry box faces / exact-level planes and minimal 2-layer slabs."""
    _concat_check(FIELDS[name]().numpy(), len(cuts) - 1, "nagae", cuts)


def test_balanced_cuts_properties():
    import random
    rng = random.Random(1)
    assert idist.balanced_cuts([1.0] * 64, 8) == [0, 8, 16, 24, 32, 40, 48, 56, 64]
    for _ in range(300):
        layers = rng.randint(4, 80)
        world = rng.randint(1, layers // 2)
        cost = [rng.random() ** 4 * rng.choice([1, 100]) for _ in range(layers)]
        cuts = idist.balanced_cuts(cost, world)
        assert idist.check_cuts(layers + 1, world, cuts) == cuts
    # a lump (an axis-aligned face) is isolated: the slab holding it is much thinner than the others
    cost = [1.0] * 100
    cost[30] = 60.0
    cuts = idist.balanced_cuts(cost, 4)
    width = [b - a for a, b in zip(cuts[:-1], cuts[1:])]
    r = next(i for i in range(4) if cuts[i] <= 30 < cuts[i + 1])
    assert width[r] == min(width) and max(sum(cost[a:b]) for a, b in zip(cuts[:-1], cuts[1:])) < 0.5 * sum(cost)
    with pytest.raises(RuntimeError):
        idist.balanced_cuts([1.0] * 5, 3)
    with pytest.raises(RuntimeError):
        idist.check_cuts(10, 2, [0, 1, 9])


def test_vertex_layer_histogram_counts_every_vertex_once():
    v = torch.tensor([[-1.0, 0, 0], [-0.01, 0, 0], [0.0, 0, 0], [0.99, 0, 0], [1.0, 0, 0]])
    h = idist.vertex_layer_histogram(v, 5, -1.0, 1.0)
    assert h.tolist() == [1.0, 1.0, 1.0, 2.0]


# ---- dual contouring on slabs: the ownership / numbering design, proved with the oracle ----------------------------
def _dc_rank_view(vals, its, d, rank, world, halo_below=2):
    """What rank `rank` would compute for dual contouring (dist.dc_slab_plan), derived from the oracle's global
    per-cell dual vertices and quads: returns (global ids its faces must get, ids it computes locally, n_own)."""
    X, Y, Z = vals.shape
    p = idist.dc_slab_plan(X, rank, world)
    ext_lo = max(0, p["c_lo"] - halo_below)
    layer = its.cell_coords[:, 0]                                   # x layer of every active cell
    edge_x = d["quad_edge"][:, 0] // (Y * Z)                        # plane of the lower end point of every quad's edge
    edge_x_hi = d["quad_edge"][:, 1] // (Y * Z)
    # quads the rank can evaluate: both end points inside its point planes
    can_eval = (edge_x >= ext_lo) & (edge_x_hi <= p["ext_hi"])
    used = np.zeros(len(layer), bool)
    used[d["quads"][can_eval].ravel()] = True
    visible = (layer >= ext_lo) & (layer <= p["layer_hi"]) & used
    pos = np.ascontiguousarray(d["dual_v"][visible].astype(np.float32))   # the welded vertices are the float32 dual vertices
    key = np.unique(pos.view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel())   # welded, lexicographic (x, y, z)
    local_v = np.stack([key["x"], key["y"], key["z"]], axis=1)
    b_lo = np.float32(px(p["c_lo"], X)) if rank > 0 else np.float32(-np.inf)
    b_hi = np.float32(px(p["c_hi"], X)) if rank < world - 1 else np.float32(np.inf)
    n_lo, n_hi = int((local_v[:, 0] < b_lo).sum()), int((local_v[:, 0] < b_hi).sum())
    # faces of the quads this rank emits, as positions -> local ids
    mine = (edge_x >= p["emit_lo"]) & (edge_x < p["emit_hi"])
    tri = np.repeat(mine, 2)
    want = d["f"][tri]                                               # global ids (ground truth)
    fpos = np.ascontiguousarray(d["v"][want.ravel()]).view([("x", "f4"), ("y", "f4"), ("z", "f4")]).ravel()
    local_ids = np.searchsorted(key, fpos)
    assert np.array_equal(key[np.minimum(local_ids, len(key) - 1)], fpos), "a referenced vertex is not in the rank's list"
    return want.ravel(), local_ids, n_lo, n_hi


def _dc_numbering_ok(vals, world, halo_below):
    its = oracle.get_intersection(vals, compute_normals=True)
    d = oracle.dual_contouring(its, vals.shape)
    try:
        views = [_dc_rank_view(vals, its, d, r, world, halo_below) for r in range(world)]
    except AssertionError:
        return False
    owned = [n_hi - n_lo for (_, _, n_lo, n_hi) in views]
    if sum(owned) != len(d["v"]):
        return False
    bases = np.concatenate([[0], np.cumsum(owned)])
    for r, (want, local_ids, n_lo, n_hi) in enumerate(views):
        got = idist.relabel_ids_torch(torch.from_numpy(local_ids.astype(np.int64)), n_lo, n_hi, int(bases[r]), int(bases[r + 1])).numpy()
        if not np.array_equal(got, want):
            return False
    return True


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_dc_slab_numbering_design(name, world):
    """Dual contouring on slabs (host-side design; the CUDA side is the next step): with TWO halo planes
    on both sides every rank can number the vertices its faces reference exactly as the single-device result does,
    and the owned ranges tile the global vertex list."""
    assert _dc_numbering_ok(FIELDS[name]().numpy(), world, halo_below=2)


def test_dc_slab_numbering_needs_the_second_halo_plane():
    """With ONE plane below (the marching-cubes halo) the numbering breaks as soon as dual vertices of layer c_r - 2
    are clipped onto plane c_r - 1 and interleave with those of layer c_r - 1 (noise does it)."""
    vals = FIELDS["noise24"]().numpy()
    assert not _dc_numbering_ok(vals, 2, halo_below=1) and _dc_numbering_ok(vals, 2, halo_below=2)


# ---- SparseGrid on slabs: partition of the sorted cell list, proved with the oracle -----------------------------------
def _sparse_band(vals):
    """Crossing cells of a dense field as a sparse grid: sorted cell ids + (N, 8) corner values (Morton corner order)."""
    X, Y, Z = vals.shape
    neg = vals < 0
    cnt = sum(neg[dx:dx + X - 1, dy:dy + Y - 1, dz:dz + Z - 1].astype(np.int8) for dx in (0, 1) for dy in (0, 1) for dz in (0, 1))
    x, y, z = np.nonzero((cnt > 0) & (cnt < 8))
    cells = (x * (Y - 1) + y) * (Z - 1) + z
    order = np.argsort(cells)
    x, y, z, cells = x[order], y[order], z[order], cells[order]
    v8 = np.stack([vals[x + (i >> 2 & 1), y + (i >> 1 & 1), z + (i & 1)] for i in range(8)], axis=1).astype(np.float32)
    return cells.astype(np.int64), v8


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["cuboid33_faces_on_slab_planes", "sphere40", "csg36", "noise24", "planes_exactly_at_level"])
def test_sparse_slabs_concatenate_to_the_single_device_mesh(name, world):
    """SparseGrid marching cubes on slabs (host-side design, dist.sparse_slab_select): every rank runs the ordinary
    sparse extraction on its owned cells plus one ghost layer on either side, keeps the faces of its owned cells,
    owns the vertices between its two threshold planes and relabels exactly like the dense path."""
    vals = FIELDS[name]().numpy()
    shape = vals.shape
    cells, v8 = _sparse_band(vals)
    gv, gf, _ = oracle.mc_sparse(v8, cells, shape)
    parts = []
    for r in range(world):
        ext, owned = (m.numpy() for m in idist.sparse_slab_select(cells, shape, r, world))
        c = idist.partition_cells(shape[0], world)
        below = ext & ~owned & (cells // ((shape[1] - 1) * (shape[2] - 1)) < c[r])
        v_ext, f_ext, _ = oracle.mc_sparse(v8[ext], cells[ext], shape)
        n_below = len(oracle.mc_sparse(v8[below], cells[below], shape)[1]) if below.any() else 0
        n_own = len(oracle.mc_sparse(v8[owned], cells[owned], shape)[1]) if owned.any() else 0
        b_lo = np.float32(px(c[r], shape[0])) if r > 0 else np.float32(-np.inf)
        b_hi = np.float32(px(c[r + 1], shape[0])) if r < world - 1 else np.float32(np.inf)
        n_lo = int((v_ext[:, 0] < b_lo).sum()) if len(v_ext) else 0
        n_hi = int((v_ext[:, 0] < b_hi).sum()) if len(v_ext) else 0
        parts.append((v_ext[n_lo:n_hi], f_ext[n_below:n_below + n_own], n_lo, n_hi))
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    vs = [p[0] for p in parts]
    fs = [idist.relabel_ids_torch(torch.from_numpy(p[1]), p[2], p[3], int(bases[r]), int(bases[r + 1])).numpy() for r, p in enumerate(parts)]
    v, f = np.concatenate(vs), np.concatenate(fs)
    assert v.shape == gv.shape and np.array_equal(v.view(np.uint32), gv.view(np.uint32))
    assert np.array_equal(f, gf)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_sparse_slab_select_ghost_depths(world):
    """Host logic of the sparse slabs: marching cubes keeps one ghost layer on either side, dual contouring three below
    and two above (dist.sparse_slab_select); the owned cells of all ranks partition the list."""
    shape = (33, 9, 11)
    rng = np.random.default_rng(7)
    cells = np.sort(rng.choice(32 * 8 * 10, size=900, replace=False)).astype(np.int64)
    layer = cells // (8 * 10)
    c = idist.partition_cells(shape[0], world)
    owned_total = np.zeros(len(cells), dtype=int)
    for r in range(world):
        for below, above in ((1, 1), (3, 2)):
            ext, owned = (m.numpy() for m in idist.sparse_slab_select(cells, shape, r, world, below=below, above=above))
            assert np.array_equal(owned, (layer >= c[r]) & (layer < c[r + 1]))
            assert np.array_equal(ext, (layer >= c[r] - below) & (layer < c[r + 1] + above))
            assert not (owned & ~ext).any()
        owned_total += owned
    assert (owned_total == 1).all()
