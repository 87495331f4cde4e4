"""GPU tests of slab-sharded marching cubes.

* simulated ranks on ONE GPU (always runs): every rank's slab is extracted in turn with the real CUDA
  kernels (x_offset / emit range / position thresholds), parts are relabelled with the CUDA kernel and
  concatenated: must equal the single-GPU mesh bit for bit;
* real ranks, one process per GPU (needs >= 2 GPUs, else skipped): the NVLink peer transport (halo pull and
  count exchange as kernels over IPC-mapped peer memory) and the NCCL fallback (send/recv + all_gather)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fields
from isoext_b200 import sdf as S

pytestmark = pytest.mark.gpu

FIELDS = {
    "cuboid65_faces_on_slab_planes": lambda: fields.eval_field(S.CuboidSDF([1, 1, 1]), (65, 65, 65)),
    "torus_96x64x128": lambda: fields.eval_field(fields.torus(), (96, 64, 128)),
    "noise_40x20x24": lambda: fields.noise((40, 20, 24), 3),
    "csg72": lambda: fields.eval_field(fields.csg_box_minus_sphere(), (72, 72, 72)),
}


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_simulated_ranks_concatenate_to_single_gpu_mesh(iso, name, world, method):
    from isoext_b200 import dist as idist
    vals = FIELDS[name]().cuda()
    g = iso.UniformGrid(list(vals.shape))
    g.set_values(vals)
    gv, gf = iso.marching_cubes(g, 0.0, method)
    parts = []
    for r in range(world):
        sg = idist.SlabGrid(list(vals.shape), rank=r, world=world)
        p = sg.plan
        sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])     # stands in for set_owned_values + halo exchange
        parts.append(idist.marching_cubes_local(sg, 0.0, method))
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    vs, fs = [], []
    for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
        vs.append(v_own); fs.append(f)
    v, f = torch.cat(vs), torch.cat(fs)
    assert v.shape == gv.shape and torch.equal(v.view(torch.int32), gv.view(torch.int32))
    assert torch.equal(f, gf)


@pytest.mark.parametrize("name,cuts", [("cuboid65_faces_on_slab_planes", [0, 2, 32, 34, 48, 64]), ("csg72", [0, 30, 32, 71]),
                                       ("noise_40x20x24", [0, 2, 4, 37, 39])])
def test_simulated_ranks_with_uneven_cuts(iso, name, cuts):
    """Load-balancing cuts (dist.balanced_cuts) only move work: same mesh bit for bit."""
    from isoext_b200 import dist as idist
    vals = FIELDS[name]().cuda()
    g = iso.UniformGrid(list(vals.shape))
    g.set_values(vals)
    gv, gf = iso.marching_cubes(g)
    world, parts = len(cuts) - 1, []
    for r in range(world):
        sg = idist.SlabGrid(list(vals.shape), rank=r, world=world, cuts=cuts)
        p = sg.plan
        sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])
        parts.append(idist.marching_cubes_local(sg))
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
    v, f = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    assert torch.equal(v.view(torch.int32), gv.view(torch.int32)) and torch.equal(f, gf)


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_dual_contouring_simulated_ranks_concatenate_to_single_gpu_mesh(iso, name, world):
    """DC on slabs, CUDA side (dist.dual_contouring_local on SlabGrid(dc=True): two halo planes below, dual vertices of
    the ghost layers welded too, quads of the owned planes, ownership by position): the per-rank parts, relabelled
    by the CUDA kernel, concatenate to the single-GPU dual-contouring mesh bit for bit (V bits and F)."""
    from isoext_b200 import dist as idist
    vals = FIELDS[name]().cuda()
    g = iso.UniformGrid(list(vals.shape))
    g.set_values(vals)
    gv, gf = iso.dual_contouring(g)
    parts = []
    for r in range(world):
        sg = idist.SlabGrid(list(vals.shape), rank=r, world=world, dc=True)
        p = sg.plan
        assert p["ext_lo"] == max(0, p["c_lo"] - 2)
        sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])
        parts.append(idist.dual_contouring_local(sg))
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
    v, f = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    assert v.shape == gv.shape and torch.equal(v.view(torch.int32), gv.view(torch.int32))
    assert torch.equal(f, gf)
    # the same slabs serve marching cubes (one more ghost layer below changes nothing)
    mv, mf = iso.marching_cubes(g)
    mparts = []
    for r in range(world):
        sg = idist.SlabGrid(list(vals.shape), rank=r, world=world, dc=True)
        p = sg.plan
        sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])
        mparts.append(idist.marching_cubes_local(sg))
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in mparts])])
    for r, (v_own, f, n_lo, n_hi) in enumerate(mparts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
    assert torch.equal(torch.cat([p[0] for p in mparts]).view(torch.int32), mv.view(torch.int32))
    assert torch.equal(torch.cat([p[1] for p in mparts]), mf)


def test_dual_contouring_slabs_with_uneven_cuts_and_level(iso):
    from isoext_b200 import dist as idist
    vals = FIELDS["csg72"]().cuda()
    g = iso.UniformGrid(list(vals.shape))
    g.set_values(vals)
    for level, cuts in ((0.0, [0, 30, 32, 71]), (0.02, [0, 2, 4, 60, 71]), (-0.01, [0, 35, 37, 39, 71])):
        gv, gf = iso.dual_contouring(g, level)
        world, parts = len(cuts) - 1, []
        for r in range(world):
            sg = idist.SlabGrid(list(vals.shape), rank=r, world=world, cuts=cuts, dc=True)
            p = sg.plan
            sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])
            parts.append(idist.dual_contouring_local(sg, level))
        bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
        for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
            idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
        assert torch.equal(torch.cat([p[0] for p in parts]).view(torch.int32), gv.view(torch.int32))
        assert torch.equal(torch.cat([p[1] for p in parts]), gf)


def _band_grid(iso, vals):
    """Sparse narrow band of a dense field: every crossing cell, with its own 8 corner values."""
    import oracle  # noqa: F401  (test infrastructure only)
    from test_dist_cpu import _sparse_band
    cells, v8 = _sparse_band(vals.cpu().numpy())
    g = iso.SparseGrid(list(vals.shape))
    g.add_cells(torch.from_numpy(cells).to(torch.int32).cuda())
    g.set_values(torch.from_numpy(v8).cuda())
    return g


@pytest.mark.parametrize("method", ["nagae", "lorensen"])
@pytest.mark.parametrize("world", [2, 3, 6])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_sparse_grid_simulated_slabs_concatenate_to_single_device_mesh(iso, name, world, method):
    """SparseGrid on slabs, CUDA side (dist.SparseSlab: owned cells + one ghost layer either side, emit range over the
    cell list, ownership by position): parts relabelled by the CUDA kernel concatenate to the single-device sparse
    mesh bit for bit."""
    from isoext_b200 import dist as idist
    vals = FIELDS[name]()
    g = _band_grid(iso, vals)
    gv, gf = iso.marching_cubes(g, 0.0, method)
    parts = [idist.marching_cubes_sparse_local(idist.SparseSlab(g, rank=r, world=world), 0.0, method) for r in range(world)]
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
    v, f = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    assert v.shape == gv.shape and torch.equal(v.view(torch.int32), gv.view(torch.int32))
    assert torch.equal(f, gf)


@pytest.mark.parametrize("holes", [False, True])
@pytest.mark.parametrize("world", [2, 3, 6])
@pytest.mark.parametrize("name", sorted(FIELDS))
def test_sparse_dual_contouring_simulated_slabs_concatenate_to_single_device_mesh(iso, name, world, holes):
    """Sparse dual contouring on slabs (dist.SparseSlab(dc=True): 3 ghost layers below, 2 above; every quad of the local
    list marks its cells as used, only owned cells emit; ownership by position): the parts concatenate to the
    single-device sparse DC mesh bit for bit.  ``holes``: a third of the band's cells is removed at random, so whether a
    dual vertex exists at all depends on which neighbour cells are present -- also across the slab boundaries."""
    from isoext_b200 import dist as idist
    vals = FIELDS[name]()
    g = _band_grid(iso, vals)
    if holes:
        gen = torch.Generator().manual_seed(3)
        drop = g.get_cell_indices()[(torch.rand(g.get_num_cells(), generator=gen) < 0.33).cuda()]
        keep_vals = g.get_values()
        keep_mask = ~torch.isin(g.get_cell_indices(), drop)
        g.remove_cells(drop)
        g.set_values(keep_vals[keep_mask].contiguous())
    gv, gf = iso.dual_contouring(g)
    parts = [idist.dual_contouring_sparse_local(idist.SparseSlab(g, rank=r, world=world, dc=True)) for r in range(world)]
    if gv is None:
        assert all(len(p[1]) == 0 for p in parts)
        return
    bases = np.concatenate([[0], np.cumsum([len(p[0]) for p in parts])])
    for r, (v_own, f, n_lo, n_hi) in enumerate(parts):
        idist.relabel_faces_(f, n_lo, n_hi, int(bases[r]), int(bases[r + 1]))
    v, f = torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
    assert v.shape == gv.shape and torch.equal(v.view(torch.int32), gv.view(torch.int32))
    assert torch.equal(f, gf)


def test_c3_2048_csg_two_slabs_equal_single_gpu():
    """BASELINE.json configs[2] at FULL size on one GPU: the 2048^3 CSG field extracted whole and as two simulated
    slabs with load-balanced cuts must agree bit for bit; the mesh is a closed genus-0 surface whose x-sorted
    vertices all lie on the box or the sphere.  (The reference cannot represent this grid at all.)"""
    import isoext_b200 as iso
    from isoext_b200 import dist as idist
    if torch.cuda.mem_get_info()[0] < 110e9:
        pytest.skip("needs ~100 GB of free device memory")
    n = 2048
    fn = fields.csg_box_minus_sphere()
    g = iso.UniformGrid([n] * 3)
    ax = fields.axis(n).cuda()
    vals = g.values_view()
    for x0 in range(0, n, 16):
        P = torch.stack(torch.meshgrid(ax[x0:x0 + 16], ax, ax, indexing="ij"), dim=-1)
        vals[x0:x0 + 16] = fn(P)
        del P
    v, f = iso.marching_cubes(g)
    # x-major sorted, strictly increasing lexicographically
    a, b = v[:-1], v[1:]
    lt = (a[:, 0] < b[:, 0]) | ((a[:, 0] == b[:, 0]) & ((a[:, 1] < b[:, 1]) | ((a[:, 1] == b[:, 1]) & (a[:, 2] < b[:, 2]))))
    assert bool(lt.all())
    del a, b, lt
    # closed 2-manifold of genus 0: every edge in exactly two triangles, V - E + F = 2
    f64 = f.long()
    e = torch.cat([f64[:, [0, 1]], f64[:, [1, 2]], f64[:, [2, 0]]])
    key = torch.minimum(e[:, 0], e[:, 1]) * len(v) + torch.maximum(e[:, 0], e[:, 1])
    del e, f64
    uniq, cnt = torch.unique(key, return_counts=True)
    assert bool((cnt == 2).all()) and len(v) - len(uniq) + len(f) == 2
    del key, uniq, cnt
    # every vertex lies on the surface (linear interpolation of an exact SDF: within a fraction of a cell)
    worst = 0.0
    for i in range(0, len(v), 1 << 22):
        worst = max(worst, float(fn(v[i:i + (1 << 22)]).abs().max()))
    assert worst < 0.5 * 2.0 / (n - 1)
    # the CPU oracle on windows of x layers of the full-size grid: a box face (1.5 M vertices in one layer), box + sphere,
    # the far box face -- bit-equal triangles
    import fullsize
    from isoext_b200 import _lib
    checked = sum(fullsize.check_layers_vs_oracle(_lib.lib(), vals, v, f, a, b) for a, b in ((408, 411), (1300, 1302), (1636, 1639)))
    assert checked > 1000000
    # two slabs with cuts balanced by the measured load
    hist = idist.vertex_layer_histogram(v, n, -1.0, 1.0)
    cuts = idist.balanced_cuts((2.6e-6 + 0.43e-9 * hist).tolist(), 2, ghost_cost=(0.43e-9 * hist).tolist())
    parts = []
    for r in range(2):
        sg = idist.SlabGrid([n] * 3, rank=r, world=2, cuts=cuts)
        p = sg.plan
        sg._ext.copy_(vals[p["ext_lo"]:p["ext_hi"] + 1])
        parts.append(idist.marching_cubes_local(sg))
        del sg
    idist.relabel_faces_(parts[0][1], parts[0][2], parts[0][3], 0, len(parts[0][0]))
    idist.relabel_faces_(parts[1][1], parts[1][2], parts[1][3], len(parts[0][0]), len(parts[0][0]) + len(parts[1][0]))
    assert torch.equal(torch.cat([parts[0][0], parts[1][0]]).view(torch.int32), v.view(torch.int32))
    assert torch.equal(torch.cat([parts[0][1], parts[1][1]]), f)


def _nccl_worker(rank, world, port, transport, field_name):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ISOEXT_B200_PEER="0" if transport == "nccl" else "1")
    if transport == "peer_fails":      # one rank cannot map its peers: ALL ranks must fall back to NCCL (no hang, same mesh)
        os.environ["ISOEXT_B200_PEER_TEST_FAIL_RANK"] = str(world - 1)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import isoext_b200 as iso
        from isoext_b200 import dist as idist
        vals = FIELDS[field_name]().cuda()
        sg = idist.SlabGrid(list(vals.shape))
        assert (sg._peer is not None) == (transport == "peer")       # "peer_fails": every rank fell back together
        lo, hi = sg.owned_point_range()
        g = iso.UniformGrid(list(vals.shape))
        # several steps with changing values: epochs, the overwrite guard and the count ring all advance
        for step, (scale, level) in enumerate([(1.0, 0.0), (-1.0, 0.0), (1.0, 0.03), (0.5, 0.01), (1.0, 5.0)]):
            cur = (vals * scale).contiguous()
            sg.set_owned_values(cur[lo:hi].contiguous())
            v_own, f_own = idist.marching_cubes(sg, level)
            v, f = idist.gather_mesh(v_own, f_own)
            g.set_values(cur)
            gv, gf = iso.marching_cubes(g, level)
            if gv is None:
                assert len(v) == 0 and len(f) == 0, f"rank {rank} step {step}: expected an empty mesh"
                continue
            assert torch.equal(v.view(torch.int32), gv.view(torch.int32)) and torch.equal(f, gf), f"rank {rank} step {step}: mismatch"
        # dual contouring on slabs that fetch two planes from below (peer pull of 2 + 2 planes / NCCL send-recv)
        sgd = idist.SlabGrid(list(vals.shape), dc=True)
        sgd.set_owned_values(vals[lo:hi].contiguous())
        g.set_values(vals)
        for level in (0.0, 0.02):
            dv, df = idist.gather_mesh(*idist.dual_contouring(sgd, level))
            gdv, gdf = iso.dual_contouring(g, level)
            assert torch.equal(dv.view(torch.int32), gdv.view(torch.int32)) and torch.equal(df, gdf), f"rank {rank}: DC slabs level {level}"
            mv, mf = idist.gather_mesh(*idist.marching_cubes(sgd, level))
            gmv, gmf = iso.marching_cubes(g, level)
            assert torch.equal(mv.view(torch.int32), gmv.view(torch.int32)) and torch.equal(mf, gmf), f"rank {rank}: MC on dc slabs"
        sgd.close()
        # SparseGrid on slabs: every rank keeps its part of the band (+ ghost cells); only two counts cross ranks
        band = _band_grid(iso, vals)
        sv, sf = idist.gather_mesh(*idist.marching_cubes_sparse(idist.SparseSlab(band)))
        gsv, gsf = iso.marching_cubes(band)
        assert torch.equal(sv.view(torch.int32), gsv.view(torch.int32)) and torch.equal(sf, gsf), f"rank {rank}: sparse slabs"
        dv, df = idist.gather_mesh(*idist.dual_contouring_sparse(idist.SparseSlab(band, dc=True)))
        gdv, gdf = iso.dual_contouring(band)
        assert torch.equal(dv.view(torch.int32), gdv.view(torch.int32)) and torch.equal(df, gdf), f"rank {rank}: sparse DC slabs"
        # same values again without a new set_owned_values, and the explicit no-exchange variant
        v_own, f_own = idist.marching_cubes(sg, 0.0)
        v2_own, f2_own = idist.marching_cubes(sg, 0.0, exchange=False)
        assert torch.equal(v_own, v2_own) and torch.equal(f_own, f2_own)
        # re-cut the slabs by measured load: the global mesh must not change
        hist = idist.vertex_layer_histogram(v_own, vals.shape[0], -1.0, 1.0)
        cuts = idist.balanced_cuts((hist + 1.0).tolist(), world, ghost_cost=hist.tolist())
        sg.close()
        sg = idist.SlabGrid(list(vals.shape), cuts=cuts)
        lo, hi = sg.owned_point_range()
        sg.set_owned_values(vals[lo:hi].contiguous())
        v, f = idist.gather_mesh(*idist.marching_cubes(sg))
        g.set_values(vals)
        gv, gf = iso.marching_cubes(g)
        assert torch.equal(v.view(torch.int32), gv.view(torch.int32)) and torch.equal(f, gf), f"rank {rank}: balanced cuts {cuts}"
        sg.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl", "peer_fails"])
@pytest.mark.parametrize("field_name", ["cuboid65_faces_on_slab_planes", "torus_96x64x128"])
def test_real_ranks(field_name, transport):
    """Real ranks, one process per GPU: NVLink peer transport (csrc/peer.cu) and the NCCL fallback."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    mp.spawn(_nccl_worker, args=(world, 29610 + {"peer": 1, "nccl": 0, "peer_fails": 2}[transport], transport, field_name), nprocs=world, join=True)
