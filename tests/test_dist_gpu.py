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


def _nccl_worker(rank, world, port, transport, field_name):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), ISOEXT_B200_PEER="1" if transport == "peer" else "0")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import isoext_b200 as iso
        from isoext_b200 import dist as idist
        vals = FIELDS[field_name]().cuda()
        sg = idist.SlabGrid(list(vals.shape))
        assert (sg._peer is not None) == (transport == "peer")
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


@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("field_name", ["cuboid65_faces_on_slab_planes", "torus_96x64x128"])
def test_real_ranks(field_name, transport):
    """Real ranks, one process per GPU: NVLink peer transport (csrc/peer.cu) and the NCCL fallback."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(n, 4)
    mp.spawn(_nccl_worker, args=(world, 29610 + (1 if transport == "peer" else 0), transport, field_name), nprocs=world, join=True)
