"""Slab-sharded multi-GPU extraction (new capability: the reference is single-GPU).

One process per GPU (``torch.distributed``, NCCL over NVLink; gloo works for the host logic on CPU).
The global (X, Y, Z) grid is cut along dim 0 (the slowest axis, so slabs are contiguous):

  * rank r owns cell layers ``[c_r, c_{r+1})`` and stores point planes ``[c_r, c_{r+1})`` (+ the last
    plane on the last rank);
  * for extraction it works on the EXTENDED slab of point planes ``[c_r - 1, c_{r+1} + 1]``: one halo
    plane below, two above (P2P send/recv of (Y, Z) planes).  The ghost cell layers on either side are
    evaluated (they decide which vertices on the shared planes exist) but emit no faces;
  * the welded, position-sorted vertex list of the extended slab is split by the *position* thresholds
    ``px[c_r]`` and ``px[c_{r+1}]``: a vertex with ``px[c_r] <= x < px[c_{r+1}]`` is owned by rank r.
    Because both neighbours evaluate the same two cell layers around a shared plane with bit-identical
    arithmetic, they agree on every vertex near it, so a face that references a vertex owned by a
    neighbour can compute that vertex's global id locally from the all-gathered counts;
  * per-rank vertex and face lists concatenate (in rank order) to exactly the single-GPU result: V is
    sorted by x first and F is in ascending (x-major) cell order.
Only counts cross ranks in the extraction step.

Two transports.  On one node (the normal case: NVLink / NVSwitch) every rank exports its value slab and a
small SYNC block through CUDA IPC and the two exchange steps are kernels of the library over peer memory
(csrc/peer.cu): ``k_peer_pull`` waits for the neighbour's epoch flag and copies the halo planes straight out
of the neighbour's slab; ``k_relabel_peer`` reads the lower ranks' vertex counts from their SYNC blocks, sums
them and relabels the faces -- no host rendezvous, no NCCL call and no extra host synchronisation on the data
path.  Otherwise (``ISOEXT_B200_PEER=0``, several nodes, gloo on CPU) the same steps run as NCCL/gloo
``batch_isend_irecv`` of planes plus one ``all_gather`` of two int64 per rank.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import socket

import torch
import torch.distributed as dist

# ---- pure host logic (no CUDA): unit-tested on CPU with gloo ---------------------------------------


def partition_cells(X: int, world: int) -> list[int]:
    """Cell-layer boundaries c_0..c_world along dim 0 (X points -> X-1 cell layers), as even as possible."""
    layers = X - 1
    if layers < 2 * world:
        raise RuntimeError(f"cannot cut {layers} cell layers into {world} slabs of at least 2 layers")
    return [(layers * r) // world for r in range(world + 1)]


def check_cuts(X: int, world: int, cuts) -> list[int]:
    """Validate explicit cell-layer boundaries c_0..c_world (every slab needs at least 2 layers: the halo
    planes of a rank come from its immediate neighbours only)."""
    cuts = [int(c) for c in cuts]
    if len(cuts) != world + 1 or cuts[0] != 0 or cuts[-1] != X - 1:
        raise RuntimeError(f"cuts must be {world + 1} boundaries from 0 to {X - 1}")
    if any(b - a < 2 for a, b in zip(cuts[:-1], cuts[1:])):
        raise RuntimeError("every slab needs at least 2 cell layers")
    return cuts


def balanced_cuts(layer_cost, world: int, ghost_cost=None) -> list[int]:
    """Cell-layer boundaries c_0..c_world that minimise the most expensive slab (exact dynamic programme).

    ``layer_cost[l]``: cost of owning cell layer l (X-1 non-negative numbers).  ``ghost_cost[l]`` (optional): what a
    slab pays for layer l when l is one of its two ghost layers (the layer just below its first and the one just
    above its last: they are evaluated for welding but emit no faces) -- normally the surface part of
    ``layer_cost``.  Every slab gets at least 2 layers.  Any boundaries give the same global mesh bit for bit;
    these only move work, e.g. away from the ranks next to an axis-aligned face."""
    cost = torch.as_tensor([float(c) for c in layer_cost], dtype=torch.float64)
    L = int(cost.numel())
    if L < 2 * world:
        raise RuntimeError(f"cannot cut {L} cell layers into {world} slabs of at least 2 layers")
    ghost = torch.zeros(L, dtype=torch.float64) if ghost_cost is None else torch.as_tensor([float(c) for c in ghost_cost], dtype=torch.float64)
    if world == 1:
        return [0, L]
    step = max(1, (L + 4095) // 4096)                 # boundaries on a coarser lattice for very deep grids
    pos = list(range(0, L, step))
    if pos[-1] != L:
        pos.append(L)
    P = torch.tensor(pos)
    prefix = torch.cat([torch.zeros(1, dtype=torch.float64), cost.cumsum(0)])
    below = torch.cat([torch.zeros(1, dtype=torch.float64), ghost])        # below[a] = ghost cost of layer a-1
    above = torch.cat([ghost, torch.zeros(1, dtype=torch.float64)])        # above[b] = ghost cost of layer b
    # C[i, j] = cost of the slab [pos_i, pos_j)
    C = (prefix[P][None, :] - prefix[P][:, None]) + below[P][:, None] + above[P][None, :]
    C = torch.where((P[None, :] - P[:, None]) >= 2, C, torch.full_like(C, float("inf")))
    n = len(pos)
    dp = C[0].clone()                                  # one slab [0, pos_j)
    choice = []
    for _ in range(1, world):
        M = torch.maximum(dp[:, None], C)              # last slab [pos_i, pos_j) after an optimal prefix ending at pos_i
        dp, arg = M.min(dim=0)
        choice.append(arg)
    if not math.isfinite(float(dp[n - 1])):
        raise RuntimeError(f"cannot cut {L} cell layers into {world} slabs of at least 2 layers")
    cuts, j = [L], n - 1
    for arg in reversed(choice):
        j = int(arg[j])
        cuts.append(pos[j])
    return [0] + cuts[::-1]


def vertex_layer_histogram(v_own: torch.Tensor, X: int, aabb_min_x: float, aabb_max_x: float, group=None) -> torch.Tensor:
    """Vertices per cell layer of the GLOBAL mesh (sum over ranks of this rank's owned vertices): the measured
    surface load behind ``balanced_cuts``."""
    layers = X - 1
    if v_own.is_cuda:
        from . import _lib
        from .grid import _stream_ptr
        h32 = torch.empty(layers, dtype=torch.int32, device=v_own.device)
        v = v_own.contiguous()
        with torch.cuda.device(v_own.device):
            _lib.check(_lib.lib().isoext_vertex_layer_histogram(v.data_ptr(), v.shape[0], float(aabb_min_x), float(aabb_max_x), layers,
                                                                h32.data_ptr(), _stream_ptr()))
        h = h32.to(torch.float64)
    elif v_own.numel():      # host tensors (CPU unit tests of the balancer)
        t = (v_own[:, 0].double() - aabb_min_x) / (aabb_max_x - aabb_min_x) * layers
        h = torch.bincount(t.floor().clamp_(0, layers - 1).long(), minlength=layers).to(torch.float64)
    else:
        h = torch.zeros(layers, dtype=torch.float64, device=v_own.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(h, group=group)
    return h


def slab_plan(X: int, rank: int, world: int, cuts=None, halo_below: int = 1) -> dict:
    """Everything rank `rank` needs to know about its slab (all plane indices are GLOBAL).
    ``halo_below``: planes fetched from the lower neighbour -- 1 for marching cubes, 2 when the slab also serves
    dual contouring (dc_slab_plan explains why)."""
    c = partition_cells(X, world) if cuts is None else check_cuts(X, world, cuts)
    c_lo, c_hi = c[rank], c[rank + 1]
    own_lo, own_hi = c_lo, (c_hi if rank < world - 1 else X)        # owned point planes [own_lo, own_hi)
    ext_lo, ext_hi = max(0, c_lo - halo_below), min(X - 1, c_hi + 1)   # extended slab planes [ext_lo, ext_hi]
    return dict(c_lo=c_lo, c_hi=c_hi, own_lo=own_lo, own_hi=own_hi, ext_lo=ext_lo, ext_hi=ext_hi,
                n_ext=ext_hi - ext_lo + 1, emit_lo=c_lo - ext_lo, emit_hi=c_hi - ext_lo,
                halo_below=list(range(ext_lo, c_lo)) if rank > 0 else [],             # from rank-1 (its last planes)
                halo_above=([p for p in (c_hi, c_hi + 1) if p <= X - 1] if rank < world - 1 else []))


def dc_slab_plan(X: int, rank: int, world: int, cuts=None) -> dict:
    """Slab geometry for dual contouring (the CUDA side is dist.dual_contouring on a SlabGrid(dc=True)).

    Rank r emits the quads of the sign-change edges whose lower end point lies in planes ``[c_r, c_{r+1})`` and owns the
    welded vertices with ``px[c_r] <= x < px[c_{r+1}]``.  A quad joins cells of layers ``p_lo.x - 1`` and ``p_lo.x``, so
    faces reference dual vertices anywhere inside layer ``c_r - 1``; to number them from the top of the lower
    neighbour's range the rank must know EVERY global vertex between them and ``px[c_r]`` -- including vertices of
    layer ``c_r - 2`` that were clipped exactly onto plane ``c_r - 1``.  Hence two halo planes on both sides:
    point planes ``[c_r - 2, c_{r+1} + 1]``, dual vertices of cell layers ``[c_r - 2, c_{r+1}]``, usage marks from
    the quads of edges in planes ``[c_r - 2, c_{r+1} + 1]`` (checked on CPU in tests/test_dist_cpu.py)."""
    c = partition_cells(X, world) if cuts is None else check_cuts(X, world, cuts)
    c_lo, c_hi = c[rank], c[rank + 1]
    ext_lo, ext_hi = max(0, c_lo - 2), min(X - 1, c_hi + 1)
    return dict(c_lo=c_lo, c_hi=c_hi, ext_lo=ext_lo, ext_hi=ext_hi, layer_lo=ext_lo, layer_hi=min(X - 2, c_hi),
                emit_lo=c_lo, emit_hi=(c_hi if rank < world - 1 else X))


def sparse_slab_select(cell_idx, shape, rank: int, world: int, cuts=None, below: int = 1, above: int = 1):
    """Slab sharding of a SparseGrid (host logic; the sorted cell list is partitioned by x layer):
    returns boolean masks ``(ext, owned)`` over ``cell_idx``.  Rank r emits the faces of its ``owned`` cells (layers
    ``[c_r, c_{r+1})``) and welds the vertices of the ``ext`` cells (``below`` / ``above`` ghost layers: one on either
    side for marching cubes), so that -- as in the dense path -- it knows every vertex near its two threshold planes
    and can number by position.

    Dual contouring needs ``below=3, above=2``: its quads reach one layer down; the previous rank's vertices that
    interleave with those of layer c_r - 1 include dual vertices of layer c_r - 2 clipped onto plane c_r - 1, and
    whether such a vertex exists at all ("used by some quad") depends on which cells of layer c_r - 3 are present;
    likewise the next rank's first vertices (x == px[c_{r+1}]) come from layer c_{r+1} as well, whose usage depends on
    the presence of cells in layer c_{r+1} + 1."""
    X, Y, Z = (int(v) for v in shape)
    c = partition_cells(X, world) if cuts is None else check_cuts(X, world, cuts)
    layer = torch.as_tensor(cell_idx).to(torch.int64) // ((Y - 1) * (Z - 1))
    owned = (layer >= c[rank]) & (layer < c[rank + 1])
    ext = (layer >= c[rank] - below) & (layer <= c[rank + 1] - 1 + above)
    return ext, owned


def exchange_halos(ext: torch.Tensor, plan: dict, rank: int, world: int, group=None) -> None:
    """Fill the halo planes of the extended slab `ext` ((n_ext, Y, Z), owned planes already in place).

    Sends my first two owned planes down to rank-1 and my last owned plane(s) up to rank+1 (as many as the halo
    below is deep: every rank uses the same depth)."""
    if world == 1:
        return
    ops, keep = [], []
    base = plan["ext_lo"]
    depth = max(1, plan["c_lo"] - plan["ext_lo"]) if rank > 0 else None

    def plane(g):
        return ext[g - base]

    def peer(r):
        return dist.get_global_rank(group, r) if group is not None else r

    if rank > 0:
        for g in (plan["own_lo"], plan["own_lo"] + 1):                 # they are rank-1's halo_above
            t = plane(g).contiguous(); keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, peer(rank - 1), group))
        for g in plan["halo_below"]:
            ops.append(dist.P2POp(dist.irecv, plane(g), peer(rank - 1), group))
    if rank < world - 1:
        up = plan.get("depth_up", 1)                                     # rank+1's halo_below
        for g in range(plan["c_hi"] - up, plan["c_hi"]):
            t = plane(g).contiguous(); keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, peer(rank + 1), group))
        for g in plan["halo_above"]:
            ops.append(dist.P2POp(dist.irecv, plane(g), peer(rank + 1), group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()


def global_bases(n_own: int, n_tri: int, device, group=None):
    """All-gather (owned vertices, triangles) and return (vertex base, face base, totals, all counts)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = torch.tensor([n_own, n_tri], dtype=torch.int64, device=device)
    if world == 1:
        return 0, 0, (n_own, n_tri), mine[None]
    flat = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(flat, mine, group=group)
    rank = dist.get_rank(group)
    allc_h = flat.view(world, 2).cpu()
    vb = int(allc_h[:rank, 0].sum())
    fb = int(allc_h[:rank, 1].sum())
    return vb, fb, (int(allc_h[:, 0].sum()), int(allc_h[:, 1].sum())), allc_h


def relabel_ids_torch(F: torch.Tensor, n_lo: int, n_hi: int, base_mine: int, base_next: int) -> torch.Tensor:
    """Reference implementation (pure torch) of the id map the CUDA kernel isoext_relabel_faces applies."""
    ids = F.long()
    out = torch.where(ids < n_lo, base_mine - (n_lo - ids),
                      torch.where(ids < n_hi, base_mine + (ids - n_lo), base_next + (ids - n_hi)))
    return out.to(F.dtype)


# ---- GPU layer -------------------------------------------------------------------------------------
class SlabGrid:
    """This rank's slab of a global UniformGrid (same constructor arguments as ``UniformGrid``).

    ``cuts`` (optional, identical on all ranks): explicit cell-layer boundaries c_0..c_world instead of the even
    split, e.g. from ``balanced_cuts`` -- the extracted mesh does not depend on them."""

    def __init__(self, shape, aabb_min=(-1.0, -1.0, -1.0), aabb_max=(1.0, 1.0, 1.0), default_value=3.4028234663852886e38,
                 group=None, device=None, rank=None, world=None, cuts=None, dc=False):
        from . import _lib
        from .grid import _Workspace
        self.shape = tuple(int(s) for s in shape)
        self.aabb_min = tuple(float(v) for v in aabb_min)
        self.aabb_max = tuple(float(v) for v in aabb_max)
        self.group = group
        # rank/world may be given explicitly to simulate a slab without a process group (tests)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.cuts = None if cuts is None else check_cuts(self.shape[0], self.world, cuts)   # explicit slab boundaries
        self.halo_below = 2 if dc else 1          # dual contouring needs two planes from the lower neighbour
        self.plan = slab_plan(self.shape[0], self.rank, self.world, self.cuts, self.halo_below)
        self.plan["depth_up"] = self.halo_below
        _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        X, Y, Z = self.shape
        self._peer = None
        want_peer = (rank is None and world is None and self.world > 1 and dist.is_initialized() and self.device.type == "cuda"
                     and os.environ.get("ISOEXT_B200_PEER", "1") != "0")
        if want_peer:
            self._peer_setup((self.plan["n_ext"], Y, Z))
        if self._peer is None:
            self._ext = torch.empty((self.plan["n_ext"], Y, Z), dtype=torch.float32, device=self.device)
        self._ext.fill_(float(default_value))
        self._ws = _Workspace()
        self._cap_hint = 0
        self._hints = {}
        lib = _lib.lib()
        px = lambda i: float(lib.isoext_axis_position(i, X, self.aabb_min[0], self.aabb_max[0]))
        self.thresholds = (px(self.plan["c_lo"]) if self.rank > 0 else -math.inf,
                           px(self.plan["c_hi"]) if self.rank < self.world - 1 else math.inf)

    def owned_point_range(self):
        return self.plan["own_lo"], self.plan["own_hi"]

    def local_points(self) -> int:
        return int(self._ext.numel())

    def owned_values(self) -> torch.Tensor:
        """View of the owned planes.  Read it freely; to WRITE new values use ``set_owned_values`` (or call
        ``_guard_overwrite()`` first): with the peer transport the neighbours pull their halo planes straight out of
        this storage, and an in-place fill must not overtake the pull of the previous exchange."""
        p = self.plan
        return self._ext[p["own_lo"] - p["ext_lo"]:p["own_hi"] - p["ext_lo"]]

    def set_owned_values(self, values: torch.Tensor) -> None:
        own = self.owned_values()
        if tuple(values.shape) != tuple(own.shape):
            raise RuntimeError("Cannot set values with different shapes")
        self._guard_overwrite()
        own.copy_(values)

    def exchange_halos(self, overlap: bool = False):
        """Fill the halo planes.  ``overlap=True`` (peer transport only): the pull runs on a side stream and the call
        returns ``(planes_below, planes_above, event)`` for ``mc_dense_raw(halo=...)`` -- the volume stream over the owned
        planes then starts while the halo planes are still in flight; otherwise returns None (halo complete in stream
        order)."""
        if self._peer is not None:
            return self._peer_exchange(overlap)
        exchange_halos(self._ext, self.plan, self.rank, self.world, self.group)
        return None

    # ---- NVLink peer transport (csrc/peer.cu) -------------------------------------------------------
    def _peer_setup(self, ext_shape) -> None:
        """Collective.  Allocate the value slab and the SYNC block as IPC-exportable memory, exchange the handles and
        map the neighbours' slabs and everybody's SYNC block.  Leaves ``self._peer = None`` (NCCL transport) when the
        ranks are not all on one host."""
        from . import _lib
        lib = _lib.lib()
        n_floats = int(ext_shape[0]) * int(ext_shape[1]) * int(ext_shape[2])
        words = int(lib.isoext_peer_sync_words())
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=self.group)
        if len(set(hosts)) != 1:
            return
        # Every step that can fail on one rank only (allocation, IPC export, mapping a neighbour: GPUs without peer
        # access, MIG slices, containers that share a host name across nodes) is followed by an all-reduce of a success
        # flag, so that ALL ranks fall back to the NCCL transport together instead of one raising inside a collective.
        def all_ok(ok: bool) -> bool:
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            return bool(flag.item())

        ext_ptr, sync_ptr = C.c_void_p(), C.c_void_p()
        h_ext, h_sync = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
        peer_ext, peer_sync = {}, {}

        def release():
            for ptr in list(peer_ext.values()) + list(peer_sync.values()):
                lib.isoext_peer_close(ptr)
            for ptr in (ext_ptr, sync_ptr):
                if ptr.value:
                    lib.isoext_peer_free(ptr.value)

        with torch.cuda.device(self.device):
            ok = lib.isoext_peer_alloc(n_floats * 4, C.byref(ext_ptr), h_ext) == 0
            ok = ok and lib.isoext_peer_alloc(words * 8, C.byref(sync_ptr), h_sync) == 0
            fail_rank = os.environ.get("ISOEXT_B200_PEER_TEST_FAIL_RANK")      # test hook: this rank "cannot" map its peers
            if not all_ok(ok):
                release()
                return
            sync = _device_view(sync_ptr.value, words, torch.int64, self.device)
            sync.zero_()
            torch.cuda.synchronize()
            allh = [None] * self.world
            dist.all_gather_object(allh, (bytes(h_ext), bytes(h_sync)), group=self.group)
            ok = True
            for r in range(self.world):
                if r == self.rank or not ok:
                    continue
                p = C.c_void_p()
                ok = lib.isoext_peer_open((C.c_ubyte * 64).from_buffer_copy(allh[r][1]), C.byref(p)) == 0
                if ok:
                    peer_sync[r] = p.value
                if ok and abs(r - self.rank) == 1:
                    p = C.c_void_p()
                    ok = lib.isoext_peer_open((C.c_ubyte * 64).from_buffer_copy(allh[r][0]), C.byref(p)) == 0
                    if ok:
                        peer_ext[r] = p.value
            if fail_rank is not None and int(fail_rank) == self.rank:
                ok = False
            if not all_ok(ok):
                del sync
                torch.cuda.synchronize()
                dist.barrier(group=self.group)      # nobody still maps what is about to be freed
                release()
                return
            self._ext = _device_view(ext_ptr.value, n_floats, torch.float32, self.device).view(ext_shape)
        err = torch.zeros(1, dtype=torch.int32).pin_memory()     # written by the kernels on a wait timeout
        sync_ptrs = (C.c_void_p * 32)(*[peer_sync.get(r) for r in range(min(self.world, 32))])
        self._peer = dict(ext_ptr=ext_ptr.value, sync_ptr=sync_ptr.value, sync=sync, peer_ext=peer_ext, peer_sync=peer_sync,
                          sync_ptrs=sync_ptrs, err=err, epoch=0, counts_epoch=0)
        # best effort for a SlabGrid that is dropped without close(): free this rank's exported allocations (the
        # collective unmapping of close() is not possible from a finalizer)
        import weakref
        self._finalizer = weakref.finalize(self, _free_exported, self._peer)

    def _peer_check(self) -> None:
        if self._peer is not None and int(self._peer["err"][0]) != 0:
            raise RuntimeError("isoext_b200.dist: timed out waiting for a peer rank (NVLink peer transport)")

    def _guard_overwrite(self) -> None:
        """Before the owned planes are overwritten the neighbours must have pulled their halos of the last exchange."""
        pr = self._peer
        if pr is None or pr["epoch"] == 0:
            return
        from . import _lib
        from .grid import _stream_ptr
        done = [pr["peer_sync"][r] + 8 if r in pr["peer_ext"] else None for r in (self.rank - 1, self.rank + 1)]
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().isoext_peer_wait(done[0], done[1], pr["epoch"], pr["err"].data_ptr(), _stream_ptr()))

    def _peer_exchange(self, overlap: bool = False):
        from . import _lib
        from .grid import _stream_ptr
        lib, pr, p = _lib.lib(), self._peer, self.plan
        X, Y, Z = self.shape
        plane = Y * Z
        self._peer_check()
        pr["epoch"] += 1
        e = pr["epoch"]
        base = self._ext.data_ptr()
        seg = [(None, None, 0, None), (None, None, 0, None)]
        if self.rank > 0 and p["halo_below"]:
            q = slab_plan(X, self.rank - 1, self.world, self.cuts, self.halo_below)
            g0 = p["halo_below"][0]                                     # contiguous planes [g0, c_lo) of the lower neighbour
            src = pr["peer_ext"][self.rank - 1] + (g0 - q["ext_lo"]) * plane * 4
            seg[0] = (base + (g0 - p["ext_lo"]) * plane * 4, src, plane * len(p["halo_below"]), pr["peer_sync"][self.rank - 1])
        if self.rank < self.world - 1 and p["halo_above"]:
            q = slab_plan(X, self.rank + 1, self.world, self.cuts, self.halo_below)
            g0 = p["halo_above"][0]
            src = pr["peer_ext"][self.rank + 1] + (g0 - q["ext_lo"]) * plane * 4
            seg[1] = (base + (g0 - p["ext_lo"]) * plane * 4, src, plane * len(p["halo_above"]), pr["peer_sync"][self.rank + 1])
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            side = main
            if overlap:
                if pr.get("side") is None:
                    pr["side"] = torch.cuda.Stream(device=self.device)
                side = pr["side"]
                side.wait_stream(main)      # this epoch's values are in place, the previous extraction has read its halo
            with torch.cuda.stream(side):
                st = _stream_ptr()
                _lib.check(lib.isoext_peer_publish(pr["sync_ptr"], e, st))                    # my values of epoch e are in place
                _lib.check(lib.isoext_peer_halo_pull(seg[0][0], seg[0][1], seg[0][2], seg[0][3], seg[1][0], seg[1][1], seg[1][2],
                                                     seg[1][3], e, pr["err"].data_ptr(), st))
                _lib.check(lib.isoext_peer_publish(pr["sync_ptr"] + 8, e, st))                # I have pulled: neighbours may overwrite
                if overlap:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    return (len(p["halo_below"]), len(p["halo_above"]), ev)
        return None

    def close(self) -> None:
        """Collective: unmap the peers' memory and free the exported allocations."""
        pr, self._peer = self._peer, None
        if pr is None:
            return
        if getattr(self, "_finalizer", None) is not None:
            self._finalizer.detach()
        from . import _lib
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for ptr in list(pr["peer_ext"].values()) + list(pr["peer_sync"].values()):
                lib.isoext_peer_close(ptr)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            self._ext = torch.empty(0, dtype=torch.float32, device=self.device)
            pr["sync"] = None
            lib.isoext_peer_free(pr["ext_ptr"])
            lib.isoext_peer_free(pr["sync_ptr"])


def _free_exported(pr: dict) -> None:
    """Finalizer of a SlabGrid that was never close()d: frees the rank's own IPC-exported slab and SYNC block."""
    try:
        from . import _lib
        lib = _lib.lib()
        pr["sync"] = None
        torch.cuda.synchronize()
        lib.isoext_peer_free(pr["ext_ptr"])
        lib.isoext_peer_free(pr["sync_ptr"])
    except Exception:      # interpreter shutdown: the driver reclaims the memory
        pass


class _RawDevice:
    """Zero-copy torch view of a raw device allocation (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _device_view(ptr: int, n: int, dtype, device) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.int64: "<i8"}[dtype]
    return torch.as_tensor(_RawDevice(ptr, n, typestr), device=device)


def marching_cubes_local(sg: SlabGrid, level: float = 0.0, method: str = "nagae", halo=None):
    """The rank-local part of the distributed extraction (no communication): returns
    ``(v_own, f_local, n_lo, n_hi)`` where ``f_local`` still holds extended-slab vertex ids.
    ``halo``: what ``sg.exchange_halos(overlap=True)`` returned (halo planes still in flight on a side stream)."""
    from .mc import _method_id, mc_dense_raw
    mid = _method_id(method)
    p = sg.plan
    X, Y, Z = sg.shape
    with torch.cuda.device(sg.device):
        v_own, f, n_lo, n_hi, cap = mc_dense_raw(sg._ext, (p["n_ext"], Y, Z), sg.aabb_min, sg.aabb_max, level, mid, sg._ws,
                                                 cap_hint=sg._cap_hint, x_offset=p["ext_lo"], x_global=X,
                                                 emit_range=(p["emit_lo"], p["emit_hi"]), x_thresholds=sg.thresholds,
                                                 hints=sg._hints, halo=halo)
    sg._cap_hint = cap
    if v_own is None:
        return (torch.empty((0, 3), dtype=torch.float32, device=sg.device),
                torch.empty((0, 3), dtype=torch.int32, device=sg.device), 0, 0)
    return v_own, f, n_lo, n_hi


def relabel_faces_(f: torch.Tensor, n_lo: int, n_hi: int, base_mine: int, base_next: int) -> None:
    """In-place local -> global vertex ids (CUDA kernel isoext_relabel_faces)."""
    from . import _lib
    from .grid import _stream_ptr
    if f.numel():
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().isoext_relabel_faces(f.data_ptr(), f.numel(), n_lo, n_hi, base_mine, base_next, _stream_ptr()))


def marching_cubes(sg: SlabGrid, level: float = 0.0, method: str = "nagae", exchange: bool = True):
    """Distributed ``marching_cubes``: returns this rank's part ``(v_own, f_own)`` of the global mesh.

    Concatenating the parts of ranks 0..R-1 gives exactly the single-GPU ``(v, f)`` (global vertex ids).
    Either element may be an empty tensor on a rank whose slab misses the surface."""
    halo = sg.exchange_halos(overlap=True) if exchange else None
    v_own, f, n_lo, n_hi = marching_cubes_local(sg, level, method, halo=halo)
    globalize_faces_(sg, f, n_lo, n_hi, peer=exchange)
    return v_own, f


def globalize_faces_(sg: SlabGrid, f: torch.Tensor, n_lo: int, n_hi: int, peer: bool = True) -> None:
    """The cross-rank step after the local extraction (collective): local vertex ids in ``f`` -> global ids.

    Peer transport (only right after a peer halo exchange of the same call, whose epoch tags the counts): publish
    this rank's counts in its SYNC block and let the relabel kernel sum the lower ranks' counts over NVLink.
    Otherwise: ``all_gather`` of the counts + the plain relabel kernel."""
    pr = sg._peer
    if pr is not None and peer:
        from . import _lib
        from .grid import _stream_ptr
        lib = _lib.lib()
        sg._peer_check()          # the local extraction synchronised the stream: a pull timeout is visible now
        with torch.cuda.device(sg.device):
            st = _stream_ptr()
            _lib.check(lib.isoext_peer_publish_counts(pr["sync_ptr"], pr["epoch"], n_hi - n_lo, int(f.shape[0]), st))
            _lib.check(lib.isoext_relabel_faces_peer(f.data_ptr(), f.numel(), n_lo, n_hi, n_hi - n_lo, pr["sync_ptrs"], sg.rank,
                                                     pr["epoch"], None, pr["err"].data_ptr(), st))
        return
    vb, fb, totals, _ = global_bases(n_hi - n_lo, int(f.shape[0]), sg.device, sg.group)
    relabel_faces_(f, n_lo, n_hi, vb, vb + (n_hi - n_lo))


def dual_contouring_local(sg: SlabGrid, level: float = 0.0, reg: float = 1e-2, svd_tol: float = 1e-6):
    """Rank-local part of the distributed dual contouring (no communication): ``(v_own, f_local, n_lo, n_hi)``.

    The slab (SlabGrid(dc=True): point planes [c_r - 2, c_{r+1} + 1]) computes the intersections, normals and dual
    vertices of ALL its cell layers, welds them, emits the quads of the sign-change edges whose origin lies in its own
    planes [c_r, c_{r+1}) and owns the welded vertices with px[c_r] <= x < px[c_{r+1}] (dc_slab_plan)."""
    from .dc import dc_dense_raw, its_dense_raw
    if sg.halo_below < 2 and sg.world > 1:
        raise RuntimeError("dual contouring on slabs needs SlabGrid(..., dc=True): two halo planes below")
    p = sg.plan
    X, Y, Z = sg.shape
    with torch.cuda.device(sg.device):
        its, cap = its_dense_raw(sg._ext, (p["n_ext"], Y, Z), sg.aabb_min, sg.aabb_max, level, True, sg._ws, cap_hint=sg._cap_hint,
                                 x_offset=p["ext_lo"], x_global=X)
        sg._cap_hint = cap
        emit_hi = (p["c_hi"] if sg.rank < sg.world - 1 else X) - p["ext_lo"]
        v, f, _, _, n_lo, n_hi = dc_dense_raw(sg, its, reg, svd_tol, emit_range=(p["emit_lo"], emit_hi),
                                              x_thresholds=sg.thresholds, with_counts=True)
    if v is None:
        return (torch.empty((0, 3), dtype=torch.float32, device=sg.device),
                torch.empty((0, 3), dtype=torch.int32, device=sg.device), 0, 0)
    return v[n_lo:n_hi], f, n_lo, n_hi


def dual_contouring(sg: SlabGrid, level: float = 0.0, reg: float = 1e-2, svd_tol: float = 1e-6, exchange: bool = True):
    """Distributed ``dual_contouring`` (normals from the field, like the default of the single-device call): returns
    this rank's part ``(v_own, f_own)`` of the global mesh; the parts of ranks 0..R-1 concatenate to exactly the
    single-GPU ``(v, f)``."""
    if exchange:
        sg.exchange_halos()
    v_own, f, n_lo, n_hi = dual_contouring_local(sg, level, reg, svd_tol)
    globalize_faces_(sg, f, n_lo, n_hi, peer=exchange)
    return v_own, f


class SparseSlab:
    """This rank's part of a SparseGrid cut along x: the owned cells (x layers [c_r, c_{r+1})) plus one
    ghost layer of cells on either side, with their corner values.  Built from any SparseGrid that holds at least
    those cells (e.g. the whole band); afterwards the rank needs nothing from its neighbours but two counts, because
    sparse cells carry their own corner values -- there is no halo to exchange."""

    def __init__(self, grid, group=None, rank=None, world=None, cuts=None, dc: bool = False):
        from .sparse import SparseGrid
        self.group = group
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        shape = grid.shape
        self.dc = bool(dc)     # dual contouring: 3 ghost layers below, 2 above (sparse_slab_select)
        ext, owned = sparse_slab_select(grid._cells, shape, self.rank, self.world, cuts, below=3 if dc else 1, above=2 if dc else 1)
        c = partition_cells(shape[0], self.world) if cuts is None else check_cuts(shape[0], self.world, cuts)
        local = SparseGrid(list(shape), grid.aabb_min, grid.aabb_max, grid.default_value, device=grid.device)
        local._cells = grid._cells[ext].contiguous()
        local._values = grid._values[ext].contiguous()
        local._int32_api = grid._int32_api
        self.local, self.device, self.shape = local, grid.device, shape
        pos = torch.nonzero(owned[ext]).flatten()        # the owned cells are a contiguous run of the sorted list
        self.emit = (int(pos[0]), int(pos[-1]) + 1) if pos.numel() else (0, 0)
        from . import _lib
        px = lambda i: float(_lib.lib().isoext_axis_position(i, shape[0], grid.aabb_min[0], grid.aabb_max[0]))
        self.thresholds = (px(c[self.rank]) if self.rank > 0 else -math.inf,
                           px(c[self.rank + 1]) if self.rank < self.world - 1 else math.inf)
        self._peer = None

    def set_values(self, values8: torch.Tensor) -> None:
        self.local.set_values(values8)


def marching_cubes_sparse_local(ss: SparseSlab, level: float = 0.0, method: str = "nagae"):
    """Rank-local part (no communication): ``(v_own, f_local, n_lo, n_hi)``."""
    from .mc import _method_id
    from .sparse import mc_sparse
    v, f, n_lo, n_hi = mc_sparse(ss.local, level, _method_id(method), emit_range=ss.emit, x_thresholds=ss.thresholds, with_counts=True)
    if v is None:
        return (torch.empty((0, 3), dtype=torch.float32, device=ss.device),
                torch.empty((0, 3), dtype=torch.int32, device=ss.device), 0, 0)
    return v[n_lo:n_hi], f, n_lo, n_hi


def marching_cubes_sparse(ss: SparseSlab, level: float = 0.0, method: str = "nagae"):
    """Distributed ``marching_cubes`` over a slab-partitioned SparseGrid: this rank's ``(v_own, f_own)``; the parts
    of ranks 0..R-1 concatenate to exactly the single-device sparse mesh (global vertex ids)."""
    v_own, f, n_lo, n_hi = marching_cubes_sparse_local(ss, level, method)
    vb, fb, totals, _ = global_bases(n_hi - n_lo, int(f.shape[0]), ss.device, ss.group)
    relabel_faces_(f, n_lo, n_hi, vb, vb + (n_hi - n_lo))
    return v_own, f


def dual_contouring_sparse_local(ss: SparseSlab, level: float = 0.0, reg: float = 1e-2, svd_tol: float = 1e-6):
    """Rank-local part of dual contouring over a slab-partitioned SparseGrid (no communication):
    ``(v_own, f_local, n_lo, n_hi)``.  Needs ``SparseSlab(..., dc=True)``."""
    from .sparse import dc_sparse_raw, its_sparse
    if not ss.dc and ss.world > 1:
        raise RuntimeError("dual contouring on sparse slabs needs SparseSlab(..., dc=True): 3 ghost layers below, 2 above")
    its = its_sparse(ss.local, level, True)
    v, f, _, _, n_lo, n_hi = dc_sparse_raw(ss.local, its, reg, svd_tol, emit_range=ss.emit, x_thresholds=ss.thresholds, with_counts=True)
    if v is None:
        return (torch.empty((0, 3), dtype=torch.float32, device=ss.device),
                torch.empty((0, 3), dtype=torch.int32, device=ss.device), 0, 0)
    return v[n_lo:n_hi], f, n_lo, n_hi


def dual_contouring_sparse(ss: SparseSlab, level: float = 0.0, reg: float = 1e-2, svd_tol: float = 1e-6):
    """Distributed ``dual_contouring`` over a slab-partitioned SparseGrid: this rank's ``(v_own, f_own)``; the parts of
    ranks 0..R-1 concatenate to exactly the single-device sparse DC mesh (global vertex ids)."""
    v_own, f, n_lo, n_hi = dual_contouring_sparse_local(ss, level, reg, svd_tol)
    vb, fb, totals, _ = global_bases(n_hi - n_lo, int(f.shape[0]), ss.device, ss.group)
    relabel_faces_(f, n_lo, n_hi, vb, vb + (n_hi - n_lo))
    return v_own, f


def gather_mesh(v_own: torch.Tensor, f_own: torch.Tensor, group=None):
    """Concatenate the per-rank parts on every rank (surface-sized all_gather)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return v_own, f_own
    sizes = torch.tensor([v_own.shape[0], f_own.shape[0]], dtype=torch.int64, device=v_own.device)
    flat = torch.empty(world * 2, dtype=torch.int64, device=v_own.device)
    dist.all_gather_into_tensor(flat, sizes, group=group)
    alls = flat.view(world, 2).cpu()
    vs = [torch.empty((int(alls[r, 0]), 3), dtype=v_own.dtype, device=v_own.device) for r in range(world)]
    fs = [torch.empty((int(alls[r, 1]), 3), dtype=f_own.dtype, device=f_own.device) for r in range(world)]
    dist.all_gather(vs, v_own.contiguous(), group=group)
    dist.all_gather(fs, f_own.contiguous(), group=group)
    return torch.cat(vs), torch.cat(fs)
