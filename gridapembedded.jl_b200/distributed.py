"""Multi-GPU cut: z-slab partition of the Cartesian background model, one process per GPU.

Reference pattern: `src/Distributed/DistributedDiscretizations.jl:29-40` -- every rank cuts its local part of the model
(`map(local_views(bgmodel)) do m; cut(cutter,m,geo) end`), ghost consistency is checked with a `consistent!` exchange
(`:136-174`) and integrals are summed over the ranks by GridapDistributed.

Here the partition is along the LAST direction (slowest index), so each rank owns a contiguous range of global cell
ids and the global layout is the rank-major concatenation of the local ones.  Every background cell depends only on
its own nodes (`CutTriangulations.jl:589-614` isolates cells), therefore the cut itself needs NO data-path collective:
  * analytical leaves are re-evaluated on the shared node plane (bit-identical by construction, exactly like the
    reference re-evaluates on ghost nodes);
  * NODAL leaves owned per rank need the upper boundary node plane of the slab from the next rank:
    `exchange_ghost_plane` (point-to-point send/recv = the reference's `consistent!`);
  * global offsets of the concatenated layout = exclusive scan over the ranks of the local counts (`global_offsets`,
    one small all_gather); global integrals = all_reduce(sum) (`global_sum`).
`torch.distributed` (NCCL over NVLink on the GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple


def slab_range(nlast: int, rank: int, world: int) -> Tuple[int, int]:
    """Layers [k0,k1) of the last direction owned by `rank` (balanced contiguous split; the first `nlast % world`
    ranks get one more layer)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if world > nlast:
        raise ValueError("more ranks than layers")
    base, rem = divmod(nlast, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def estimate_cut_cells_per_layer(model, geo, coarse: int = 128):
    """Host estimate of the number of cut n-cubes in every layer of the last direction, from the signs of the level sets
    on a coarse node grid: blocks of f x f (x f) fine cells with the SAME f in every direction (f = ceil(max partition /
    coarse); anisotropic blocks would weight the surface orientations differently from the fine cells).  A block with
    mixed signs for some level set stands for about f^(D-1) fine cut cells (cut cells scale like area / h^2) spread over
    its f fine layers.  Needs closed-form or host-callable leaves; returns None for value-vector (DiscreteGeometry) leaves."""
    import numpy as np
    from .csg import find_unique_leaves
    from .geometries import LeafFunction
    payloads, _ = find_unique_leaves(geo.get_tree())
    if not all(isinstance(p, LeafFunction) for p in payloads):
        return None
    D = model.D
    part = list(model.partition)
    f = max(1, -(-max(part) // max(1, coarse)))
    cpart = [-(-n // f) for n in part]
    h = [(model.pmax[d] - model.pmin[d]) / part[d] for d in range(D)]
    axes = [model.pmin[d] + f * h[d] * np.arange(cpart[d] + 1) for d in range(D)]
    grids = np.meshgrid(*axes, indexing="ij")
    x = np.stack([g.ravel() for g in grids], axis=1)
    shape = tuple(c + 1 for c in cpart)
    cut = np.zeros(tuple(cpart), dtype=bool)
    for p in payloads:
        out = (p(x) > 0.0).reshape(shape)
        any_ = np.zeros(tuple(cpart), dtype=bool)
        all_ = np.ones(tuple(cpart), dtype=bool)
        for corner in range(1 << D):
            sl = tuple(slice((corner >> d) & 1, ((corner >> d) & 1) + cpart[d]) for d in range(D))
            any_ |= out[sl]
            all_ &= out[sl]
        cut |= any_ & ~all_
    per_block_layer = cut.reshape(-1, cpart[-1]).sum(axis=0).astype(np.float64)
    per_fine_layer = per_block_layer * float(f) ** (D - 1) / f  # f^(D-1) fine cut cells per block over f layers
    return np.repeat(per_fine_layer, f)[:part[-1]]


def balanced_slab_ranges(model, geo, world: int, cut_weight: float = 600.0, coarse: int = 128) -> List[Tuple[int, int]]:
    """Work-balanced z-slabs: contiguous layer ranges [k0,k1) per rank such that the estimated work
    `layers * cubes_per_layer + cut_weight * cut cubes` is as even as possible (the classification sweep costs the same
    for every cube, the subtriangulation and the integrals are proportional to the cut cells -- SURVEY.md section 8e:
    "load is unbalanced in the cut work").  Deterministic and communication-free: every rank computes the same
    partition from the geometry.  Falls back to the even split when the level sets cannot be evaluated on the host."""
    import numpy as np
    nz = model.partition[-1]
    if world > nz:
        raise ValueError("more ranks than layers")
    est = estimate_cut_cells_per_layer(model, geo, coarse) if world > 1 else None
    if est is None:
        return [slab_range(nz, r, world) for r in range(world)]
    cubes_per_layer = model.num_cubes // nz
    w = cubes_per_layer + float(cut_weight) * est
    pre = np.concatenate([[0.0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        target = pre[-1] * r / world
        k = int(np.searchsorted(pre, target, side="left"))
        if k > 0 and abs(pre[k - 1] - target) < abs(pre[min(k, nz)] - target):
            k -= 1
        k = max(bounds[-1] + 1, min(k, nz - (world - r)))  # every rank keeps at least one layer
        bounds.append(k)
    bounds.append(nz)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def slab_cell_offset(model, k0: int) -> int:
    """Global id offset of the first cell of the slab starting at layer k0."""
    layer_cubes = model.num_cubes // model.partition[-1]
    return k0 * layer_cubes * (model.simplices_per_cube if model.simplexified else 1)


def slab_node_offset(model, k0: int) -> int:
    layer_nodes = model.num_nodes // (model.partition[-1] + 1)
    return k0 * layer_nodes


def _dist():
    import torch.distributed as dist
    return dist


def _active(group=None) -> bool:
    dist = _dist()
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def global_offsets(local_counts: Dict[str, int], group=None, device=None) -> Tuple[Dict[str, int], Dict[str, int]]:
    """Exclusive scan over the ranks of the local counts: (offsets of this rank, global totals).  One all_gather of
    len(local_counts) int64 per rank."""
    keys = sorted(local_counts)
    if not _active(group):
        return {k: 0 for k in keys}, {k: int(local_counts[k]) for k in keys}
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = torch.tensor([int(local_counts[k]) for k in keys], dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    allc = torch.stack(out).cpu()
    offs = allc[:rank].sum(0) if rank > 0 else torch.zeros_like(allc[0])
    tot = allc.sum(0)
    return ({k: int(offs[i]) for i, k in enumerate(keys)}, {k: int(tot[i]) for i, k in enumerate(keys)})


def global_sum(values: Sequence[float], group=None, device=None) -> List[float]:
    """all_reduce(sum) of a few float64 numbers (volumes, surface measures, ...)."""
    if not _active(group):
        return [float(v) for v in values]
    import torch
    dist = _dist()
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.cpu()]


def exchange_ghost_plane(owned_values, layer_nodes: int, group=None):
    """NODAL level sets distributed by z-slabs: `owned_values` holds the node planes this rank owns (torch tensor,
    a multiple of layer_nodes values; rank r owns planes k0..k1-1, the last rank also the top plane).  Returns the
    (k1-k0+1)-plane vector the slab cut needs, i.e. with the first plane of the next rank appended.  One send + one
    recv of layer_nodes doubles per rank boundary (the reference's `consistent!`,
    DistributedDiscretizations.jl:164-170)."""
    if not _active(group):
        return owned_values
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def grank(r):
        return dist.get_global_rank(group, r) if group is not None else r

    reqs = []
    ghost = None
    if rank > 0:  # my first plane is the ghost plane of the previous rank
        reqs.append(dist.isend(owned_values[:layer_nodes].contiguous(), dst=grank(rank - 1), group=group))
    if rank < world - 1:
        ghost = torch.empty(layer_nodes, dtype=owned_values.dtype, device=owned_values.device)
        reqs.append(dist.irecv(ghost, src=grank(rank + 1), group=group))
    for r in reqs:
        r.wait()
    return owned_values if ghost is None else torch.cat([owned_values, ghost])


def exchange_ghost_plane_(slab_values, layer_nodes: int, group=None):
    """In-place form of `exchange_ghost_plane` for a re-cut loop: `slab_values` is the persistent (k1-k0+1)-plane buffer of
    this rank (torch tensor); its first plane is sent to the previous rank and its last plane is overwritten with the
    first plane of the next rank (the last rank owns its top plane).  No allocation, one send + one recv of layer_nodes
    doubles per rank boundary."""
    if not _active(group):
        return slab_values
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def grank(r):
        return dist.get_global_rank(group, r) if group is not None else r

    reqs = []
    if rank > 0:
        reqs.append(dist.isend(slab_values[:layer_nodes], dst=grank(rank - 1), group=group))
    if rank < world - 1:
        reqs.append(dist.irecv(slab_values[slab_values.numel() - layer_nodes:], src=grank(rank + 1), group=group))
    for r in reqs:
        r.wait()
    return slab_values


def init_library_communicator(ctx, group=None):
    """Give `ctx` its own NCCL communicator (the library's, `gecut_comm_init`) over the ranks of a torch.distributed group:
    rank 0 creates the unique id, the 128 bytes travel by one `broadcast_object_list` (gloo or nccl -- the host program's
    existing channel, MPI.Bcast in the Julia host), every rank joins.  After this the data path of a multi-GPU step
    (ghost planes, global sums, consistency flag, sizes) runs inside libgecut.so on the context's stream."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("torch.distributed is not initialised")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(box, src=src, group=group)
    ctx.comm_init(box[0], rank, world)
    return rank, world


class DistributedLevelSetCutter:
    """`cut(::DistributedDiscreteModel, geo)` analogue: each rank cuts its z-slab on its own GPU."""

    def __init__(self, ctx=None, group=None, nodal_slab_local: bool = False):
        dist = _dist()
        self.group = group
        on = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        self.ctx = ctx
        self.nodal_slab_local = nodal_slab_local

    def slab(self, model) -> Tuple[int, int]:
        return slab_range(model.partition[-1], self.rank, self.world)

    def cut(self, model, geo):
        from .cutters import LevelSetCutter
        slab = self.slab(model) if self.world > 1 else None
        return LevelSetCutter(self.ctx, slab=slab,
                              nodal_slab_local=self.nodal_slab_local and slab is not None).cut(model, geo)

    def global_layout(self, cutgeo, device=None):
        """(offsets of this rank, global totals) of the concatenated layout."""
        s = cutgeo.sizes
        local = {"cutcells": int(s.num_cutcells), "subcells": int(s.num_subcells),
                 "subcell_points": int(s.num_subcell_points), "subfacets": int(s.num_subfacets),
                 "subfacet_points": int(s.num_subfacet_points), "bgcells": int(s.num_bgcells)}
        return global_offsets(local, self.group, device)
