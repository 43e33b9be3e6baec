"""AgFEM cell aggregation (host mirror of `src/AgFEM/CellAggregation.jl:1-102`).

`aggregate(strategy, cut[, cut_facets][, geo or name][, in_or_out])` returns `cell_to_cellin` (Int32, 1-based root cell of
every active cell, 0 for the others), computed on the device by `gecut_aggregate`; the neighbour search uses the closed
form topology of the Cartesian model, the facet states come from `cut_facets` / `compute_bgfacet_to_inoutcut`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .csg import Geometry
from .cutters import LevelSetCutter
from .discretizations import IN, OUT, EmbeddedDiscretization
from .facets import EmbeddedFacetDiscretization, _call_cut_facets


class AggregateCutCellsByThreshold:
    """`AggregateCutCellsByThreshold(threshold)` (`CellAggregation.jl:64-68`)."""

    def __init__(self, threshold: float):
        if threshold < 0.0 or threshold > 1.0:
            raise ValueError("Agg. threshold must be in [0,1]")
        self.threshold = float(threshold)


def AggregateAllCutCells() -> AggregateCutCellsByThreshold:
    """`AggregateAllCutCells() = AggregateCutCellsByThreshold(1.0)` (`CellAggregation.jl:70`)."""
    return AggregateCutCellsByThreshold(1.0)


def aggregate(strategy: AggregateCutCellsByThreshold, cut: EmbeddedDiscretization, *args, device_output: bool = False):
    """`aggregate(strategy, cut)`, `(strategy, cut, geo)`, `(strategy, cut, name, in_or_out)`,
    `(strategy, cut, geo, in_or_out)`, `(strategy, cut, cut_facets[, geo or name[, in_or_out]])`
    (`CellAggregation.jl:1-102`)."""
    args = list(args)
    facets: Optional[EmbeddedFacetDiscretization] = None
    if args and isinstance(args[0], EmbeddedFacetDiscretization):
        facets = args.pop(0)
    geo = args.pop(0) if args else None
    in_or_out = args.pop(0) if args else IN
    if args:
        raise TypeError("too many arguments")
    if in_or_out not in (IN, OUT):
        raise ValueError("in_or_out must be IN or OUT")
    if not isinstance(strategy, AggregateCutCellsByThreshold):
        raise NotImplementedError("abstract method")  # CellAggregation.jl:14-16
    own = None
    if facets is None:
        # compute_bgfacet_to_inoutcut(cut.bgmodel, geo) (`CellAggregation.jl:79`): classification of the facet grid only
        # (of the WHOLE geometry, so that the level-set rows line up with cut.program(geo))
        own = facets = _call_cut_facets(LevelSetCutter(cut.ctx), cut.bgmodel, cut.geo, True)
    prog = cut.program(geo)
    # compute_bgfacet_to_inoutcut(cut_facets, geo) uses cut_facets' OWN oid_to_ls (`CellAggregation.jl:90`): the facet
    # discretization may come from another Geometry object or leaf order, so the geometry is compiled a second time over it
    fprog = facets.program(geo)
    ctx = cut.ctx
    nc = cut.num_bgcells
    try:
        if device_output:
            import torch
            out = torch.empty(nc, dtype=torch.int32, device=f"cuda:{ctx.device}")
            ctx.check(ctx._lib.gecut_aggregate(cut.handle, facets.handle, prog.ctypes.data, prog.size, fprog.ctypes.data,
                                               fprog.size, int(in_or_out), C.c_double(strategy.threshold),
                                               C.c_void_p(out.data_ptr()), 1),
                      "gecut_aggregate")
        else:
            out = np.empty(nc, np.int32)
            ctx.check(ctx._lib.gecut_aggregate(cut.handle, facets.handle, prog.ctypes.data, prog.size, fprog.ctypes.data,
                                               fprog.size, int(in_or_out), C.c_double(strategy.threshold),
                                               C.c_void_p(out.ctypes.data), 0),
                      "gecut_aggregate")
    finally:
        if own is not None:
            own.close()
    return out


def aggregate_slabs(strategy: AggregateCutCellsByThreshold, cuts, facets=None, geo=None, in_or_out=IN, comm: bool = False):
    """Aggregation over z-slabs (`src/Distributed/DistributedAgFEM.jl:52-144`): `cuts` = the slab discretizations of one
    model in ascending layer order (results of `LevelSetCutter(ctx, slab=(k0, k1)).cut`), `facets` = the matching slab facet
    discretizations (default: classified here).  Returns one `cell_to_cellin` per slab with GLOBAL 1-based root ids; the
    concatenation equals `aggregate` on the whole model.  `comm=True`: ONE slab per rank (`cuts` is this rank's cut), ghost
    layers and termination through the library's NCCL communicator (`gecut_comm_aggregate`, collective)."""
    if not isinstance(strategy, AggregateCutCellsByThreshold):
        raise NotImplementedError("abstract method")
    if in_or_out not in (IN, OUT):
        raise ValueError("in_or_out must be IN or OUT")
    single = isinstance(cuts, EmbeddedDiscretization)
    cuts = [cuts] if single else list(cuts)
    if comm and len(cuts) != 1:
        raise ValueError("comm=True takes the one slab this rank owns")
    ctx = cuts[0].ctx
    own = []
    if facets is None:
        facets = []
        for c in cuts:
            slab = c.slab if getattr(c, "slab", None) is not None else None
            f = _call_cut_facets(LevelSetCutter(ctx, slab=slab), c.bgmodel, c.geo, True)
            own.append(f)
            facets.append(f)
    else:
        facets = [facets] if isinstance(facets, EmbeddedFacetDiscretization) else list(facets)
    if len(facets) != len(cuts):
        raise ValueError("one facet discretization per slab")
    prog = cuts[0].program(geo)
    fprog = facets[0].program(geo)
    outs = [np.empty(c.num_bgcells, np.int32) for c in cuts]
    try:
        if comm:
            ctx.check(ctx._lib.gecut_comm_aggregate(cuts[0].handle, facets[0].handle, prog.ctypes.data, prog.size,
                                                    fprog.ctypes.data, fprog.size, int(in_or_out),
                                                    C.c_double(strategy.threshold), C.c_void_p(outs[0].ctypes.data), 0),
                      "gecut_comm_aggregate")
        else:
            n = len(cuts)
            hc = (C.c_void_p * n)(*[c.handle for c in cuts])
            hf = (C.c_void_p * n)(*[f.handle for f in facets])
            ho = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
            ctx.check(ctx._lib.gecut_aggregate_slabs(hc, hf, n, prog.ctypes.data, prog.size, fprog.ctypes.data, fprog.size,
                                                     int(in_or_out), C.c_double(strategy.threshold), ho, 0),
                      "gecut_aggregate_slabs")
    finally:
        for f in own:
            f.close()
    return outs[0] if single else outs


def GhostSkeleton(cut: EmbeddedDiscretization, in_or_out=None, geo=None, ids: bool = False):
    """`GhostSkeleton(cut[, in_or_out=ACTIVE_IN][, geo or name])` (`EmbeddedDiscretizations.jl:465-543`): the
    `facet_to_mask` (bool per bg facet, Gridap facet numbering) the reference hands to `SkeletonTriangulation(model, mask)`;
    with `ids=True` the ascending 1-based facet ids instead."""
    from .discretizations import ACTIVE_IN, ActiveInOrOut, CutInOrOut
    if isinstance(in_or_out, (str, Geometry)) and geo is None:
        in_or_out, geo = None, in_or_out
    if in_or_out is None:
        in_or_out = ACTIVE_IN
    if isinstance(in_or_out, tuple):  # PHYSICAL_IN / PHYSICAL_OUT (`:508`)
        raise NotImplementedError("Not implemented but not needed in practice. Ghost stabilization can be integrated in "
                                  "full facets.")
    side = in_or_out if isinstance(in_or_out, int) else in_or_out.in_or_out
    if side not in (IN, OUT, 0):
        raise ValueError("in_or_out must be ACTIVE_IN, ACTIVE_OUT or IN / OUT / CUT")
    prog = cut.program(geo)
    ctx = cut.ctx
    n = C.c_int64()
    if ids:
        ctx.check(ctx._lib.gecut_ghost_skeleton(cut.handle, prog.ctypes.data, prog.size, int(side), C.byref(n), None, None),
                  "gecut_ghost_skeleton")
        out = np.empty(int(n.value), np.int32)
        if out.size:
            ctx.check(ctx._lib.gecut_ghost_skeleton(cut.handle, prog.ctypes.data, prog.size, int(side), C.byref(n), None,
                                                    C.c_void_p(out.ctypes.data)), "gecut_ghost_skeleton")
        return out
    mask = np.empty(_local_num_facets(cut.bgmodel, None), np.int8)
    ctx.check(ctx._lib.gecut_ghost_skeleton(cut.handle, prog.ctypes.data, prog.size, int(side), C.byref(n),
                                            C.c_void_p(mask.ctypes.data), None), "gecut_ghost_skeleton")
    return mask.astype(bool)


def _local_num_facets(model, slab) -> int:
    part = list(model.partition)
    if slab is not None:
        part[-1] = int(slab[1]) - int(slab[0])
    D = model.D
    nf = sum(int(np.prod([part[e] + (1 if e == d else 0) for e in range(D)])) for d in range(D))
    if model.simplexified:
        # facets of the TRI / TET mesh: the n-cube faces split into 1 / 2 simplices + 1 / 6 facets inside every cube
        nf = (1 if D == 2 else 2) * nf + (1 if D == 2 else 6) * int(np.prod(part))
    return nf


def GhostSkeleton_slabs(cuts, in_or_out=None, geo=None, comm: bool = False):
    """`GhostSkeleton` over z-slabs (the reference's ranks build it on their local models, `src/Distributed`): one
    `facet_to_mask` per slab in the facet numbering of the slab's local model (`LevelSetCutter(ctx, slab).cut_facets`); the
    cells across a shared plane are read from the adjacent slab (`comm=True`: ONE slab per rank, boundary cell layers over the
    library's NCCL communicator)."""
    from .discretizations import ACTIVE_IN
    if isinstance(in_or_out, (str, Geometry)) and geo is None:
        in_or_out, geo = None, in_or_out
    if in_or_out is None:
        in_or_out = ACTIVE_IN
    if isinstance(in_or_out, tuple):
        raise NotImplementedError("Not implemented but not needed in practice. Ghost stabilization can be integrated in "
                                  "full facets.")
    side = in_or_out if isinstance(in_or_out, int) else in_or_out.in_or_out
    if side not in (IN, OUT, 0):
        raise ValueError("in_or_out must be ACTIVE_IN, ACTIVE_OUT or IN / OUT / CUT")
    single = isinstance(cuts, EmbeddedDiscretization)
    cuts = [cuts] if single else list(cuts)
    if comm and len(cuts) != 1:
        raise ValueError("comm=True takes the one slab this rank owns")
    ctx = cuts[0].ctx
    prog = cuts[0].program(geo)
    masks = [np.empty(_local_num_facets(c.bgmodel, getattr(c, "slab", None)), np.int8) for c in cuts]
    n = len(cuts)
    counts = (C.c_int64 * n)()
    if comm:
        ctx.check(ctx._lib.gecut_comm_ghost_skeleton(cuts[0].handle, prog.ctypes.data, prog.size, int(side), counts,
                                                     C.c_void_p(masks[0].ctypes.data)), "gecut_comm_ghost_skeleton")
    else:
        hc = (C.c_void_p * n)(*[c.handle for c in cuts])
        hm = (C.c_void_p * n)(*[m.ctypes.data for m in masks])
        ctx.check(ctx._lib.gecut_ghost_skeleton_slabs(hc, n, prog.ctypes.data, prog.size, int(side), counts, hm),
                  "gecut_ghost_skeleton_slabs")
    for m, k in zip(masks, counts):
        assert int(m.sum()) == int(k)
    out = [m.astype(bool) for m in masks]
    return out[0] if single else out
