"""`cut_facets`, `EmbeddedFacetDiscretization` and the triangulations built on it (host mirror).

Reference: `src/LevelSetCutters/LevelSetCutters.jl:107-142,159-173` (`cut_facets`, `compute_bgfacet_to_inoutcut`),
`src/Interfaces/EmbeddedFacetDiscretizations.jl:28-35` (the struct), `:54-103` (`SkeletonTriangulation`),
`:112-201` (`BoundaryTriangulation`), `:216-248` (`compute_bgfacet_to_inoutcut(cut,geo)`,
`compute_subfacet_to_inout(cut,geo)`).

The facet grid of the Cartesian model is never materialised on the host: facet ids follow Gridap's first-encounter
numbering (assumption A8) and the device decodes them in closed form.  n-cube models only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np

from . import _lib
from .csg import Geometry, compile_postfix, get_geometry
from .cutters import LevelSetCutter, flatten_geometry, make_grid, make_leaf_records, _nodal_pointers
from .discretizations import (IN, OUT, CUT, PHYSICAL_IN, ActiveInOrOut, CutInOrOut, EmbeddedDiscretization,
                              SubCellData)


class EmbeddedFacetDiscretization:
    """`EmbeddedFacetDiscretization{Dc,Dp,T}` (`EmbeddedFacetDiscretizations.jl:28-35`): `bgmodel`,
    `ls_to_facet_to_inoutcut`, `subfacets::SubCellData{D-1,D}`, `ls_to_subfacet_to_inout`, `oid_to_ls`, `geo`."""

    def __init__(self, ctx: _lib.Context, handle, bgmodel, geo: Geometry, oid_to_ls: dict):
        self.ctx = ctx
        self._h = handle
        self.bgmodel = bgmodel
        self.geo = geo
        self.oid_to_ls = oid_to_ls
        s = _lib.FacetSizes()
        ctx.check(ctx._lib.gecut_facet_result_sizes(self._h, C.byref(s)), "gecut_facet_result_sizes")
        self.sizes = s
        self._host = None

    D = property(lambda self: int(self.sizes.D))
    L = property(lambda self: int(self.sizes.num_levelsets))
    num_bgfacets = property(lambda self: int(self.sizes.num_bgfacets))
    num_cutfacets = property(lambda self: int(self.sizes.num_cutfacets))
    handle = property(lambda self: self._h)

    def close(self):
        if self._h:
            self.ctx._lib.gecut_facet_result_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_host(self, topology: bool = False) -> dict:
        """Copy the layout to the host; `topology` adds the facet grid itself (`facet_to_nodes`, `facet_is_boundary`)."""
        if self._host is not None and (not topology or "facet_to_nodes" in self._host):
            return self._host
        s = self.sizes
        D, L, nvf = int(s.D), int(s.num_levelsets), int(s.num_vertices_per_bgfacet)
        shapes = {
            "ls_to_facet_to_inoutcut": ((L, s.num_bgfacets), np.int8),
            "cutfacet_to_facet": ((s.num_cutfacets,), np.int32),
            "sf_cell_to_points": ((s.num_subfacets, D), np.int32),
            "sf_cell_to_bgcell": ((s.num_subfacets,), np.int32),
            "sf_point_to_coords": ((s.num_subfacet_points, D), np.float64),
            "sf_point_to_rcoords": ((s.num_subfacet_points, D - 1), np.float64),
            "ls_to_subfacet_to_inout": ((L, s.num_subfacets), np.int8),
        }
        if topology:
            shapes["facet_to_nodes"] = ((s.num_bgfacets, nvf), np.int32)
            shapes["facet_is_boundary"] = ((s.num_bgfacets,), np.int8)
        host = {}
        b = _lib.FacetBuffers()
        for name, (shape, dt) in shapes.items():
            a = np.empty(tuple(int(x) for x in shape), dtype=dt)
            host[name] = a
            setattr(b, name, a.ctypes.data if a.size else None)
        self.ctx.check(self.ctx._lib.gecut_facet_result_copy(self._h, C.byref(b)), "gecut_facet_result_copy")
        self._host = host
        return host

    @property
    def ls_to_facet_to_inoutcut(self) -> np.ndarray:
        return self.to_host()["ls_to_facet_to_inoutcut"]

    @property
    def subfacets(self) -> SubCellData:
        h = self.to_host()
        return SubCellData(h["sf_cell_to_points"], h["sf_cell_to_bgcell"], h["sf_point_to_coords"],
                           h["sf_point_to_rcoords"])

    @property
    def ls_to_subfacet_to_inout(self) -> np.ndarray:
        return self.to_host()["ls_to_subfacet_to_inout"]

    def program(self, geo_or_name: Union[None, str, Geometry]) -> np.ndarray:
        if geo_or_name is None:
            geo = self.geo
        elif isinstance(geo_or_name, str):
            geo = get_geometry(self.geo, geo_or_name)
        else:
            geo = geo_or_name
        prog = compile_postfix(geo.get_tree(), self.oid_to_ls)
        if len(prog) > _lib.MAX_PROGRAM:
            raise NotImplementedError("geometry tree too large for the selector kernels")
        return np.asarray(prog, dtype=np.int32)


def _call_cut_facets(cutter: LevelSetCutter, background, geom, classify_only: bool) -> EmbeddedFacetDiscretization:
    """With a z-slab cutter (`LevelSetCutter(ctx, slab=(k0, k1))`) the result is the facet discretization of the rank's LOCAL
    model -- layers [k0, k1) as a Cartesian model of its own, local facet numbering, global node coordinates -- the way every
    rank of the reference's distributed mode cuts its own local model (`src/Distributed/DistributedDiscretizations.jl:42-53`)."""
    ctx = cutter.ctx
    num_nodal, planes = None, None
    if cutter.nodal_slab_local and cutter.slab is not None:
        layer_nodes = background.num_nodes // (background.partition[-1] + 1)
        num_nodal = layer_nodes * (cutter.slab[1] - cutter.slab[0] + 1)
        planes = (cutter.slab[0], cutter.slab[1])
    leaves, nodal, oid_to_ls = flatten_geometry(background, geom, num_nodal=num_nodal, nodal_planes=planes)
    grid = make_grid(background, cutter.slab)
    grid.nodal_slab_local = 1 if num_nodal is not None else 0
    recs = make_leaf_records(leaves)
    ptrs, on_device, keep = _nodal_pointers(ctx, background, nodal)
    out = C.c_void_p()
    rc = ctx._lib.gecut_cut_facets(ctx.handle, C.byref(grid), recs, len(leaves), ptrs, 1 if on_device else 0,
                                   1 if classify_only else 0, C.byref(out))
    ctx.check(rc, "gecut_cut_facets")
    del keep
    local = background
    if cutter.slab is not None:  # the local model of the rank (facet ids, boundary flags and sizes refer to it)
        from .cartesian import CartesianDiscreteModel
        k0, k1 = int(cutter.slab[0]), int(cutter.slab[1])
        D = background.D
        dz = background.sizes[D - 1]
        pmin, pmax, part = list(background.pmin), list(background.pmax), list(background.partition)
        pmax[D - 1] = pmin[D - 1] + k1 * dz
        pmin[D - 1] = pmin[D - 1] + k0 * dz
        part[D - 1] = k1 - k0
        local = CartesianDiscreteModel(tuple(pmin), tuple(pmax), tuple(part), simplexified=background.simplexified)
    return EmbeddedFacetDiscretization(ctx, out, local, geom, oid_to_ls)


def _cutter_cut_facets(self, background, geom):
    """`cut_facets(cutter::LevelSetCutter, background, geom)` (`LevelSetCutters.jl:107-110`)."""
    return _call_cut_facets(self, background, geom, False)


def _cutter_compute_bgfacet_to_inoutcut(self, background, geom) -> np.ndarray:
    """`compute_bgfacet_to_inoutcut(::LevelSetCutter, model, geom)` (`LevelSetCutters.jl:169-173`)."""
    disc = _call_cut_facets(self, background, geom, True)
    return compute_bgfacet_to_inoutcut(disc, geom)


LevelSetCutter.cut_facets = _cutter_cut_facets
LevelSetCutter.compute_bgfacet_to_inoutcut = _cutter_compute_bgfacet_to_inoutcut


def cut_facets(*args):
    """`cut_facets(background, geom)`, `cut_facets(cutter, background, geom)`, `cut_facets(cut::EmbeddedDiscretization,
    geo)` or `cut_facets(cut::EmbeddedFacetDiscretization, ...)` (`LevelSetCutters.jl:107-128`, `Cutters.jl:49-51,66-68`)."""
    if isinstance(args[0], EmbeddedFacetDiscretization):
        return args[0]
    if isinstance(args[0], EmbeddedDiscretization):
        cut, geo = args[0], (args[1] if len(args) > 1 else args[0].geo)
        return LevelSetCutter(cut.ctx).cut_facets(cut.bgmodel, geo)
    if len(args) == 2:
        return LevelSetCutter().cut_facets(*args)
    cutter, background, geom = args
    return cutter.cut_facets(background, geom)


def compute_bgfacet_to_inoutcut(*args) -> np.ndarray:
    """`compute_bgfacet_to_inoutcut(cut_facets, geo)` (`EmbeddedFacetDiscretizations.jl:216-231`),
    `(model, geom)` or `(cutter, model, geom)` (`LevelSetCutters.jl:159-173`)."""
    if isinstance(args[0], EmbeddedFacetDiscretization):
        cut = args[0]
        prog = cut.program(args[1] if len(args) > 1 else None)
        out = np.empty(cut.num_bgfacets, dtype=np.int8)
        cut.ctx.check(cut.ctx._lib.gecut_bgfacet_to_inoutcut(cut.handle, prog.ctypes.data, prog.size, out.ctypes.data),
                      "gecut_bgfacet_to_inoutcut")
        return out
    if len(args) == 2:
        return LevelSetCutter().compute_bgfacet_to_inoutcut(*args)
    cutter, background, geom = args
    return cutter.compute_bgfacet_to_inoutcut(background, geom)


def compute_subfacet_to_inout(cut: EmbeddedFacetDiscretization, geo=None) -> np.ndarray:
    """`compute_subfacet_to_inout(cut, geo)` (`EmbeddedFacetDiscretizations.jl:233-248`)."""
    prog = cut.program(geo)
    out = np.empty(int(cut.sizes.num_subfacets), dtype=np.int8)
    cut.ctx.check(cut.ctx._lib.gecut_subfacet_to_inout(cut.handle, prog.ctypes.data, prog.size, out.ctypes.data),
                  "gecut_subfacet_to_inout")
    return out


class CutFacetTriangulation:
    """`BoundaryTriangulation(cut_facets, in_or_out, geo)` or one side of `SkeletonTriangulation(cut_facets, ...)`:
    the selected sub-facets (`SubFacetBoundaryTriangulation`, `EmbeddedFacetDiscretizations.jl:266-313`) and the
    selected whole facets, which the reference `lazy_append`s (`:147-156`)."""

    def __init__(self, cut: EmbeddedFacetDiscretization, handle, which: int):
        self.cut = cut
        self.ctx = cut.ctx
        self._h = handle
        self.which = which
        a, b = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx._lib.gecut_facet_selection_sizes(self._h, C.byref(a), C.byref(b)),
                       "gecut_facet_selection_sizes")
        self.num_sub, self.num_bg = int(a.value), int(b.value)
        self._ids = None

    handle = property(lambda self: self._h)

    def num_cells(self) -> int:
        return self.num_sub + self.num_bg

    def close(self):
        if self._h:
            self.ctx._lib.gecut_facet_selection_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _copy(self):
        if self._ids is None:
            sub = np.empty(self.num_sub, np.int32)
            bg = np.empty(self.num_bg, np.int32)
            self.ctx.check(self.ctx._lib.gecut_facet_selection_copy(
                self._h, sub.ctypes.data if sub.size else None, bg.ctypes.data if bg.size else None),
                "gecut_facet_selection_copy")
            self._ids = (sub, bg)
        return self._ids

    @property
    def sub_ids(self) -> np.ndarray:
        return self._copy()[0]

    @property
    def bg_ids(self) -> np.ndarray:
        return self._copy()[1]

    def measure(self, return_values: bool = False):
        """`sum(∫(1)dΓ)` with `Measure(trian, 2)`."""
        total = C.c_double()
        vals = np.empty(self.num_sub, np.float64) if return_values else None
        self.ctx.check(self.ctx._lib.gecut_facet_selection_measure(
            self._h, vals.ctypes.data if (vals is not None and vals.size) else None, C.byref(total)),
            "gecut_facet_selection_measure")
        return (float(total.value), vals) if return_values else float(total.value)


def _select(cut: EmbeddedFacetDiscretization, which: int, in_or_out, geo) -> CutFacetTriangulation:
    if isinstance(in_or_out, (str, Geometry)) and geo is None:
        in_or_out, geo = PHYSICAL_IN, in_or_out
    prog = cut.program(geo)
    if isinstance(in_or_out, tuple):
        a, b = in_or_out
        if not (isinstance(a, CutInOrOut) and isinstance(b, int) and a.in_or_out == b):
            raise NotImplementedError("only (CutInOrOut(s), s) tuples (PHYSICAL_IN / PHYSICAL_OUT) are supported")
        kind, side = _lib.SEL_PHYSICAL, b
    elif isinstance(in_or_out, CutInOrOut):
        kind, side = _lib.SEL_CUTONLY, in_or_out.in_or_out
    elif isinstance(in_or_out, ActiveInOrOut):
        kind, side = _lib.SEL_ACTIVE, in_or_out.in_or_out
    elif isinstance(in_or_out, int):
        kind, side = _lib.SEL_BGONLY, in_or_out
    else:
        raise TypeError(f"unsupported selector {in_or_out!r}")
    if side not in (IN, OUT):
        raise NotImplementedError("side must be IN or OUT")
    out = C.c_void_p()
    cut.ctx.check(cut.ctx._lib.gecut_select_facets(cut.handle, which, kind, prog.ctypes.data, prog.size, side,
                                                   C.byref(out)), "gecut_select_facets")
    return CutFacetTriangulation(cut, out, which)


def BoundaryTriangulation(cut: EmbeddedFacetDiscretization, in_or_out=PHYSICAL_IN, geo=None) -> CutFacetTriangulation:
    """`BoundaryTriangulation(cut::EmbeddedFacetDiscretization[, in_or_out][, geo or name])` (`:112-201`) on all
    boundary facets of the model (`tags=nothing`)."""
    return _select(cut, _lib.FACETS_BOUNDARY, in_or_out, geo)


def SkeletonTriangulation(cut: EmbeddedFacetDiscretization, in_or_out=PHYSICAL_IN, geo=None) -> CutFacetTriangulation:
    """`SkeletonTriangulation(cut::EmbeddedFacetDiscretization[, in_or_out][, geo or name])` (`:54-103`): the plus and
    the minus side select the same facets, so one selection stands for both."""
    return _select(cut, _lib.FACETS_SKELETON, in_or_out, geo)


def facet_topology(model, slab=None):
    """`(facet_to_nodes, facet_is_boundary)` of `Grid(ReferenceFE{D-1}, model)` in Gridap's numbering, from the closed
    forms of the library on the HOST (`gecut_facet_topology_host`: index arithmetic only, no device) -- what
    `get_face_nodes(model, D-1)` and the one-cell-around mask give in the reference.  `slab=(k0, k1)`: the local model of
    the slab."""
    from .cutters import make_grid
    L = _lib.lib()
    grid = make_grid(model, slab)
    nf, nvf = C.c_int64(0), C.c_int32(0)
    rc = L.gecut_facet_topology_host(C.byref(grid), 0, 0, None, None, C.byref(nf), C.byref(nvf))
    if rc != _lib.GECUT_OK:
        raise _lib.GecutError(f"gecut_facet_topology_host failed with status {rc}")
    nodes = np.empty((nf.value, nvf.value), np.int32)
    isb = np.empty(nf.value, np.int8)
    rc = L.gecut_facet_topology_host(C.byref(grid), 0, nf.value, C.c_void_p(nodes.ctypes.data), C.c_void_p(isb.ctypes.data),
                                     C.byref(nf), C.byref(nvf))
    if rc != _lib.GECUT_OK:
        raise _lib.GecutError(f"gecut_facet_topology_host failed with status {rc}")
    return nodes, isb
