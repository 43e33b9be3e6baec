"""`EmbeddedDiscretization`, `SubCellData`, `SubFacetData` and the triangulation selectors of the host mirror.

Reference: `src/Interfaces/Interfaces.jl:70-89` (state constants and selector types),
`src/Interfaces/EmbeddedDiscretizations.jl:36-45,91-105,255-339,408-463`,
`src/Interfaces/SubCellTriangulations.jl:2-17`, `src/Interfaces/SubFacetTriangulations.jl:2-21`.

The arrays live on the device inside a `gecut_result`; the properties below copy them to the host on first
access (numpy arrays with the reference's element types: Int32 1-based ids, Int8 states, Float64 AoS points).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib
from .csg import Geometry, compile_postfix, get_geometry

IN, OUT, INTERFACE, CUT = -1, 1, 0, 0  # Interfaces.jl:70-73


class CutInOrOut:  # Interfaces.jl:75-77
    def __init__(self, in_or_out: int):
        self.in_or_out = in_or_out

    def __repr__(self):
        return f"CutInOrOut({self.in_or_out})"


class ActiveInOrOut:  # Interfaces.jl:84-86
    def __init__(self, in_or_out: int):
        self.in_or_out = in_or_out

    def __repr__(self):
        return f"ActiveInOrOut({self.in_or_out})"


CUT_IN, CUT_OUT = CutInOrOut(IN), CutInOrOut(OUT)
PHYSICAL_IN, PHYSICAL_OUT = (CUT_IN, IN), (CUT_OUT, OUT)
PHYSICAL = PHYSICAL_IN
ACTIVE_IN, ACTIVE_OUT = ActiveInOrOut(IN), ActiveInOrOut(OUT)
ACTIVE = ACTIVE_IN


class SubCellData:
    """`SubCellData{Dr,Dp,Tp,Tr}` (`SubCellTriangulations.jl:2-7`).  `cell_to_points` is the Table's data reshaped
    to (ncells, D+1); `cell_to_points_ptrs` gives the 1-based ptrs vector."""

    def __init__(self, cell_to_points, cell_to_bgcell, point_to_coords, point_to_rcoords):
        self.cell_to_points = cell_to_points
        self.cell_to_bgcell = cell_to_bgcell
        self.point_to_coords = point_to_coords
        self.point_to_rcoords = point_to_rcoords

    @property
    def cell_to_points_ptrs(self) -> np.ndarray:
        n, s = self.cell_to_points.shape
        return (1 + s * np.arange(n + 1, dtype=np.int64)).astype(np.int32)


class SubFacetData:
    """`SubFacetData{Dp,T,Tn}` (`SubFacetTriangulations.jl:2-8`)."""

    def __init__(self, facet_to_points, facet_to_normal, facet_to_bgcell, point_to_coords, point_to_rcoords):
        self.facet_to_points = facet_to_points
        self.facet_to_normal = facet_to_normal
        self.facet_to_bgcell = facet_to_bgcell
        self.point_to_coords = point_to_coords
        self.point_to_rcoords = point_to_rcoords

    @property
    def facet_to_points_ptrs(self) -> np.ndarray:
        n, s = self.facet_to_points.shape
        return (1 + s * np.arange(n + 1, dtype=np.int64)).astype(np.int32)


def _alloc(shape, dtype, pinned: bool):
    if pinned:
        import torch
        tdt = {np.int8: torch.int8, np.int32: torch.int32, np.float64: torch.float64}[dtype]
        return torch.empty(shape, dtype=tdt, pin_memory=True).numpy()
    return np.empty(shape, dtype=dtype)


class EmbeddedDiscretization:
    """`EmbeddedDiscretization{Dc,T}` (`EmbeddedDiscretizations.jl:36-45`): `bgmodel`,
    `ls_to_bgcell_to_inoutcut`, `subcells`, `ls_to_subcell_to_inout`, `subfacets`, `ls_to_subfacet_to_inout`,
    `oid_to_ls`, `geo` -- device resident, copied to the host lazily."""

    def __init__(self, ctx: _lib.Context, handle, bgmodel, geo: Geometry, oid_to_ls: dict):
        self.ctx = ctx
        self._h = handle
        self.bgmodel = bgmodel
        self.geo = geo
        self.oid_to_ls = oid_to_ls
        s = _lib.Sizes()
        ctx.check(ctx._lib.gecut_result_sizes(self._h, C.byref(s)), "gecut_result_sizes")
        self.sizes = s
        self._host = None

    # ---- sizes ---------------------------------------------------------------------------
    @property
    def D(self):
        return int(self.sizes.D)

    @property
    def L(self):
        return int(self.sizes.num_levelsets)

    @property
    def num_bgcells(self):
        return int(self.sizes.num_bgcells)

    @property
    def num_cutcells(self):
        return int(self.sizes.num_cutcells)

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            self.ctx._lib.gecut_result_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- device access ----------------------------------------------------------------------
    def device_ptrs(self) -> _lib.Buffers:
        b = _lib.Buffers()
        self.ctx.check(self.ctx._lib.gecut_result_device_ptrs(self._h, C.byref(b)), "gecut_result_device_ptrs")
        return b

    # ---- host copy ------------------------------------------------------------------------------
    def to_host(self, pinned: bool = False, buffers: Optional[dict] = None, share_aliased_points: bool = True) -> dict:
        """Copy the whole layout to the host (the arrays a Julia `EmbeddedDiscretization` owns)."""
        if self._host is not None and buffers is None:
            return self._host
        s = self.sizes
        D, L = int(s.D), int(s.num_levelsets)
        shapes = {
            "ls_to_bgcell_to_inoutcut": ((L, s.num_bgcells), np.int8),
            "cutcell_to_cell": ((s.num_cutcells,), np.int32),
            "sc_cell_to_points": ((s.num_subcells, D + 1), np.int32),
            "sc_cell_to_bgcell": ((s.num_subcells,), np.int32),
            "sc_point_to_coords": ((s.num_subcell_points, D), np.float64),
            "sc_point_to_rcoords": ((s.num_subcell_points, D), np.float64),
            "ls_to_subcell_to_inout": ((L, s.num_subcells), np.int8),
            "sf_facet_to_points": ((s.num_subfacets, D), np.int32),
            "sf_facet_to_bgcell": ((s.num_subfacets,), np.int32),
            "sf_facet_to_normal": ((s.num_subfacets, D), np.float64),
            "sf_point_to_coords": ((s.num_subfacet_points, D), np.float64),
            "sf_point_to_rcoords": ((s.num_subfacet_points, D), np.float64),
            "ls_to_subfacet_to_inout": ((L, s.num_subfacets), np.int8),
        }
        host = buffers if buffers is not None else {}
        b = _lib.Buffers()
        # one level set: the subfacet point arrays are the subcell point arrays (the reference aliases them until
        # merge_sub_face_data copies them); share ONE host array for both instead of copying 2 x Nscp x D doubles twice
        alias = bool(s.subfacet_points_alias) and share_aliased_points
        skip = {"sf_point_to_coords": "sc_point_to_coords", "sf_point_to_rcoords": "sc_point_to_rcoords"} if alias else {}
        for name, (shape, dt) in shapes.items():
            if name in skip:
                continue
            if name not in host:
                host[name] = _alloc(tuple(int(x) for x in shape), dt, pinned)
            a = host[name]
            setattr(b, name, a.ctypes.data if a.size else None)
        self.ctx.check(self.ctx._lib.gecut_result_copy(self._h, C.byref(b)), "gecut_result_copy")
        for name, src in skip.items():
            host[name] = host[src]
        if buffers is None:
            self._host = host
        return host

    @property
    def ls_to_bgcell_to_inoutcut(self) -> np.ndarray:
        return self.to_host()["ls_to_bgcell_to_inoutcut"]

    @property
    def cutcell_to_cell(self) -> np.ndarray:
        return self.to_host()["cutcell_to_cell"]

    @property
    def subcells(self) -> SubCellData:
        h = self.to_host()
        return SubCellData(h["sc_cell_to_points"], h["sc_cell_to_bgcell"], h["sc_point_to_coords"],
                           h["sc_point_to_rcoords"])

    @property
    def ls_to_subcell_to_inout(self) -> np.ndarray:
        return self.to_host()["ls_to_subcell_to_inout"]

    @property
    def subfacets(self) -> SubFacetData:
        h = self.to_host()
        return SubFacetData(h["sf_facet_to_points"], h["sf_facet_to_normal"], h["sf_facet_to_bgcell"],
                            h["sf_point_to_coords"], h["sf_point_to_rcoords"])

    @property
    def ls_to_subfacet_to_inout(self) -> np.ndarray:
        return self.to_host()["ls_to_subfacet_to_inout"]

    # ---- programs ------------------------------------------------------------------------------
    def program(self, geo_or_name: Union[None, str, Geometry]) -> np.ndarray:
        """Postfix program of `geo` (or of the named sub-geometry, `CSG/Geometries.jl:128-154`) over this
        discretization's level-set rows (`oid_to_ls`, `EmbeddedDiscretizations.jl:95-101`)."""
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


def get_background_model(cut: EmbeddedDiscretization):
    return cut.bgmodel


def compute_bgcell_to_inoutcut(cut: EmbeddedDiscretization, geo=None) -> np.ndarray:
    """`compute_bgcell_to_inoutcut(cut,geo)` (`EmbeddedDiscretizations.jl:91-105`)."""
    prog = cut.program(geo)
    out = np.empty(cut.num_bgcells, dtype=np.int8)
    cut.ctx.check(cut.ctx._lib.gecut_bgcell_to_inoutcut(cut.handle, prog.ctypes.data, prog.size, out.ctypes.data),
                  "gecut_bgcell_to_inoutcut")
    return out


def compute_subcell_to_inout(cut: EmbeddedDiscretization, geo=None) -> np.ndarray:
    """`compute_subcell_to_inout(cut,geo)` (`EmbeddedDiscretizations.jl:325-339`)."""
    prog = cut.program(geo)
    out = np.empty(int(cut.sizes.num_subcells), dtype=np.int8)
    cut.ctx.check(cut.ctx._lib.gecut_subcell_to_inout(cut.handle, prog.ctypes.data, prog.size, out.ctypes.data),
                  "gecut_subcell_to_inout")
    return out


class CutTriangulation:
    """Result of `Triangulation(cut, in_or_out, geo)` / `EmbeddedBoundary(cut, geo1[, geo2])`: a device-resident
    `gecut_selection`.  `kind` is "physical", "cut", "bg", "active" or "boundary"; the parts the reference would
    `lazy_append` (`EmbeddedDiscretizations.jl:297-301`) are `sub_ids` (subcells / subfacets, 1-based, stable
    order) and `bg_ids` (uncut bg cells)."""

    def __init__(self, cut: EmbeddedDiscretization, handle, kind: str, use_sub: bool, use_bg: bool):
        self.cut = cut
        self.ctx = cut.ctx
        self._h = handle
        self.kind = kind
        self.use_sub = use_sub
        self.use_bg = use_bg
        a, b = C.c_int64(), C.c_int64()
        self.ctx.check(self.ctx._lib.gecut_selection_sizes(self._h, C.byref(a), C.byref(b)), "gecut_selection_sizes")
        self.num_sub = int(a.value) if use_sub else 0
        self.num_bg = int(b.value) if use_bg else 0
        self._n_sub_raw, self._n_bg_raw = int(a.value), int(b.value)
        self._ids = None

    @property
    def handle(self):
        return self._h

    def num_cells(self) -> int:
        return self.num_sub + self.num_bg

    def close(self):
        if self._h:
            self.ctx._lib.gecut_selection_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _copy(self):
        if self._ids is None:
            sub = np.empty(self._n_sub_raw, np.int32)
            ori = np.empty(self._n_sub_raw, np.int8)
            bg = np.empty(self._n_bg_raw, np.int32)
            self.ctx.check(self.ctx._lib.gecut_selection_copy(
                self._h, sub.ctypes.data if sub.size else None,
                ori.ctypes.data if (ori.size and self.kind == "boundary") else None,
                bg.ctypes.data if bg.size else None), "gecut_selection_copy")
            self._ids = (sub, ori, bg)
        return self._ids

    @property
    def sub_ids(self) -> np.ndarray:
        return self._copy()[0] if self.use_sub else np.zeros(0, np.int32)

    @property
    def orientation(self) -> np.ndarray:
        return self._copy()[1]

    @property
    def bg_ids(self) -> np.ndarray:
        return self._copy()[2] if self.use_bg else np.zeros(0, np.int32)

    def gather(self):
        """`SubCellData(st,newcells)` (`SubCellTriangulations.jl:9-17`) or `SubFacetData(st,newfacets,orientation)`
        (`SubFacetTriangulations.jl:10-21`): re-indexed table + bgcell (+ oriented normals); points are shared."""
        D = self.cut.D
        n = self._n_sub_raw
        if self.kind == "boundary":
            tp = np.empty((n, D), np.int32)
            tb = np.empty(n, np.int32)
            nr = np.empty((n, D), np.float64)
            self.ctx.check(self.ctx._lib.gecut_selection_gather(
                self._h, tp.ctypes.data if n else None, tb.ctypes.data if n else None, nr.ctypes.data if n else None),
                "gecut_selection_gather")
            sf = self.cut.subfacets
            return SubFacetData(tp, nr, tb, sf.point_to_coords, sf.point_to_rcoords)
        tp = np.empty((n, D + 1), np.int32)
        tb = np.empty(n, np.int32)
        self.ctx.check(self.ctx._lib.gecut_selection_gather(
            self._h, tp.ctypes.data if n else None, tb.ctypes.data if n else None, None), "gecut_selection_gather")
        sc = self.cut.subcells
        return SubCellData(tp, tb, sc.point_to_coords, sc.point_to_rcoords)


def Triangulation(cut: EmbeddedDiscretization, in_or_out=PHYSICAL_IN, geo=None) -> CutTriangulation:
    """`Triangulation(cut[,in_or_out][,geo or name])` (`EmbeddedDiscretizations.jl:255-323`)."""
    if isinstance(in_or_out, (str, Geometry)) and geo is None:  # Triangulation(cut,name) / Triangulation(cut,geo)
        in_or_out, geo = PHYSICAL_IN, in_or_out
    prog = cut.program(geo)
    lib = cut.ctx._lib
    out = C.c_void_p()
    if isinstance(in_or_out, tuple):
        a, b = in_or_out
        if not (isinstance(a, CutInOrOut) and isinstance(b, int) and a.in_or_out == b):
            raise NotImplementedError("only (CutInOrOut(s), s) tuples (PHYSICAL_IN / PHYSICAL_OUT) are supported")
        kind, side, use = "physical", b, (True, True)
        sel_kind = _lib.SEL_PHYSICAL
    elif isinstance(in_or_out, CutInOrOut):
        kind, side, use = "cut", in_or_out.in_or_out, (True, False)
        sel_kind = _lib.SEL_CUTONLY
    elif isinstance(in_or_out, ActiveInOrOut):
        kind, side, use = "active", in_or_out.in_or_out, (False, True)
        sel_kind = _lib.SEL_ACTIVE
    elif isinstance(in_or_out, int):
        kind, side, use = "bg", in_or_out, (False, True)
        sel_kind = _lib.SEL_BGONLY
    else:
        raise TypeError(f"unsupported selector {in_or_out!r}")
    if side not in (IN, OUT):
        raise NotImplementedError("side must be IN or OUT")
    cut.ctx.check(lib.gecut_select_cells(cut.handle, sel_kind, prog.ctypes.data, prog.size, side, C.byref(out)),
                  "gecut_select_cells")
    return CutTriangulation(cut, out, kind, *use)


def EmbeddedBoundary(cut: EmbeddedDiscretization, geo1=None, geo2=None) -> CutTriangulation:
    """`EmbeddedBoundary(cut[,geo1[,geo2]])` (`EmbeddedDiscretizations.jl:408-463`)."""
    p1 = cut.program(geo1)
    out = C.c_void_p()
    if geo2 is None:
        rc = cut.ctx._lib.gecut_select_boundary(cut.handle, p1.ctypes.data, p1.size, None, 0, C.byref(out))
    else:
        p2 = cut.program(geo2)
        rc = cut.ctx._lib.gecut_select_boundary(cut.handle, p1.ctypes.data, p1.size, p2.ctypes.data, p2.size,
                                                C.byref(out))
    cut.ctx.check(rc, "gecut_select_boundary")
    return CutTriangulation(cut, out, "boundary", True, False)
