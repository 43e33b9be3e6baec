"""ORACLE / TEST INFRASTRUCTURE -- ctypes binding of oracle/_build/libgecut_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgecut_oracle.so")

IN, OUT, CUT, INTERFACE = -1, 1, 0, 0


class OGrid(C.Structure):
    _fields_ = [("D", C.c_int32), ("simplexified", C.c_int32), ("pmin", C.c_double * 3), ("pmax", C.c_double * 3),
                ("partition", C.c_int32 * 3), ("pad", C.c_int32)]


class OLeaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("x0", C.c_double * 3), ("v", C.c_double * 3),
                ("R", C.c_double), ("r", C.c_double)]


_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "gecut_oracle.cpp")
    stale = (not os.path.exists(LIB_PATH)) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["bash", os.path.join(HERE, "build.sh")])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.gecut_oracle_cut.restype = C.c_void_p
        L.gecut_oracle_cut.argtypes = [C.POINTER(OGrid), C.POINTER(OLeaf), C.c_int, C.POINTER(C.c_void_p)]
        L.gecut_oracle_free.argtypes = [C.c_void_p]
        L.gecut_oracle_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.gecut_oracle_array.restype = C.c_void_p
        L.gecut_oracle_array.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        L.gecut_oracle_bgcell_to_inoutcut.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gecut_oracle_subcell_to_inout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gecut_oracle_select_subcells.restype = C.c_int64
        L.gecut_oracle_select_subcells.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gecut_oracle_select_bgcells.restype = C.c_int64
        L.gecut_oracle_select_bgcells.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gecut_oracle_select_boundary.restype = C.c_int64
        L.gecut_oracle_select_boundary.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                   C.c_void_p]
        L.gecut_oracle_subcell_measures.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gecut_oracle_subfacet_measures.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gecut_oracle_bgcell_measure.restype = C.c_double
        L.gecut_oracle_bgcell_measure.argtypes = [C.c_void_p, C.c_int32]
        L.gecut_oracle_cutcell_laplacian.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gecut_oracle_bgcell_laplacian.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.gecut_oracle_eval_leaf.argtypes = [C.POINTER(OLeaf), C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
        L.gecut_oracle_node_values.argtypes = [C.POINTER(OGrid), C.POINTER(OLeaf), C.c_void_p]
        L.gecut_oracle_node_coords.argtypes = [C.POINTER(OGrid), C.c_void_p]
        L.gecut_oracle_cut_facets.restype = C.c_void_p
        L.gecut_oracle_cut_facets.argtypes = [C.POINTER(OGrid), C.POINTER(OLeaf), C.c_int, C.POINTER(C.c_void_p)]
        L.gecut_oracle_facets_free.argtypes = [C.c_void_p]
        L.gecut_oracle_facets_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.gecut_oracle_facets_array.restype = C.c_void_p
        L.gecut_oracle_facets_array.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        L.gecut_oracle_bgfacet_to_inoutcut.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gecut_oracle_facets_subfacet_to_inout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gecut_oracle_select_facets.restype = C.c_int64
        L.gecut_oracle_select_facets.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.gecut_oracle_facets_subfacet_measures.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gecut_oracle_ghost_skeleton.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gecut_oracle_aggregate.restype = C.c_int
        L.gecut_oracle_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.gecut_oracle_bgfacet_measure.restype = C.c_double
        L.gecut_oracle_bgfacet_measure.argtypes = [C.c_void_p, C.c_int32]
        _lib = L
    return _lib


def make_grid(model) -> OGrid:
    g = OGrid()
    g.D = model.D
    g.simplexified = 1 if model.simplexified else 0
    for d in range(model.D):
        g.pmin[d] = model.pmin[d]
        g.pmax[d] = model.pmax[d]
        g.partition[d] = model.partition[d]
    return g


def make_leaf(kind, x0=(0, 0, 0), v=(0, 0, 0), R=0.0, r=0.0) -> OLeaf:
    lf = OLeaf()
    lf.kind = kind
    for d in range(3):
        lf.x0[d] = float(x0[d]) if d < len(x0) else 0.0
        lf.v[d] = float(v[d]) if d < len(v) else 0.0
    lf.R = float(R)
    lf.r = float(r)
    return lf


_DTYPES = {
    "ls_to_bgcell_to_inoutcut": np.int8, "cutcell_to_cell": np.int32,
    "subcells.cell_to_points": np.int32, "subcells.cell_to_bgcell": np.int32,
    "subcells.point_to_coords": np.float64, "subcells.point_to_rcoords": np.float64,
    "ls_to_subcell_to_inout": np.int8,
    "subfacets.facet_to_points": np.int32, "subfacets.facet_to_bgcell": np.int32,
    "subfacets.facet_to_normal": np.float64,
    "subfacets.point_to_coords": np.float64, "subfacets.point_to_rcoords": np.float64,
    "ls_to_subfacet_to_inout": np.int8,
}


class OracleCut:
    """Result of the oracle's `cut(LevelSetCutter(), model, geo)`: the arrays of the
    `EmbeddedDiscretization` (EmbeddedDiscretizations.jl:36-45) as numpy copies."""

    def __init__(self, model, leaves: Sequence[OLeaf], nodal_values: Optional[List[Optional[np.ndarray]]] = None):
        L = len(leaves)
        arr = (OLeaf * L)(*leaves)
        ptrs = (C.c_void_p * L)()
        self._keep = []
        for i in range(L):
            if nodal_values is not None and nodal_values[i] is not None:
                a = np.ascontiguousarray(nodal_values[i], dtype=np.float64)
                assert a.shape[0] == model.num_nodes
                self._keep.append(a)
                ptrs[i] = a.ctypes.data
            else:
                ptrs[i] = None
        self.model = model
        self.grid = make_grid(model)
        self.h = lib().gecut_oracle_cut(C.byref(self.grid), arr, L, ptrs)
        s = (C.c_int64 * 10)()
        lib().gecut_oracle_sizes(self.h, s)
        (self.Nc, self.Nn, self.Ncut, self.Nsc, self.Nscp, self.Nsf, self.Nsfp, self.L, self.D, self.nv) = [int(x) for x in s]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().gecut_oracle_free(self.h)
                self.h = None
        except Exception:
            pass

    def array(self, name: str, ls: int = 0) -> np.ndarray:
        n = C.c_int64()
        p = lib().gecut_oracle_array(self.h, name.encode(), ls, C.byref(n))
        if n.value < 0:
            raise KeyError(name)
        dt = np.dtype(_DTYPES[name])
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n.value * dt.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt).copy()

    def rows(self, name: str) -> np.ndarray:
        return np.stack([self.array(name, ls) for ls in range(self.L)]) if self.L else np.zeros((0, 0), np.int8)

    # selectors -------------------------------------------------------------------------
    @staticmethod
    def _prog(prog):
        a = np.ascontiguousarray(prog, dtype=np.int32)
        return a, a.ctypes.data, int(a.size)

    def bgcell_to_inoutcut(self, prog) -> np.ndarray:
        a, p, n = self._prog(prog)
        out = np.zeros(self.Nc, np.int8)
        lib().gecut_oracle_bgcell_to_inoutcut(self.h, p, n, out.ctypes.data)
        return out

    def subcell_to_inout(self, prog) -> np.ndarray:
        a, p, n = self._prog(prog)
        out = np.zeros(self.Nsc, np.int8)
        lib().gecut_oracle_subcell_to_inout(self.h, p, n, out.ctypes.data)
        return out

    def select_subcells(self, prog, side: int) -> np.ndarray:
        a, p, n = self._prog(prog)
        ids = np.zeros(max(self.Nsc, 1), np.int32)
        k = lib().gecut_oracle_select_subcells(self.h, p, n, side, ids.ctypes.data)
        return ids[:k].copy()

    def select_bgcells(self, prog, side: int, active: bool = False) -> np.ndarray:
        a, p, n = self._prog(prog)
        ids = np.zeros(max(self.Nc, 1), np.int32)
        k = lib().gecut_oracle_select_bgcells(self.h, p, n, side, 1 if active else 0, ids.ctypes.data)
        return ids[:k].copy()

    def select_boundary(self, prog1, prog2=None):
        a1, p1, n1 = self._prog(prog1)
        if prog2 is not None:
            a2, p2, n2 = self._prog(prog2)
        else:
            p2, n2 = None, 0
        ids = np.zeros(max(self.Nsf, 1), np.int32)
        ori = np.zeros(max(self.Nsf, 1), np.int8)
        k = lib().gecut_oracle_select_boundary(self.h, p1, n1, p2, n2, ids.ctypes.data, ori.ctypes.data)
        return ids[:k].copy(), ori[:k].copy()

    # integration -----------------------------------------------------------------------
    def subcell_measures(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(ids.size, np.float64)
        lib().gecut_oracle_subcell_measures(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def subfacet_measures(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(ids.size, np.float64)
        lib().gecut_oracle_subfacet_measures(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def bgcell_measure(self, cell1: int) -> float:
        return float(lib().gecut_oracle_bgcell_measure(self.h, int(cell1)))

    def cutcell_laplacian(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int32)
        K = np.zeros((self.Ncut, self.nv, self.nv), np.float64)
        lib().gecut_oracle_cutcell_laplacian(self.h, ids.ctypes.data, ids.size, K.ctypes.data)
        return K

    def bgcell_laplacian(self, cell1: int) -> np.ndarray:
        K = np.zeros((self.nv, self.nv), np.float64)
        lib().gecut_oracle_bgcell_laplacian(self.h, int(cell1), K.ctypes.data)
        return K


_FACET_DTYPES = {
    "ls_to_facet_to_inoutcut": np.int8, "cutfacet_to_facet": np.int32, "facet_is_boundary": np.int8,
    "facet_nodes": np.int64,
    "subfacets.cell_to_points": np.int32, "subfacets.cell_to_bgcell": np.int32,
    "subfacets.point_to_coords": np.float64, "subfacets.point_to_rcoords": np.float64,
    "ls_to_subfacet_to_inout": np.int8,
}

BOUNDARY_FACETS, SKELETON_FACETS = 0, 1
SEL_CUT, SEL_BG, SEL_ACTIVE = 0, 1, 2


class OracleFacetCut:
    """Result of the oracle's `cut_facets(LevelSetCutter(), model, geo)`: the arrays of the
    `EmbeddedFacetDiscretization` (EmbeddedFacetDiscretizations.jl:28-35) as numpy copies."""

    def __init__(self, model, leaves: Sequence[OLeaf], nodal_values: Optional[List[Optional[np.ndarray]]] = None):
        L = len(leaves)
        arr = (OLeaf * L)(*leaves)
        ptrs = (C.c_void_p * L)()
        self._keep = []
        for i in range(L):
            if nodal_values is not None and nodal_values[i] is not None:
                a = np.ascontiguousarray(nodal_values[i], dtype=np.float64)
                assert a.shape[0] == model.num_nodes
                self._keep.append(a)
                ptrs[i] = a.ctypes.data
            else:
                ptrs[i] = None
        self.model = model
        self.grid = make_grid(model)
        self.h = lib().gecut_oracle_cut_facets(C.byref(self.grid), arr, L, ptrs)
        s = (C.c_int64 * 8)()
        lib().gecut_oracle_facets_sizes(self.h, s)
        (self.Nf, self.Ncutf, self.Nsf, self.Nsfp, self.L, self.D, self.nvf, self.Nboundary) = [int(x) for x in s]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().gecut_oracle_facets_free(self.h)
                self.h = None
        except Exception:
            pass

    def array(self, name: str, ls: int = 0) -> np.ndarray:
        n = C.c_int64()
        p = lib().gecut_oracle_facets_array(self.h, name.encode(), ls, C.byref(n))
        if n.value < 0:
            raise KeyError(name)
        dt = np.dtype(_FACET_DTYPES[name])
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n.value * dt.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt).copy()

    def rows(self, name: str) -> np.ndarray:
        return np.stack([self.array(name, ls) for ls in range(self.L)])

    def bgfacet_to_inoutcut(self, prog) -> np.ndarray:
        a, p, n = OracleCut._prog(prog)
        out = np.zeros(self.Nf, np.int8)
        lib().gecut_oracle_bgfacet_to_inoutcut(self.h, p, n, out.ctypes.data)
        return out

    def subfacet_to_inout(self, prog) -> np.ndarray:
        a, p, n = OracleCut._prog(prog)
        out = np.zeros(self.Nsf, np.int8)
        lib().gecut_oracle_facets_subfacet_to_inout(self.h, p, n, out.ctypes.data)
        return out

    def select(self, prog, which: int, mode: int, side: int) -> np.ndarray:
        a, p, n = OracleCut._prog(prog)
        ids = np.zeros(max(self.Nsf, self.Nf, 1), np.int32)
        k = lib().gecut_oracle_select_facets(self.h, p, n, which, mode, side, ids.ctypes.data)
        return ids[:k].copy()

    def subfacet_measures(self, ids) -> np.ndarray:
        ids = np.ascontiguousarray(ids, np.int32)
        out = np.zeros(ids.size, np.float64)
        lib().gecut_oracle_facets_subfacet_measures(self.h, ids.ctypes.data, ids.size, out.ctypes.data)
        return out

    def bgfacet_measure(self, facet1: int) -> float:
        return float(lib().gecut_oracle_bgfacet_measure(self.h, int(facet1)))


def oracle_aggregate(oc: "OracleCut", fc: "OracleFacetCut", prog, loc: int = IN, threshold: float = 1.0):
    """`aggregate(AggregateCutCellsByThreshold(threshold), cut, cut_facets, geo, loc)` (CellAggregation.jl:83-189):
    returns (cell_to_cellin Int32 1-based, all_aggregated)."""
    a, p, n = OracleCut._prog(prog)
    out = np.zeros(oc.Nc, np.int32)
    rc = lib().gecut_oracle_aggregate(oc.h, fc.h, p, n, int(loc), float(threshold), out.ctypes.data)
    return out, rc == 0


def oracle_ghost_skeleton(oc: "OracleCut", fc: "OracleFacetCut", prog, in_or_out: int = IN) -> np.ndarray:
    """`facet_to_mask` of `GhostSkeleton(cut, ActiveInOrOut(in_or_out), geo)` (EmbeddedDiscretizations.jl:504-543)."""
    a, p, n = OracleCut._prog(prog)
    out = np.zeros(fc.Nf, np.int8)
    lib().gecut_oracle_ghost_skeleton(oc.h, fc.h, p, n, int(in_or_out), out.ctypes.data)
    return out.astype(bool)


def oracle_cut_facets(model, geo) -> "OracleFacetCut":
    import importlib
    gecut = importlib.import_module("gecut")
    leaves, nodal, oid_to_ls = gecut.cutters.flatten_geometry(model, geo, host_only=True)
    oleaves = [make_leaf(l.kind, l.x0, l.v, l.R, l.r) for l in leaves]
    oc = OracleFacetCut(model, oleaves, nodal)
    oc.oid_to_ls = oid_to_ls
    oc.geo = geo
    return oc


def node_values(model, leaf: OLeaf) -> np.ndarray:
    out = np.zeros(model.num_nodes, np.float64)
    g = make_grid(model)
    lib().gecut_oracle_node_values(C.byref(g), C.byref(leaf), out.ctypes.data)
    return out


def node_coords(model) -> np.ndarray:
    out = np.zeros((model.num_nodes, model.D), np.float64)
    g = make_grid(model)
    lib().gecut_oracle_node_coords(C.byref(g), out.ctypes.data)
    return out


def oracle_cut(model, geo) -> "OracleCut":
    """Convenience: flatten a host-side geometry (product mirror classes) and cut with the oracle.
    Returns (OracleCut, oid_to_ls)."""
    import importlib
    gecut = importlib.import_module("gecut")
    leaves, nodal, oid_to_ls = gecut.cutters.flatten_geometry(model, geo, host_only=True)
    oleaves = [make_leaf(l.kind, l.x0, l.v, l.R, l.r) for l in leaves]
    oc = OracleCut(model, oleaves, nodal)
    oc.oid_to_ls = oid_to_ls
    oc.geo = geo
    return oc
