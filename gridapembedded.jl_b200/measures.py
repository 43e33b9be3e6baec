"""`Measure(trian, degree)` and the integrals the hot path produces on the device.

Reference call sites: `examples/PoissonCSGCutFEM/PoissonCSGCutFEM.jl:47-51,63-70`,
`test/InterfacesTests/EmbeddedDiscretizationsTests.jl:48-56`.  Gridap's `Measure/CellQuadrature/∫` engine is
external to the reference; its semantics for this path are restated in SURVEY.md section 3.4: the degree-2
reference rule of the sub-simplex mapped through its P1 map, `w_q*|det J_s|`, background shape functions evaluated
at the glued reference point.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .discretizations import CutTriangulation


class Measure:
    """`Measure(trian, degree)`; only `degree == 2` (= `2*order` for first-order spaces, the value every example of
    the reference uses) is implemented on the device."""

    def __init__(self, trian: CutTriangulation, degree: int = 2):
        if degree != 2:
            raise NotImplementedError("only degree-2 quadrature is implemented on the device")
        self.trian = trian
        self.degree = degree


def integrate_measure(dΩ: Measure, return_values: bool = False):
    """`sum(∫(1)dΩ)` (volume of a PHYSICAL triangulation, or surface of an EmbeddedBoundary).
    With `return_values` also the per-sub-entity contributions (selection order)."""
    t = dΩ.trian
    total = C.c_double()
    vals = np.empty(t._n_sub_raw, np.float64) if return_values else None
    t.ctx.check(t.ctx._lib.gecut_integrate_measure(
        t.handle, vals.ctypes.data if (vals is not None and vals.size) else None, C.byref(total)),
        "gecut_integrate_measure")
    if return_values:
        return float(total.value), vals
    return float(total.value)


def integrate_laplacian(dΩ: Measure) -> Tuple[np.ndarray, np.ndarray]:
    """Element matrices of `∫(∇v⋅∇u)dΩ` for the bg shape functions: `(K_cut, K_bg)` with `K_cut[k]` the
    contribution of the selected subcells of cut cell `cutcell_to_cell[k]` (Ncut x nv x nv) and `K_bg[type]` the
    matrix of an uncut cell (ntypes x nv x nv)."""
    t = dΩ.trian
    cut = t.cut
    nv = int(cut.sizes.num_vertices_per_bgcell)
    model = cut.bgmodel
    ntypes = model.simplices_per_cube if model.simplexified else 1
    K_cut = np.empty((cut.num_cutcells, nv, nv), np.float64)
    K_bg = np.empty((ntypes, nv, nv), np.float64)
    t.ctx.check(t.ctx._lib.gecut_integrate_laplacian(
        t.handle, K_cut.ctypes.data if K_cut.size else None, K_bg.ctypes.data), "gecut_integrate_laplacian")
    return K_cut, K_bg


def get_cell_measure(trian: CutTriangulation, device_output: bool = False):
    """`get_cell_measure(trian, bgtrian)` (`src/AgFEM/CellAggregation.jl:111-113`): measure of `trian` per background
    cell (slab-local numbering) -- the input of the AgFEM aggregation.  numpy array, or a CUDA torch tensor."""
    nc = int(trian.cut.sizes.num_bgcells)
    if device_output:
        import torch
        out = torch.empty(nc, dtype=torch.float64, device=f"cuda:{trian.ctx.device}")
        trian.ctx.check(trian.ctx._lib.gecut_cell_measure(trian.handle, C.c_void_p(out.data_ptr()), 1), "gecut_cell_measure")
        return out
    out = np.empty(nc, np.float64)
    trian.ctx.check(trian.ctx._lib.gecut_cell_measure(trian.handle, C.c_void_p(out.ctypes.data), 0), "gecut_cell_measure")
    return out
