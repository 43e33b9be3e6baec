"""numpy helpers for the cut_facets tests (CPU oracle and GPU parity): facet -> cell adjacency of a Cartesian model
from the facet node lists, outward normals, and the integrands of the reference's known-answer checks
(test/InterfacesTests/EmbeddedFacetDiscretizationsTests.jl:18-68)."""
import numpy as np

import helpers as H


def facet_geometry(model, facet_nodes):
    """facet_nodes (Nf, nvf) 0-based node ids -> (direction d the facet is normal to, node index of the facet along d,
    lower-corner cell multi-index of the facet in the other directions)."""
    D = model.D
    nn = [n + 1 for n in model.partition]
    ijk = np.zeros(facet_nodes.shape + (D,), dtype=np.int64)
    rem = facet_nodes.copy()
    for d in range(D):
        ijk[..., d] = rem % nn[d]
        rem = rem // nn[d]
    const = np.all(ijk == ijk[:, :1, :], axis=1)  # (Nf, D): coordinate shared by all nodes of the facet
    assert np.all(const.sum(1) == 1)
    d = np.argmax(const, axis=1)
    lo = ijk.min(axis=1)  # (Nf, D)
    return d, lo


def facet_cells(model, facet_nodes):
    """(minus, plus): 0-based ids of the cell on the low / high side of every facet along its normal (-1: none)."""
    D = model.D
    n = model.partition
    d, lo = facet_geometry(model, facet_nodes)
    Nf = len(d)
    stride = np.cumprod([1] + list(n[:-1]))
    ar = np.arange(Nf)
    pos = lo[ar, d]
    base = lo.copy()
    base[ar, d] = 0
    lin = (base * stride[None, :]).sum(1)
    minus = np.where(pos > 0, lin + (pos - 1) * stride[d], -1)
    plus = np.where(pos < np.asarray(n)[d], lin + pos * stride[d], -1)
    return d, pos, minus, plus


def bg_field_at_x(model, x, vn):
    """Q1 bg FE function with nodal values vn at physical points x (N,D) of an n-cube Cartesian model."""
    D = model.D
    h = np.asarray(model.sizes)
    lo = np.asarray(model.pmin)
    n = np.asarray(model.partition)
    t = (x - lo) / h
    c = np.clip(np.floor(t).astype(np.int64), 0, n - 1)
    eta = t - c
    nn = n + 1
    v = np.zeros(len(x))
    for a in range(2 ** D):
        w = np.ones(len(x))
        node = np.zeros(len(x), dtype=np.int64)
        stride = 1
        for d in range(D):
            bit = (a >> d) & 1
            w *= eta[:, d] if bit else 1 - eta[:, d]
            node += (c[:, d] + bit) * stride
            stride *= nn[d]
        v += w * vn[node]
    return v


def integrate_fun_on_subfacets(c2p, X, fun, degree=2):
    """sum over the (D-1)-simplices (c2p rows, 1-based ids into X) of ∫ fun(x) dS."""
    if len(c2p) == 0:
        return 0.0
    ds = c2p.shape[1] - 1
    xi, w = H.simplex_rule(ds, degree)
    N = H.p1_shape(xi)
    Xc = X[c2p - 1]
    meas = H.simplex_meas(Xc)
    total = 0.0
    for q in range(len(w)):
        x = np.einsum("i,nid->nd", N[q], Xc)
        total += (w[q] * meas * fun(x)).sum()
    return total


def integrate_fun_on_bgfacets(model, facet_nodes, fun):
    """sum over whole bg facets (rows of 0-based node ids) of ∫ fun(x) dS with the tensor Gauss rule of degree 2."""
    if len(facet_nodes) == 0:
        return 0.0
    D = model.D
    Xn = model.node_coordinates()[facet_nodes]  # (N, nvf, D)
    if D == 2:
        e = Xn[:, 1] - Xn[:, 0]
        meas = np.sqrt((e ** 2).sum(1))
        total = 0.0
        for q in range(2):
            x = Xn[:, 0] + H.GAUSS2[q] * e
            total += (0.5 * meas * fun(x)).sum()
        return total
    a = Xn[:, 1] - Xn[:, 0]
    b = Xn[:, 2] - Xn[:, 0]
    meas = np.sqrt((np.cross(a, b) ** 2).sum(1))
    total = 0.0
    for q in range(4):
        x = Xn[:, 0] + H.GAUSS2[q & 1] * a + H.GAUSS2[q >> 1] * b
        total += (0.25 * meas * fun(x)).sum()
    return total


def boundary_flux(model, fc, prog, vn, gradu, side=-1):
    """∫_{Γf} v n·∇u over BoundaryTriangulation(cutgeo_facets, PHYSICAL): the IN sub-facets of the cut boundary
    facets plus the IN boundary facets, n = outward normal of the model.  `fc` is an (Oracle)FacetCut-like object with
    .array/.select; ids are 1-based."""
    import oracle_binding as ob
    D = model.D
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    d, pos, minus, plus = facet_cells(model, fn)
    sign = np.where(minus < 0, -1.0, 1.0)  # low-side boundary facets look towards -e_d
    X = fc.array("subfacets.point_to_coords").reshape(-1, D)
    c2p = fc.array("subfacets.cell_to_points").reshape(-1, D)
    c2bg = fc.array("subfacets.cell_to_bgcell")
    sub = fc.select(prog, ob.BOUNDARY_FACETS, ob.SEL_CUT, side)
    bg = fc.select(prog, ob.BOUNDARY_FACETS, ob.SEL_BG, side)
    total = 0.0
    for dd in range(D):
        for sg in (-1.0, 1.0):
            k = sub[(d[c2bg[sub - 1] - 1] == dd) & (sign[c2bg[sub - 1] - 1] == sg)]
            total += sg * gradu[dd] * integrate_fun_on_subfacets(c2p[k - 1], X, lambda x: bg_field_at_x(model, x, vn))
            kb = bg[(d[bg - 1] == dd) & (sign[bg - 1] == sg)]
            total += sg * gradu[dd] * integrate_fun_on_bgfacets(model, fn[kb - 1], lambda x: bg_field_at_x(model, x, vn))
    return total


def cellwise_divergence_defect(model, oc, fc, p_cells, p_facets, measures_sub=None, measures_bgfacet=None):
    """Per bg cell K: D*|K∩Ω| - [ Σ_{facets F of K} ± x_d |F∩Ω| + ∫_{Γ∩K} x·n ]  (divergence theorem for F(x) = x on
    K∩Ω), which ties the IN part of EVERY bg facet (interior ones too) from cut_facets to the subcells and the
    embedded boundary from cut().  Returns the vector of defects (one per bg cell)."""
    import oracle_binding as ob
    D = model.D
    Nc = model.num_cells
    # volumes
    vol = np.zeros(Nc)
    ids = oc.select_subcells(p_cells, -1)
    c2bg = oc.array("subcells.cell_to_bgcell")
    np.add.at(vol, c2bg[ids - 1] - 1, oc.subcell_measures(ids))
    bg = oc.select_bgcells(p_cells, -1)
    cellvol = float(np.prod(model.sizes))
    vol[bg - 1] += cellvol
    # embedded boundary: ∫ x·n per subfacet (x·n is constant on a flat facet; use the first vertex)
    fids, ori = oc.select_boundary(p_cells)
    fX = oc.array("subfacets.point_to_coords").reshape(-1, D)
    f2p = oc.array("subfacets.facet_to_points").reshape(-1, D)
    f2bg = oc.array("subfacets.facet_to_bgcell")
    f2n = oc.array("subfacets.facet_to_normal").reshape(-1, D)
    nrm = f2n[fids - 1] * ori[:, None]
    x0 = fX[f2p[fids - 1][:, 0] - 1]
    flux = np.zeros(Nc)
    np.add.at(flux, f2bg[fids - 1] - 1, (x0 * nrm).sum(1) * oc.subfacet_measures(fids))
    # bg facets: IN area of every facet
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    d, pos, minus, plus = facet_cells(model, fn)
    area = np.zeros(fc.Nf)
    for which in (ob.BOUNDARY_FACETS, ob.SKELETON_FACETS):
        sub = fc.select(p_facets, which, ob.SEL_CUT, -1)
        sc2bg = fc.array("subfacets.cell_to_bgcell")
        np.add.at(area, sc2bg[sub - 1] - 1, fc.subfacet_measures(sub))
        whole = fc.select(p_facets, which, ob.SEL_BG, -1)
        h = np.asarray(model.sizes)
        full = np.array([np.prod(np.delete(h, dd)) for dd in range(D)])
        area[whole - 1] += full[d[whole - 1]]
    xd = np.asarray(model.pmin)[d] + pos * np.asarray(model.sizes)[d]
    m = minus >= 0
    np.add.at(flux, minus[m], (xd * area)[m])   # facet is the HIGH face of its minus cell: n = +e_d
    m = plus >= 0
    np.add.at(flux, plus[m], -(xd * area)[m])   # LOW face of its plus cell: n = -e_d
    return D * vol - flux
