"""Pins the CPU oracle's `cut_facets` against the known answers the reference holds for it
(test/InterfacesTests/EmbeddedFacetDiscretizationsTests.jl:18-68: divergence theorem with the physical part of the
model boundary) and against `cut()` itself through the cell-wise divergence theorem."""
import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, disk, sphere, cylinder, union, intersect, setdiff, complement, get_geometry,
                   plane)
from gecut.csg import compile_postfix
import oracle_binding as ob
import helpers as H
import facet_helpers as FH
from test_oracle_kat import divergence_theorem, csg_geometry

IN, OUT, CUT = -1, 1, 0


def both(model, geo):
    oc = ob.oracle_cut(model, geo)
    fc = ob.oracle_cut_facets(model, geo)
    assert fc.oid_to_ls == oc.oid_to_ls or list(fc.oid_to_ls.values()) == list(oc.oid_to_ls.values())
    return oc, fc


def prog(oc, name=None):
    geo = oc.geo if name is None else get_geometry(oc.geo, name)
    return compile_postfix(geo.get_tree(), oc.oid_to_ls)


def test_facet_numbering_counts():
    # Gridap: num_facets of a Cartesian model; boundary facets = facets with one cell around
    m2 = CartesianDiscreteModel((0, 1, 0, 1), (4, 3))
    fc = ob.oracle_cut_facets(m2, disk(0.3, x0=(0.5, 0.5)))
    assert fc.Nf == 4 * 4 + 5 * 3 and fc.Nboundary == 2 * (4 + 3)
    # first cell: its 4 edges in the local order [1,2],[3,4],[1,3],[2,4]
    fn = fc.array("facet_nodes").reshape(-1, 2)
    assert fn[:4].tolist() == [[0, 1], [5, 6], [0, 5], [1, 6]]
    # second cell: bottom, top, right (left is shared with the first cell)
    assert fn[4:7].tolist() == [[1, 2], [6, 7], [2, 7]]
    m3 = CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (3, 2, 2))
    fc = ob.oracle_cut_facets(m3, sphere(0.3, x0=(0.5, 0.5, 0.5)))
    assert fc.Nf == 3 * 12 + 3 * 2 + 3 * 2 + 2 * 2 and fc.Nboundary == 2 * (6 + 6 + 4)
    fn = fc.array("facet_nodes").reshape(-1, 4)
    assert fn[0].tolist() == [0, 1, 4, 5] and fn[1].tolist() == [12, 13, 16, 17]
    assert fn[2].tolist() == [0, 1, 12, 13] and fn[4].tolist() == [0, 4, 12, 16]


def test_divergence_theorem_with_boundary_facets_2d():
    # EmbeddedFacetDiscretizationsTests.jl:18-68: disk centred on a corner of the unit square
    n = 10
    model = CartesianDiscreteModel((0, 1, 0, 1), (n, n))
    geo = disk(0.72, x0=(1, 1))
    oc, fc = both(model, geo)
    p = prog(oc)
    a, b_u = divergence_theorem(oc, model, p, p)
    rng = np.random.default_rng(1234)
    vn = rng.random(model.num_nodes)  # the same field divergence_theorem draws
    b_f = FH.boundary_flux(model, fc, p, vn, np.ones(2))
    assert abs(b_f) > 1e-3
    assert abs(a - (b_u + b_f)) < 1.0e-9
    # the physical skeleton and boundary have IN and OUT parts that add up to the whole facets
    fn = fc.array("facet_nodes").reshape(-1, 2)
    for which in (ob.BOUNDARY_FACETS, ob.SKELETON_FACETS):
        tot = 0.0
        for side in (IN, OUT):
            sub = fc.select(p, which, ob.SEL_CUT, side)
            whole = fc.select(p, which, ob.SEL_BG, side)
            tot += fc.subfacet_measures(sub).sum() + sum(fc.bgfacet_measure(int(f)) for f in whole)
        nfac = fc.Nboundary if which == ob.BOUNDARY_FACETS else fc.Nf - fc.Nboundary
        assert abs(tot - nfac * 0.1) < 1e-12


def test_divergence_theorem_with_boundary_facets_3d():
    # sphere centred on the middle of the z-max face (x facets are simplexified along the other diagonal than the
    # traces of simplexify(HEX), so only models whose cut boundary facets are y- or z-normal are exact)
    n = 8
    model = CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (n, n, n))
    geo = sphere(0.37, x0=(0.5, 0.5, 1.0))
    oc, fc = both(model, geo)
    p = prog(oc)
    a, b_u = divergence_theorem(oc, model, p, p, gradu=[1, 1, -1], facet_degree=4)
    vn = np.random.default_rng(1234).random(model.num_nodes)
    b_f = FH.boundary_flux(model, fc, p, vn, np.array([1.0, 1.0, -1.0]))
    assert abs(b_f) > 1e-3
    # v n.grad(u) on a cut QUAD facet is Q1 (degree 2) -> the 3-point rule is exact
    assert abs(a - (b_u + b_f)) < 1.0e-9


@pytest.mark.parametrize("case", ["disk", "two_disks", "csg2d"])
def test_cellwise_divergence_theorem_2d(case):
    model = CartesianDiscreteModel((0, 1, 0, 1), (12, 9))
    if case == "disk":
        geo = disk(0.72, x0=(1, 1))
    elif case == "two_disks":
        geo = union(disk(0.3, x0=(0.3, 0.4)), disk(0.33, x0=(0.62, 0.55)))
    else:
        geo = setdiff(intersect(disk(0.45, x0=(0.5, 0.5)), plane(x0=(0.5, 0.7), v=(0.2, 1.0))), disk(0.2, x0=(0.4, 0.4)))
    oc, fc = both(model, geo)
    p = prog(oc)
    defect = FH.cellwise_divergence_defect(model, oc, fc, p, p)
    assert np.abs(defect).max() < 1e-12
    assert fc.Ncutf > 0 and fc.Nsf > fc.Ncutf


@pytest.mark.parametrize("case", ["sphere", "csg"])
def test_cellwise_divergence_theorem_3d(case):
    if case == "sphere":
        model = CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (7, 6, 5))
        geo = sphere(0.45, x0=(0.5, 0.55, 0.8))
    else:
        model = CartesianDiscreteModel((-0.8, -0.8, -0.8), (0.8, 0.8, 0.8), (8, 8, 8))
        geo = csg_geometry()
    oc, fc = both(model, geo)
    p = prog(oc, "csg" if case == "csg" else None)
    defect = FH.cellwise_divergence_defect(model, oc, fc, p, p)
    # cells whose two x-normal facets are not cut (see the 3-D test above)
    st = fc.bgfacet_to_inoutcut(p)
    fn = fc.array("facet_nodes").reshape(-1, 4)
    d, pos, minus, plus = FH.facet_cells(model, fn)
    bad = np.zeros(model.num_cells, bool)
    xcut = (d == 0) & (st == CUT)
    bad[minus[xcut & (minus >= 0)]] = True
    bad[plus[xcut & (plus >= 0)]] = True
    # multi-level-set x facets can also be cut by a single level set each while the CSG state is not CUT
    if fc.L > 1:
        rows = fc.rows("ls_to_facet_to_inoutcut")
        anycut = (d == 0) & (rows == CUT).any(0)
        bad[minus[anycut & (minus >= 0)]] = True
        bad[plus[anycut & (plus >= 0)]] = True
    assert (~bad).sum() > 0.1 * model.num_cells
    assert np.abs(defect[~bad]).max() < 1e-12
    # and the mismatch on the others is a discretisation effect, not a bookkeeping error
    assert np.abs(defect[bad]).max() < 0.05 * np.prod(model.sizes)


def test_facet_states_match_cell_states():
    # a facet whose cells are both IN (OUT) is IN (OUT); a CUT facet lies in CUT cells only
    model = CartesianDiscreteModel((-0.8, -0.8, -0.8), (0.8, 0.8, 0.8), (8, 8, 8))
    geo = csg_geometry()
    oc, fc = both(model, geo)
    fn = fc.array("facet_nodes").reshape(-1, 4)
    d, pos, minus, plus = FH.facet_cells(model, fn)
    for ls in range(fc.L):
        cst = oc.array("ls_to_bgcell_to_inoutcut", ls)
        fst = fc.array("ls_to_facet_to_inoutcut", ls)
        for nb in (minus, plus):
            m = nb >= 0
            assert np.all((cst[nb[m]] == CUT) | (cst[nb[m]] == fst[m]))
    # sub-facets keep their bg facet, states are +-1, points per facet id are within the facet's plane
    c2bg = fc.array("subfacets.cell_to_bgcell")
    assert np.all(np.diff(c2bg) >= 0)
    assert set(np.unique(c2bg).tolist()) == set(fc.array("cutfacet_to_facet").tolist())
    st = fc.rows("ls_to_subfacet_to_inout")
    assert set(np.unique(st).tolist()) <= {-1, 1}
    X = fc.array("subfacets.point_to_coords").reshape(-1, 3)
    R = fc.array("subfacets.point_to_rcoords").reshape(-1, 2)
    c2p = fc.array("subfacets.cell_to_points").reshape(-1, 3)
    assert R.min() >= 0 and R.max() <= 1
    # physical coordinates = facet map of the reference coordinates
    Xn = model.node_coordinates()[fn[c2bg - 1]]  # (Nsf, 4, 3)
    for v in range(3):
        r = R[c2p[:, v] - 1]
        x = Xn[:, 0] + r[:, :1] * (Xn[:, 1] - Xn[:, 0]) + r[:, 1:] * (Xn[:, 2] - Xn[:, 0])
        assert np.abs(x - X[c2p[:, v] - 1]).max() < 1e-13


# ---- simplexified models: the facet grid of the TRI / TET mesh ---------------------------------------------------------
def _simplex_cellwise_defect(model, oc, fc, p):
    """Per bg SIMPLEX K: D |K ∩ Ω| - [ Σ_{facets F of K} (x_F · n_F) |F ∩ Ω| + ∫_{Γ ∩ K} x · n ] (divergence theorem for
    F(x) = x on K ∩ Ω): ties the IN part of every bg facet of cut_facets -- boundary and interior, axis-aligned and
    diagonal -- to the subcells and the embedded boundary of cut().  A level set is linear on a simplex, so the trace of
    the cell's cut on a facet IS the facet's own cut: the defect is round-off for every cell."""
    D = model.D
    X = model.node_coordinates()
    cells = model.cell_node_ids() - 1
    Nc = cells.shape[0]
    vol = np.zeros(Nc)
    ids = oc.select_subcells(p, IN)
    c2bg = oc.array("subcells.cell_to_bgcell")
    np.add.at(vol, c2bg[ids - 1] - 1, oc.subcell_measures(ids))
    for c in oc.select_bgcells(p, IN):
        vol[c - 1] += oc.bgcell_measure(int(c))
    flux = np.zeros(Nc)
    fids, ori = oc.select_boundary(p)
    fX = oc.array("subfacets.point_to_coords").reshape(-1, D)
    f2p = oc.array("subfacets.facet_to_points").reshape(-1, D)
    f2bg = oc.array("subfacets.facet_to_bgcell")
    f2n = oc.array("subfacets.facet_to_normal").reshape(-1, D)
    nrm = f2n[fids - 1] * ori[:, None]
    x0 = fX[f2p[fids - 1][:, 0] - 1]
    np.add.at(flux, f2bg[fids - 1] - 1, (x0 * nrm).sum(1) * oc.subfacet_measures(fids))
    # IN measure of every bg facet
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    area = np.zeros(fc.Nf)
    for which in (ob.BOUNDARY_FACETS, ob.SKELETON_FACETS):
        sub = fc.select(p, which, ob.SEL_CUT, IN)
        np.add.at(area, fc.array("subfacets.cell_to_bgcell")[sub - 1] - 1, fc.subfacet_measures(sub))
        for f in fc.select(p, which, ob.SEL_BG, IN):
            area[f - 1] += fc.bgfacet_measure(int(f))
    facet_of = {tuple(sorted(r)): f for f, r in enumerate(fn.tolist())}
    lfacets = [(0, 1), (0, 2), (1, 2)] if D == 2 else [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
    for c in range(Nc):
        nodes = cells[c]
        for lf in lfacets:
            f = facet_of[tuple(sorted(int(nodes[i]) for i in lf))]
            if area[f] == 0.0:
                continue
            P = X[nodes[list(lf)]]
            opp = X[nodes[[i for i in range(D + 1) if i not in lf][0]]]
            if D == 2:
                e = P[1] - P[0]
                n = np.array([e[1], -e[0]])
            else:
                n = np.cross(P[1] - P[0], P[2] - P[0])
            n = n / np.linalg.norm(n)
            if np.dot(n, opp - P[0]) > 0:
                n = -n
            flux[c] += np.dot(P[0], n) * area[f]
    return D * vol - flux


@pytest.mark.parametrize("case", ["disk", "csg2d", "sphere", "csg3d"])
def test_cellwise_divergence_theorem_simplexified(case):
    from gecut import simplexify
    if case == "disk":
        model, geo, nm = simplexify(CartesianDiscreteModel((0, 1, 0, 1), (12, 9))), disk(0.72, x0=(1, 1)), None
    elif case == "csg2d":
        geo = setdiff(intersect(disk(0.45, x0=(0.5, 0.5)), plane(x0=(0.5, 0.7), v=(0.2, 1.0))), disk(0.2, x0=(0.4, 0.4)))
        model, nm = simplexify(CartesianDiscreteModel((0, 1, 0, 1), (12, 9))), None
    elif case == "sphere":
        model, geo, nm = simplexify(CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (6, 5, 4))), sphere(0.45, x0=(0.5, 0.55, 0.8)), None
    else:
        model, geo, nm = simplexify(CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (6, 6, 6))), csg_geometry(), "csg"
    oc, fc = both(model, geo)
    p = prog(oc, nm)
    assert fc.nvf == model.D and fc.Ncutf > 0 and fc.Nsf >= fc.Ncutf
    defect = _simplex_cellwise_defect(model, oc, fc, p)
    assert np.abs(defect).max() < 1e-12
    # boundary / skeleton split of the facet grid: IN + OUT parts add up to every facet
    for which in (ob.BOUNDARY_FACETS, ob.SKELETON_FACETS):
        tot = 0.0
        for side in (IN, OUT):
            sub = fc.select(p, which, ob.SEL_CUT, side)
            whole = fc.select(p, which, ob.SEL_BG, side)
            tot += fc.subfacet_measures(sub).sum() + sum(fc.bgfacet_measure(int(f)) for f in whole)
        isb = fc.array("facet_is_boundary").astype(bool)
        sel = np.nonzero(isb if which == ob.BOUNDARY_FACETS else ~isb)[0]
        ref = sum(fc.bgfacet_measure(int(f) + 1) for f in sel)
        assert abs(tot - ref) < 1e-11 * ref
