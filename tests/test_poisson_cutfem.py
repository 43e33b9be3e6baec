"""The reference's PDE known answer for this path (test/GridapEmbeddedTests/PoissonCutFEMTests.jl:116-117): Nitsche +
ghost-penalty CutFEM Poisson on the sphere with a manufactured linear solution, el2/ul2 < 1e-8 and eh1/uh1 < 1e-7,
assembled in numpy (tests/poisson_helpers.py) from what the path hands downstream -- the element matrices, the
SubFacetData layout with its glue and normals, the ghost skeleton.  CPU: the oracle.  GPU: the device results."""
import numpy as np
import pytest

import gecut
from gecut import CartesianDiscreteModel, sphere
from gecut.csg import compile_postfix
import oracle_binding as ob
import poisson_helpers as PH

IN, OUT, CUT = -1, 1, 0


def _model(n=10):
    geom = sphere(0.7)
    box = gecut.get_metadata(geom)
    return CartesianDiscreteModel(box.pmin, box.pmax, (n, n, n)), geom


def test_poisson_cutfem_manufactured_solution_oracle():
    model, geom = _model()
    oc = ob.oracle_cut(model, geom)
    p = compile_postfix(geom.get_tree(), oc.oid_to_ls)
    D = 3
    ids = oc.select_subcells(p, IN)
    fids, ori = oc.select_boundary(p)
    sub = {"c2p": oc.array("subcells.cell_to_points").reshape(-1, D + 1)[ids - 1],
           "c2bg": oc.array("subcells.cell_to_bgcell")[ids - 1],
           "X": oc.array("subcells.point_to_coords").reshape(-1, D),
           "R": oc.array("subcells.point_to_rcoords").reshape(-1, D)}
    fac = {"c2p": oc.array("subfacets.facet_to_points").reshape(-1, D)[fids - 1],
           "c2bg": oc.array("subfacets.facet_to_bgcell")[fids - 1],
           "X": oc.array("subfacets.point_to_coords").reshape(-1, D),
           "R": oc.array("subfacets.point_to_rcoords").reshape(-1, D),
           "N": oc.array("subfacets.facet_to_normal").reshape(-1, D)[fids - 1] * ori[:, None]}
    bg_state = oc.bgcell_to_inoutcut(p)
    el2, eh1, nghost = PH.solve_poisson_cutfem(model, bg_state, oc.array("cutcell_to_cell"), oc.cutcell_laplacian(ids),
                                               oc.bgcell_laplacian(1), sub, fac)
    assert el2 < 1.0e-8 and eh1 < 1.0e-7, (el2, eh1)
    # the ghost facets of the assembly are the ones GhostSkeleton(cutgeo) marks
    ofc = ob.oracle_cut_facets(model, geom)
    assert nghost == int(np.count_nonzero(ob.oracle_ghost_skeleton(oc, ofc, p, IN)))


@pytest.mark.gpu
def test_poisson_cutfem_manufactured_solution_gpu():
    from gecut import cut, Triangulation, EmbeddedBoundary, Measure, PHYSICAL, GhostSkeleton
    from gecut.measures import integrate_laplacian
    from gecut.discretizations import compute_bgcell_to_inoutcut
    model, geom = _model()
    cg = cut(model, geom)
    h = cg.to_host()
    om = Triangulation(cg, PHYSICAL)
    ga = EmbeddedBoundary(cg)
    K_cut, K_bg = integrate_laplacian(Measure(om, 2))
    ids, fids, ori = om.sub_ids, ga.sub_ids, ga.orientation
    sub = {"c2p": h["sc_cell_to_points"][ids - 1], "c2bg": h["sc_cell_to_bgcell"][ids - 1],
           "X": h["sc_point_to_coords"], "R": h["sc_point_to_rcoords"]}
    fac = {"c2p": h["sf_facet_to_points"][fids - 1], "c2bg": h["sf_facet_to_bgcell"][fids - 1],
           "X": h["sf_point_to_coords"], "R": h["sf_point_to_rcoords"],
           "N": h["sf_facet_to_normal"][fids - 1] * ori[:, None].astype(np.float64)}
    el2, eh1, nghost = PH.solve_poisson_cutfem(model, compute_bgcell_to_inoutcut(cg), h["cutcell_to_cell"], K_cut, K_bg[0],
                                               sub, fac)
    assert el2 < 1.0e-8 and eh1 < 1.0e-7, (el2, eh1)
    assert nghost == int(np.count_nonzero(GhostSkeleton(cg)))


def _bimaterial_model(n=30):
    from gecut import disk, complement, union, simplexify
    geo1 = disk(0.7, name="geo1")
    geo2 = complement(geo1, name="geo2")
    model = simplexify(CartesianDiscreteModel((-1.0, 1.0, -1.0, 1.0), (n, n)))
    return model, union(geo1, geo2), geo1, geo2


def test_bimaterial_poisson_manufactured_solution_oracle():
    """BimaterialPoissonCutFEMTests.jl:137-138 on the oracle: P1 matrices of both sides, get_cell_measure of both sides and
    of the interface (the stabilisation weights), EmbeddedBoundary(cutgeo, geo1, geo2) with its normals."""
    model, geo, geo1, geo2 = _bimaterial_model()
    oc = ob.oracle_cut(model, geo)
    assert oc.L == 1
    p1 = compile_postfix(geo1.get_tree(), oc.oid_to_ls)
    p2 = compile_postfix(geo2.get_tree(), oc.oid_to_ls)
    D = 2
    c2p = oc.array("subcells.cell_to_points").reshape(-1, D + 1)
    c2bg = oc.array("subcells.cell_to_bgcell")
    X, R = oc.array("subcells.point_to_coords").reshape(-1, D), oc.array("subcells.point_to_rcoords").reshape(-1, D)
    ids1, ids2 = oc.select_subcells(p1, IN), oc.select_subcells(p2, IN)
    assert np.array_equal(ids2, oc.select_subcells(p1, OUT))
    sub1 = {"c2p": c2p[ids1 - 1], "c2bg": c2bg[ids1 - 1], "X": X, "R": R}
    sub2 = {"c2p": c2p[ids2 - 1], "c2bg": c2bg[ids2 - 1], "X": X, "R": R}
    fids, ori = oc.select_boundary(p1, p2)
    fac = {"c2p": oc.array("subfacets.facet_to_points").reshape(-1, D)[fids - 1],
           "c2bg": oc.array("subfacets.facet_to_bgcell")[fids - 1],
           "X": oc.array("subfacets.point_to_coords").reshape(-1, D),
           "R": oc.array("subfacets.point_to_rcoords").reshape(-1, D),
           "N": oc.array("subfacets.facet_to_normal").reshape(-1, D)[fids - 1] * ori[:, None]}
    state1 = oc.bgcell_to_inoutcut(p1)
    nc = len(state1)

    def cell_measure(ids, st):
        m = np.bincount(c2bg[ids - 1] - 1, weights=oc.subcell_measures(ids), minlength=nc)
        for c in np.flatnonzero(state1 == st):
            m[c] = oc.bgcell_measure(int(c) + 1)
        return m

    K_bg = np.stack([oc.bgcell_laplacian(t + 1) for t in range(model.simplices_per_cube)])
    el2, eh1 = PH.solve_bimaterial_poisson(model, state1, oc.array("cutcell_to_cell"), oc.cutcell_laplacian(ids1),
                                           oc.cutcell_laplacian(ids2), K_bg, cell_measure(ids1, IN), cell_measure(ids2, OUT),
                                           sub1, sub2, fac)
    assert el2 < 1.0e-8 and eh1 < 1.0e-7, (el2, eh1)


@pytest.mark.gpu
def test_bimaterial_poisson_manufactured_solution_gpu():
    """The same known answer from the device results: Triangulation(cutgeo, PHYSICAL, geo1 / geo2), the two-geometry
    EmbeddedBoundary, get_cell_measure of both sides and the P1 element matrices of both sides."""
    from gecut import cut, Triangulation, EmbeddedBoundary, Measure, PHYSICAL_IN
    from gecut.measures import integrate_laplacian, get_cell_measure
    from gecut.discretizations import compute_bgcell_to_inoutcut
    model, geo, geo1, geo2 = _bimaterial_model()
    cg = cut(model, geo)
    h = cg.to_host()
    t1 = Triangulation(cg, PHYSICAL_IN, "geo1")
    t2 = Triangulation(cg, PHYSICAL_IN, "geo2")
    ga = EmbeddedBoundary(cg, "geo1", "geo2")
    K1_cut, K_bg = integrate_laplacian(Measure(t1, 2))
    K2_cut, _ = integrate_laplacian(Measure(t2, 2))
    ids1, ids2, fids, ori = t1.sub_ids, t2.sub_ids, ga.sub_ids, ga.orientation
    sub1 = {"c2p": h["sc_cell_to_points"][ids1 - 1], "c2bg": h["sc_cell_to_bgcell"][ids1 - 1],
            "X": h["sc_point_to_coords"], "R": h["sc_point_to_rcoords"]}
    sub2 = {"c2p": h["sc_cell_to_points"][ids2 - 1], "c2bg": h["sc_cell_to_bgcell"][ids2 - 1],
            "X": h["sc_point_to_coords"], "R": h["sc_point_to_rcoords"]}
    fac = {"c2p": h["sf_facet_to_points"][fids - 1], "c2bg": h["sf_facet_to_bgcell"][fids - 1],
           "X": h["sf_point_to_coords"], "R": h["sf_point_to_rcoords"],
           "N": h["sf_facet_to_normal"][fids - 1] * ori[:, None].astype(np.float64)}
    el2, eh1 = PH.solve_bimaterial_poisson(model, compute_bgcell_to_inoutcut(cg, "geo1"), h["cutcell_to_cell"], K1_cut, K2_cut,
                                           K_bg, get_cell_measure(t1), get_cell_measure(t2), sub1, sub2, fac)
    assert el2 < 1.0e-8 and eh1 < 1.0e-7, (el2, eh1)


def test_poisson_agfem_manufactured_solution_oracle():
    """PoissonAgFEMTests.jl:80-81 on the oracle: setdiff of two disks (two level sets), AggregateAllCutCells, Nitsche
    Poisson on the aggregated Q1 space without ghost penalty -- the consumer of `aggregate`."""
    from gecut import disk, setdiff
    R = 0.5
    geo3 = setdiff(disk(R, x0=(0.0, 0.0)), disk(R, x0=(0.8 * 2 * R, 0.0)))
    t = 1.01
    model = CartesianDiscreteModel((-t * R, -t * R), (t * R, t * R), (30, 30))
    oc = ob.oracle_cut(model, geo3)
    assert oc.L == 2
    fcut = ob.oracle_cut_facets(model, geo3)
    p = compile_postfix(geo3.get_tree(), oc.oid_to_ls)
    D = 2
    ids = oc.select_subcells(p, IN)
    fids, ori = oc.select_boundary(p)
    sub = {"c2p": oc.array("subcells.cell_to_points").reshape(-1, D + 1)[ids - 1],
           "c2bg": oc.array("subcells.cell_to_bgcell")[ids - 1],
           "X": oc.array("subcells.point_to_coords").reshape(-1, D),
           "R": oc.array("subcells.point_to_rcoords").reshape(-1, D)}
    fac = {"c2p": oc.array("subfacets.facet_to_points").reshape(-1, D)[fids - 1],
           "c2bg": oc.array("subfacets.facet_to_bgcell")[fids - 1],
           "X": oc.array("subfacets.point_to_coords").reshape(-1, D),
           "R": oc.array("subfacets.point_to_rcoords").reshape(-1, D),
           "N": oc.array("subfacets.facet_to_normal").reshape(-1, D)[fids - 1] * ori[:, None]}
    agg, ok = ob.oracle_aggregate(oc, fcut, p, IN, 1.0)
    assert ok
    el2, eh1, nslaves = PH.solve_poisson_agfem_2d(model, oc.bgcell_to_inoutcut(p), oc.array("cutcell_to_cell"),
                                                  oc.cutcell_laplacian(ids), oc.bgcell_laplacian(1), agg, sub, fac)
    assert nslaves > 0
    assert el2 < 1.0e-8 and eh1 < 1.0e-7, (el2, eh1)


def test_tracefem_projection_oracle():
    """TraceFEMTests.jl:53-54 on the oracle: a problem posed on the embedded boundary alone (P1 background functions of the
    CUT cells traced on Γ through the glue of the SubFacetData) with the CUT-cell ghost skeleton; l2 error on Γ < 1e-10."""
    from gecut import disk, simplexify
    geom = disk(0.7)
    box = gecut.get_metadata(geom)
    model = simplexify(CartesianDiscreteModel(box.pmin, box.pmax, (10, 10)))
    oc = ob.oracle_cut(model, geom)
    p = compile_postfix(geom.get_tree(), oc.oid_to_ls)
    D = 2
    fids, ori = oc.select_boundary(p)
    fac = {"c2p": oc.array("subfacets.facet_to_points").reshape(-1, D)[fids - 1],
           "c2bg": oc.array("subfacets.facet_to_bgcell")[fids - 1],
           "X": oc.array("subfacets.point_to_coords").reshape(-1, D),
           "R": oc.array("subfacets.point_to_rcoords").reshape(-1, D)}
    state = oc.bgcell_to_inoutcut(p)
    # every bg cell that carries a piece of Γ is a CUT cell (Triangulation(cutgeom, CUT, geom) is the support of the space)
    assert np.all(state[fac["c2bg"] - 1] == CUT)
    err, nghost = PH.solve_tracefem_projection_2d(model, state, fac)
    assert nghost > 0
    assert err < 1.0e-10, err
