"""Pins the CPU oracle against every known answer the reference's own tests hold for the path
(SURVEY.md section 8c): compute_case literals, disk area/perimeter, quadrilateral areas, the divergence theorem
(which jointly pins normal orientation, sub-Jacobian signs, glue ref-coords and quadrature)."""
import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, simplexify, disk, sphere, cube, cylinder, quadrilateral, union, intersect,
                   setdiff, complement, get_geometry, tube, plane)
from gecut.csg import compile_postfix
import oracle_binding as ob
import helpers as H

IN, OUT, CUT = -1, 1, 0


def ocut(model, geo):
    oc = ob.oracle_cut(model, geo)
    return oc


def prog(oc, geo_or_name=None):
    geo = oc.geo if geo_or_name is None else (get_geometry(oc.geo, geo_or_name) if isinstance(geo_or_name, str) else geo_or_name)
    return compile_postfix(geo.get_tree(), oc.oid_to_ls)


def test_compute_case_literals():
    # test/LevelSetCuttersTests/LookupTablesTests.jl:8-11
    import gen_lookup_tables as g
    assert g.compute_case([0, 0, 0]) == 1
    assert g.compute_case([1, 0, 0]) == 2
    assert g.compute_case([1, 0, 1]) == 6
    assert g.compute_case([1, 1, 1]) == 8


def physical_measures(oc, p, p_boundary=None, side=IN):
    ids = oc.select_subcells(p, side)
    bg = oc.select_bgcells(p, side)
    vol = oc.subcell_measures(ids).sum() + sum(oc.bgcell_measure(c) for c in bg[:1]) * len(bg) if not oc.model.simplexified \
        else oc.subcell_measures(ids).sum() + sum(oc.bgcell_measure(int(c)) for c in bg)
    fids, ori = oc.select_boundary(p if p_boundary is None else p_boundary)
    surf = oc.subfacet_measures(fids).sum()
    return vol, surf


@pytest.mark.parametrize("simplex", [False, True])
def test_disk_area_perimeter(simplex):
    # test/InterfacesTests/EmbeddedDiscretizationsTests.jl:11-56
    R = 0.7
    geom = disk(R)
    box = gecut.get_metadata(geom)
    model = CartesianDiscreteModel(box.pmin, box.pmax, (50, 50))
    if simplex:
        model = simplexify(model)
    oc = ocut(model, geom)
    vol, surf = physical_measures(oc, prog(oc))
    assert abs(np.pi * R ** 2 - vol) < 1.0e-3
    assert abs(surf - 2 * np.pi * R) < 1.0e-3
    # IN + OUT partition the box
    ids_out = oc.select_subcells(prog(oc), OUT)
    bg_out = oc.select_bgcells(prog(oc), OUT)
    vol_out = oc.subcell_measures(ids_out).sum() + sum(oc.bgcell_measure(int(c)) for c in bg_out)
    box_vol = np.prod(np.asarray(box.pmax) - np.asarray(box.pmin))
    assert abs(vol + vol_out - box_vol) < 1e-12


def divergence_theorem(oc, model, p_omega, p_gamma1, p_gamma2=None, seed=1234, gradu=None, facet_degree=2):
    """a = ∫_Ω ∇v·∇u, b = ∫_Γ v n·∇u for u linear and v a random bg FE function
    (test/InterfacesTests/EmbeddedDiscretizationsTests.jl:58-107)."""
    D = model.D
    rng = np.random.default_rng(seed)
    vn = rng.random(model.num_nodes)
    gradu = np.ones(D) if gradu is None else np.asarray(gradu, float)
    cn = model.cell_node_ids()
    X = oc.array("subcells.point_to_coords").reshape(-1, D)
    Rr = oc.array("subcells.point_to_rcoords").reshape(-1, D)
    c2p = oc.array("subcells.cell_to_points").reshape(-1, D + 1)
    c2bg = oc.array("subcells.cell_to_bgcell")
    ids = oc.select_subcells(p_omega, IN)
    bg = oc.select_bgcells(p_omega, IN)

    def f_omega(bgcell, eta, x):
        _, g = H.bg_field_at(model, cn, bgcell, eta, vn)
        return g @ gradu

    a = H.integrate_on_simplices(model, cn, c2p[ids - 1], c2bg[ids - 1], X, Rr, f_omega)
    a += H.integrate_on_bgcells(model, cn, bg, f_omega)
    fX = oc.array("subfacets.point_to_coords").reshape(-1, D)
    fR = oc.array("subfacets.point_to_rcoords").reshape(-1, D)
    f2p = oc.array("subfacets.facet_to_points").reshape(-1, D)
    f2bg = oc.array("subfacets.facet_to_bgcell")
    f2n = oc.array("subfacets.facet_to_normal").reshape(-1, D)
    fids, ori = oc.select_boundary(p_gamma1, p_gamma2)
    normals = f2n[fids - 1] * ori[:, None]
    ndu = normals @ gradu

    def f_gamma(bgcell, eta, x):
        v, _ = H.bg_field_at(model, cn, bgcell, eta, vn)
        return v * ndu

    b = H.integrate_on_simplices(model, cn, f2p[fids - 1], f2bg[fids - 1], fX, fR, f_gamma, degree=facet_degree)
    return a, b


@pytest.mark.parametrize("simplex", [False, True])
def test_divergence_theorem_disk(simplex):
    R = 0.7
    geom = disk(R)
    box = gecut.get_metadata(geom)
    model = CartesianDiscreteModel(box.pmin, box.pmax, (50, 50))
    if simplex:
        model = simplexify(model)
    oc = ocut(model, geom)
    a, b = divergence_theorem(oc, model, prog(oc), prog(oc))
    assert abs(a - b) < 1.0e-9
    assert abs(a) > 1e-3


@pytest.mark.parametrize("simplex", [False, True])
def test_divergence_theorem_sphere_3d(simplex):
    geom = sphere(0.7)
    box = gecut.get_metadata(geom)
    model = CartesianDiscreteModel(box.pmin, box.pmax, (12, 12, 12))
    if simplex:
        model = simplexify(model)
    oc = ocut(model, geom)
    a, b = divergence_theorem(oc, model, prog(oc), prog(oc), gradu=[1, 1, -1], facet_degree=4)
    assert abs(a - b) < 1.0e-9
    vol, surf = physical_measures(oc, prog(oc))
    assert abs(vol - 4 / 3 * np.pi * 0.7 ** 3) / vol < 0.05


def test_quadrilateral_areas():
    # test/LevelSetCuttersTests/AnalyticalGeometriesTests.jl:190-242
    model = CartesianDiscreteModel((-0.01, 2.01, -0.01, 1.01), (5, 5))
    for d1, d2 in (((1, 0), (0, 1)), ((1, 0), (1, 1))):
        geo = quadrilateral(x0=(0, 0), d1=d1, d2=d2)
        oc = ocut(model, geo)
        vol, _ = physical_measures(oc, prog(oc))
        assert abs(1.0 - vol) < 1e-9
    model = CartesianDiscreteModel((-0.01, 2.01, -0.01, 3.01), (5, 5))
    geo = quadrilateral(x0=(0, 0), d1=(1, 1), d2=(1, 2))
    oc = ocut(model, geo)
    vol, _ = physical_measures(oc, prog(oc))
    assert abs(1.0 - vol) < 1e-9


def csg_geometry():
    # examples/PoissonCSGCutFEM/PoissonCSGCutFEM.jl:14-23
    R = 0.5
    geo1 = cylinder(R, v=(1, 0, 0))
    geo2 = cylinder(R, v=(0, 1, 0))
    geo3 = cylinder(R, v=(0, 0, 1))
    geo4 = union(union(geo1, geo2), geo3, name="source")
    geo5 = sphere(1)
    geo6 = cube(L=1.5)
    geo7 = intersect(geo6, geo5)
    geo8 = setdiff(geo7, geo4, name="csg")
    return geo8


def test_csg_multi_levelset_divergence_theorem():
    n = 10
    model = CartesianDiscreteModel((-0.8, -0.8, -0.8), (0.8, 0.8, 0.8), (n, n, n))
    geo = csg_geometry()
    oc = ocut(model, geo)
    assert oc.L == 10
    p = prog(oc, "csg")
    a, b = divergence_theorem(oc, model, p, p, gradu=[1, 1, -1], facet_degree=4)
    assert abs(a - b) < 1.0e-9
    # ... and through the Q1 element matrices of the ten-level-set cut (AffineFEOperator's input in the example)
    a_mat = matrix_form(oc, model, p, [1, 1, -1])
    assert abs(a_mat - a) < 1e-12 * max(1.0, abs(a)) and abs(a_mat - b) < 1.0e-9
    # the Dirichlet boundary of the example: EmbeddedBoundary(cutgeo,"csg","source") is part of the whole boundary
    f_all, _ = oc.select_boundary(p)
    f_src, _ = oc.select_boundary(p, prog(oc, "source"))
    assert 0 < len(f_src) < len(f_all)
    assert set(f_src.tolist()) <= set(f_all.tolist())
    # volume against a Monte-Carlo estimate of the CSG body
    vol, surf = physical_measures(oc, p)

    def inside(x):
        c = np.all(np.abs(x) <= 0.75, axis=1) & ((x ** 2).sum(1) <= 1.0)
        for ax in range(3):
            o = [d for d in range(3) if d != ax]
            c &= ~((x[:, o] ** 2).sum(1) <= 0.25)
        return c

    mc = H.monte_carlo_volume(model, inside, n=2_000_000)
    assert abs(vol - mc) / mc < 0.08


def test_two_phase_union_with_complement():
    # examples/BimaterialLinElastCutFEM/BimaterialLinElastCutFEM.jl:29-56 at small size
    model = CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (9, 6, 9))
    R = 0.4
    geo1 = cylinder(R, x0=(0.0, 1.0, 0.7))
    geo2 = cylinder(R, x0=(0.0, 1.0, 2.3))
    geo3 = union(geo1, geo2, name="steel")
    geo4 = complement(geo3, name="concrete")
    geo = union(geo3, geo4)
    oc = ocut(model, geo)
    assert oc.L == 2
    ps, pc = prog(oc, "steel"), prog(oc, "concrete")
    v1, _ = physical_measures(oc, ps)
    v2, _ = physical_measures(oc, pc)
    assert abs(v1 + v2 - 18.0) < 1e-10
    f1, o1 = oc.select_boundary(ps, pc)
    f2, o2 = oc.select_boundary(pc, ps)
    assert len(f1) > 0 and np.array_equal(f1, f2) and np.array_equal(o1, -o2)


def matrix_form(oc, model, p, gradu, seed=1234):
    """sum over the active cells of v_c^T K_c u_c with the element matrices of the path (cut cells: contributions of the
    selected subcells moved to their bg cell; uncut IN cells: the bg-cell matrix), u = gradu·x, v as in divergence_theorem()."""
    rng = np.random.default_rng(seed)
    vn = rng.random(model.num_nodes)
    un = model.node_coordinates() @ np.asarray(gradu, float)
    cn = model.cell_node_ids()
    K_cut = oc.cutcell_laplacian(oc.select_subcells(p, IN))
    nodes = cn[oc.array("cutcell_to_cell") - 1] - 1
    a_mat = np.einsum("ka,kab,kb->", vn[nodes], K_cut, un[nodes])
    for c in oc.select_bgcells(p, IN):
        nd = cn[int(c) - 1] - 1
        a_mat += vn[nd] @ oc.bgcell_laplacian(int(c)) @ un[nd]
    return float(a_mat)


@pytest.mark.parametrize("dim,simplex", [(2, False), (2, True), (3, False), (3, True)])
def test_element_matrices_reproduce_the_divergence_theorem(dim, simplex):
    """The reference checks the divergence theorem once more AFTER `move_contributions` has gathered the sub-cell integrals
    of ∫∇v·∇u per background cell (test/InterfacesTests/EmbeddedDiscretizationsTests.jl:96-110).  The per-cut-cell element
    matrices of this path ARE those moved contributions (in matrix form), so with u linear (exactly representable in the
    Q1 / P1 background space) and v a random background FE function
        sum_cells v_c^T K_c u_c  ==  ∫_Ω ∇v·∇u  ==  ∫_Γ v n·∇u          (1e-9, the reference's tolerance)
    which pins the matrices -- shape-function gradients at the glued reference points, quadrature, subcell -> bg cell
    accumulation -- to the reference's own known answer, not only to the oracle."""
    if dim == 2:
        geom, part, gradu, fdeg = disk(0.7), (50, 50), [1.0, 1.0], 2
    else:
        geom, part, gradu, fdeg = sphere(0.7), (12, 12, 12), [1.0, 1.0, -1.0], 4
    box = gecut.get_metadata(geom)
    model = CartesianDiscreteModel(box.pmin, box.pmax, part)
    if simplex:
        model = simplexify(model)
    oc = ocut(model, geom)
    p = prog(oc)
    a, b = divergence_theorem(oc, model, p, p, gradu=gradu, facet_degree=fdeg)
    a_mat = matrix_form(oc, model, p, gradu)
    assert abs(a_mat - a) < 1e-12 * max(1.0, abs(a))
    assert abs(a_mat - b) < 1.0e-9
    assert abs(a_mat) > 1e-3
