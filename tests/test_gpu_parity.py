"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): bit-exact IN/OUT/CUT classification, subcell/subfacet counts, connectivity and ordering;
<= 1e-12 relative on coordinates, normals, measures and element matrices.  Coordinates/normals are in fact compared for
exact equality first (both sides use the same IEEE operations without FMA contraction) and only then with the tolerance.
"""
import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, simplexify, disk, sphere, cube, cylinder, plane, doughnut, olympic_rings,
                   quadrilateral, square, tube, popcorn, union, intersect, setdiff, complement, get_geometry,
                   DiscreteGeometry, AnalyticalGeometry, LevelSetCutter, cut, Triangulation, EmbeddedBoundary, Measure,
                   PHYSICAL, PHYSICAL_IN, PHYSICAL_OUT, ACTIVE, ACTIVE_OUT, CUT_IN, IN, OUT)
from gecut.csg import compile_postfix
from gecut.measures import integrate_measure, integrate_laplacian
from gecut.discretizations import compute_bgcell_to_inoutcut, compute_subcell_to_inout
import oracle_binding as ob

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def csg_geometry():
    R = 0.5
    geo1 = cylinder(R, v=(1, 0, 0))
    geo2 = cylinder(R, v=(0, 1, 0))
    geo3 = cylinder(R, v=(0, 0, 1))
    geo4 = union(union(geo1, geo2), geo3, name="source")
    geo5 = sphere(1)
    geo6 = cube(L=1.5)
    geo7 = intersect(geo6, geo5)
    return setdiff(geo7, geo4, name="csg")


def stokes_geometry():
    # examples/StokesTubeWithObstacleCutFEM/StokesTubeWithObstacleCutFEM.jl:17-28,51
    R, L = 0.3, 2.0
    x0 = (0.0, 0.0, 0.0)
    geo2 = tube(R, L, x0=x0)
    geo3 = sphere(0.3 * R, x0=(L / 3, 0.0, 0.0))
    geo4 = complement(get_geometry(geo2, "walls"))
    geo5 = union(geo3, geo4, name="solid")
    geo1 = setdiff(geo2, geo3, name="fluid")
    m = 0.05
    pmin = (x0[0] - m, x0[1] - R, x0[2] - R)
    pmax = (x0[0] + m + L, x0[1] + R, x0[2] + R)
    return union(geo1, geo5), pmin, pmax


def three_disks():
    # test/LevelSetCuttersTests/CutTriangulationsTests.jl:15-30
    R = 0.85
    geo1 = disk(R, name="disk1")
    geo2 = disk(R, x0=(1, 1), name="disk2")
    geo3 = disk(R, x0=(1, 0), name="disk3")
    geo5 = union(geo1, geo2)
    return union(geo5, geo3)


def bimaterial():
    R = 0.4
    geo1 = cylinder(R, x0=(0.0, 1.0, 0.7))
    geo2 = cylinder(R, x0=(0.0, 1.0, 2.3))
    geo3 = union(geo1, geo2, name="steel")
    geo4 = complement(geo3, name="concrete")
    return union(geo3, geo4)


def box_model(geo, part, simplex=False):
    box = gecut.get_metadata(geo)
    m = CartesianDiscreteModel(box.pmin, box.pmax, part)
    return simplexify(m) if simplex else m


CASES = {
    "disk_quad_50": lambda: (box_model(disk(0.7), (50, 50)), disk(0.7)),
    "disk_tri_50": lambda: (box_model(disk(0.7), (50, 50), True), disk(0.7)),
    "disk_quad_odd": lambda: (box_model(disk(0.7), (37, 45)), disk(0.7)),
    "sphere_hex_12": lambda: (box_model(sphere(0.7), (12, 12, 12)), sphere(0.7)),
    "sphere_tet_12": lambda: (box_model(sphere(0.7), (12, 12, 12), True), sphere(0.7)),
    "sphere_hex_odd": lambda: (box_model(sphere(0.7), (37, 21, 13)), sphere(0.7)),
    "sphere_tet_odd": lambda: (box_model(sphere(0.7), (35, 9, 17), True), sphere(0.7)),
    "csg_hex_10": lambda: (CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (10, 10, 10)), csg_geometry()),
    "csg_hex_20": lambda: (CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (20, 20, 20)), csg_geometry()),
    "csg_tet_8": lambda: (simplexify(CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (8, 8, 8))), csg_geometry()),
    "three_disks_40": lambda: (CartesianDiscreteModel((-1, -1), (2, 2), (40, 40)), three_disks()),
    "three_disks_tri": lambda: (simplexify(CartesianDiscreteModel((-1, -1), (2, 2), (33, 20))), three_disks()),
    "square_2d": lambda: (CartesianDiscreteModel((-1, -1), (1, 1), (24, 24)), square(L=1.1)),
    "quadrilateral": lambda: (CartesianDiscreteModel((-0.01, 2.01, -0.01, 3.01), (5, 5)),
                              quadrilateral(x0=(0, 0), d1=(1, 1), d2=(1, 2))),
    "bimaterial": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (9, 12, 18)), bimaterial()),
    "doughnut": lambda: (box_model(doughnut(0.6, 0.25), (20, 20, 8)), doughnut(0.6, 0.25)),
    "olympic": lambda: (box_model(olympic_rings(0.6, 0.1), (40, 20, 4)), olympic_rings(0.6, 0.1)),
    "stokes_tet": lambda: (lambda g: (simplexify(CartesianDiscreteModel(g[1], g[2], (16, 4, 4))), g[0]))(stokes_geometry()),
    "stokes_hex": lambda: (lambda g: (CartesianDiscreteModel(g[1], g[2], (24, 6, 6)), g[0]))(stokes_geometry()),
    "no_cut": lambda: (CartesianDiscreteModel((2, 2, 2), (3, 3, 3), (6, 5, 4)), sphere(0.7)),
    "all_in": lambda: (CartesianDiscreteModel((-.1, -.1), (.1, .1), (7, 3)), disk(0.7)),
    # models several classification tiles wide (64 x 8 x 8 / 64 x 16 x 8 cubes) with surfaces ON tile boundaries and node planes:
    # the interval culling (k_classify_cull) must hand every tile a surface touches to the exact marching kernel
    "cull_plane_on_tile_boundary": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 1., 1.), (96, 32, 24)), cull_planes()),
    "cull_sphere_through_nodes_tet": lambda: (simplexify(CartesianDiscreteModel((0., 0., 0.), (4., 2., 2.), (128, 16, 16))),
                                              sphere(0.75, x0=(2.0, 1.0, 1.0))),
    "cull_oblique_cylinders": lambda: (CartesianDiscreteModel((0., 0., 0.), (4., 1., 1.), (128, 24, 20)),
                                       union(cylinder(0.3, x0=(2.0, 0.5, 0.5), v=(1, 1, 0.5)),
                                             cylinder(0.2, x0=(1.0, 0.5, 0.5), v=(0, 0, 1)))),
    "cull_disk_2d": lambda: (CartesianDiscreteModel((0., 0.), (8., 2.), (512, 64)), disk(0.75, x0=(4.0, 1.0))),
}


def cull_planes():
    # x = 2.0 is node column 64 (the boundary of the first tile), z = 1/3 is node plane 8 (a chunk boundary): values are exactly 0
    # there (IN); the third plane is oblique
    return intersect(intersect(plane(x0=(2.0, 0., 0.), v=(1, 0, 0)), plane(x0=(0., 0., 8 * (1.0 / 24)), v=(0, 0, -1))),
                     plane(x0=(1.0, 0.5, 0.5), v=(-1, 0.3, 0.2)))


def assert_float_equal(a, b, what):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if np.array_equal(a, b):
        return
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    err = np.abs(a - b).max() / scale
    assert err <= RTOL, f"{what}: max relative error {err:.3e} > {RTOL}"


def compare_layout(gd, oc):
    s = gd.sizes
    D, L = oc.D, oc.L
    assert (s.num_bgcells, s.num_cutcells, s.num_subcells, s.num_subcell_points, s.num_subfacets,
            s.num_subfacet_points, s.num_levelsets) == (oc.Nc, oc.Ncut, oc.Nsc, oc.Nscp, oc.Nsf, oc.Nsfp, oc.L)
    h = gd.to_host()
    assert np.array_equal(h["ls_to_bgcell_to_inoutcut"], oc.rows("ls_to_bgcell_to_inoutcut"))
    assert np.array_equal(h["cutcell_to_cell"], oc.array("cutcell_to_cell"))
    assert np.array_equal(h["sc_cell_to_points"].ravel(), oc.array("subcells.cell_to_points"))
    assert np.array_equal(h["sc_cell_to_bgcell"], oc.array("subcells.cell_to_bgcell"))
    assert np.array_equal(h["ls_to_subcell_to_inout"], oc.rows("ls_to_subcell_to_inout").reshape(L, -1))
    assert np.array_equal(h["sf_facet_to_points"].ravel(), oc.array("subfacets.facet_to_points"))
    assert np.array_equal(h["sf_facet_to_bgcell"], oc.array("subfacets.facet_to_bgcell"))
    assert np.array_equal(h["ls_to_subfacet_to_inout"], oc.rows("ls_to_subfacet_to_inout").reshape(L, -1))
    assert_float_equal(h["sc_point_to_coords"].ravel(), oc.array("subcells.point_to_coords"), "subcell coords")
    assert_float_equal(h["sc_point_to_rcoords"].ravel(), oc.array("subcells.point_to_rcoords"), "subcell rcoords")
    assert_float_equal(h["sf_point_to_coords"].ravel(), oc.array("subfacets.point_to_coords"), "subfacet coords")
    assert_float_equal(h["sf_point_to_rcoords"].ravel(), oc.array("subfacets.point_to_rcoords"), "subfacet rcoords")
    assert_float_equal(h["sf_facet_to_normal"].ravel(), oc.array("subfacets.facet_to_normal"), "normals")


def named_programs(gd):
    names = [None] + [n for n in gecut.get_geometry_names(gd.geo)]
    return names


@pytest.mark.parametrize("case", sorted(CASES))
def test_cut_layout_matches_oracle(case):
    model, geo = CASES[case]()
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    compare_layout(gd, oc)


@pytest.mark.parametrize("case", ["disk_quad_50", "disk_tri_50", "sphere_hex_12", "sphere_tet_12", "csg_hex_10",
                                  "three_disks_40", "bimaterial", "stokes_tet", "no_cut", "all_in"])
def test_selectors_and_integrals_match_oracle(case):
    model, geo = CASES[case]()
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    oid = {k: v for k, v in gd.oid_to_ls.items()}
    names = named_programs(gd)
    for name in names:
        prog = gd.program(name)
        assert np.array_equal(compute_bgcell_to_inoutcut(gd, name), oc.bgcell_to_inoutcut(prog))
        if oc.Nsc:
            assert np.array_equal(compute_subcell_to_inout(gd, name), oc.subcell_to_inout(prog))
        for side, sel in ((IN, PHYSICAL_IN), (OUT, PHYSICAL_OUT)):
            t = Triangulation(gd, sel, name)
            ids = oc.select_subcells(prog, side)
            bg = oc.select_bgcells(prog, side)
            assert np.array_equal(t.sub_ids, ids)
            assert np.array_equal(t.bg_ids, bg)
            if len(ids):
                g = t.gather()
                c2p = oc.array("subcells.cell_to_points").reshape(-1, oc.D + 1)
                assert np.array_equal(g.cell_to_points, c2p[ids - 1])
                assert np.array_equal(g.cell_to_bgcell, oc.array("subcells.cell_to_bgcell")[ids - 1])
            total, vals = integrate_measure(Measure(t, 2), return_values=True)
            ovals = oc.subcell_measures(ids)
            assert_float_equal(vals, ovals, "subcell measures")
            ototal = ovals.sum() + sum(oc.bgcell_measure(int(c)) for c in bg)
            assert abs(total - ototal) <= RTOL * max(abs(ototal), 1e-300)
            K_cut, K_bg = integrate_laplacian(Measure(t, 2))
            oK = oc.cutcell_laplacian(ids)
            assert_float_equal(K_cut, oK, "cut-cell laplacian")
            ntypes = K_bg.shape[0]
            for ty in range(ntypes):
                assert_float_equal(K_bg[ty], oc.bgcell_laplacian(ty + 1), "bg laplacian")
        ta = Triangulation(gd, ACTIVE, name)
        assert np.array_equal(ta.bg_ids, oc.select_bgcells(prog, IN, active=True))
        tb = EmbeddedBoundary(gd, name)
        fids, ori = oc.select_boundary(prog)
        assert np.array_equal(tb.sub_ids, fids)
        assert np.array_equal(tb.orientation, ori)
        surf, fvals = integrate_measure(Measure(tb, 2), return_values=True)
        assert_float_equal(fvals, oc.subfacet_measures(fids), "subfacet measures")
        if len(fids):
            g = tb.gather()
            f2n = oc.array("subfacets.facet_to_normal").reshape(-1, oc.D)
            assert_float_equal(g.facet_to_normal, f2n[fids - 1] * ori[:, None], "oriented normals")
    # two-geometry boundaries
    if len(names) > 2:
        a, b = names[1], names[2]
        tb = EmbeddedBoundary(gd, a, b)
        fids, ori = oc.select_boundary(gd.program(a), gd.program(b))
        assert np.array_equal(tb.sub_ids, fids)
        assert np.array_equal(tb.orientation, ori)


def test_known_answers_on_gpu():
    # test/InterfacesTests/EmbeddedDiscretizationsTests.jl:48-56 through the public API
    R = 0.7
    for simplex in (False, True):
        geo = disk(R)
        model = box_model(geo, (50, 50), simplex)
        cutgeo = cut(model, geo)
        vol = integrate_measure(Measure(Triangulation(cutgeo, PHYSICAL), 2))
        surf = integrate_measure(Measure(EmbeddedBoundary(cutgeo), 2))
        assert abs(np.pi * R ** 2 - vol) < 1.0e-3
        assert abs(surf - 2 * np.pi * R) < 1.0e-3


def test_discrete_geometry_matches_analytical():
    """DiscreteGeometry (NODAL values) of discretize(geo) gives the same cut as the analytical tree
    (src/LevelSetCutters/DiscreteGeometries.jl:36-63; CutTriangulations.jl:443-454)."""
    import torch
    g, pmin, pmax = stokes_geometry()
    model = CartesianDiscreteModel(pmin, pmax, (16, 5, 4))
    dgeo = gecut.discretize(g, model)
    a = cut(model, g)
    b = cut(model, dgeo)
    ha, hb = a.to_host(), b.to_host()
    for k in ha:
        assert np.array_equal(ha[k], hb[k]), k
    # host numpy nodal values too
    vals = ob.node_values(model, ob.make_leaf(0, R=0.7))
    dg = DiscreteGeometry(vals, CartesianDiscreteModel((-1, -1, -1), (1, 1, 1), (9, 8, 7)), name="ball") \
        if False else None
    model2 = CartesianDiscreteModel((-1, -1, -1), (1, 1, 1), (9, 8, 7))
    vals = ob.node_values(model2, ob.make_leaf(0, R=0.7))
    c1 = cut(model2, DiscreteGeometry(vals, model2, name="ball"))
    c2 = cut(model2, sphere(0.7))
    h1, h2 = c1.to_host(), c2.to_host()
    for k in h1:
        assert np.array_equal(h1[k], h2[k]), k
    c3 = cut(model2, DiscreteGeometry(torch.from_numpy(vals).cuda(), model2, name="ball"))
    h3 = c3.to_host()
    for k in h1:
        assert np.array_equal(h1[k], h3[k]), k


@pytest.mark.parametrize("part,simplex,slab", [((70, 45), False, None), ((33, 64), True, (5, 40)), ((65, 31, 9), False, None),
                                               ((31, 33, 12), True, (2, 11)), ((96, 8, 8), False, (0, 3)),
                                               ((5, 4, 3), True, None)])
def test_nodal_sign_words_classify_like_the_closed_form(part, simplex, slab):
    """NODAL level sets go through k_nodal_signs (one streaming pass: value vector -> sign words) and the marching kernel
    reads bits; the whole result must equal the cut of the same level sets in closed form, bit for bit -- row lengths that
    are not multiples of 32 nodes, 2-D and 3-D, n-cube and simplex cells, whole model and z-slabs (device and host
    vectors, two level sets so that chunked and NODAL rows mix with an analytical one)."""
    import torch
    D = len(part)
    if D == 2:
        g1, g2 = disk(0.62, x0=(0.05, -0.1)), disk(0.3, x0=(-0.4, 0.35))
        model = CartesianDiscreteModel((-1.0, -0.9), (1.0, 0.9), part)
    else:
        g1, g2 = sphere(0.62, x0=(0.05, -0.1, 0.02)), cylinder(0.3, x0=(-0.4, 0.35, 0.0), v=(0.0, 0.0, 1.0))
        model = CartesianDiscreteModel((-1.0, -0.9, -0.8), (1.0, 0.9, 0.8), part)
    if simplex:
        model = simplexify(model)
    geo = union(g1, g2, name="both")
    leaves, _, _ = gecut.cutters.flatten_geometry(model, geo)
    v1, v2 = gecut.cutters.evaluate_leaves_at_nodes(model, leaves, device_output=False)
    cutter = LevelSetCutter(slab=slab)
    ref = cutter.cut(model, geo).to_host()
    # both NODAL (host vectors), NODAL + closed form (device vector)
    d_both = union(DiscreteGeometry(v1, model, name="a"), DiscreteGeometry(v2, model, name="b"), name="both")
    d_mix = union(DiscreteGeometry(torch.from_numpy(v1).cuda(), model, name="a"), g2, name="both")
    for dg in (d_both, d_mix):
        got = cutter.cut(model, dg).to_host()
        for k in ref:
            assert np.array_equal(ref[k], got[k]), k
    # classification-only verb on the NODAL geometry
    assert np.array_equal(cutter.compute_bgcell_to_inoutcut(model, d_both), cutter.compute_bgcell_to_inoutcut(model, geo))


@pytest.mark.parametrize("case", ["disk_tri_50", "sphere_tet_12", "sphere_tet_odd", "sphere_hex_12"])
def test_per_entity_measures_on_demand(case):
    """Simplex bg cells with one level set do not write the per-entity measures at cut time (the sums come from the fill
    kernel's tile statistics); per-entity values, cell measures and the matrices of a complemented selection are computed
    on demand and must be the bits the oracle integrates (n-cube cells, which do write them, run the same assertions)."""
    model, geo = CASES[case]()
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    prog = gd.program(None)
    for side, sel in ((IN, PHYSICAL_IN), (OUT, PHYSICAL_OUT)):
        t = Triangulation(gd, sel)
        ids = oc.select_subcells(prog, side)
        total_fast = integrate_measure(Measure(t, 2))                     # tile statistics, no per-entity array
        total, vals = integrate_measure(Measure(t, 2), return_values=True)  # per-entity values on demand
        ref = oc.subcell_measures(ids)
        assert np.array_equal(vals, ref)
        bg = oc.select_bgcells(prog, side)
        expect = ref.sum() + sum(oc.bgcell_measure(int(c)) for c in bg)
        assert abs(total - expect) <= 1e-12 * abs(expect) and abs(total_fast - expect) <= 1e-12 * abs(expect)
        K_cut, _ = integrate_laplacian(Measure(t, 2))
        assert np.allclose(K_cut, oc.cutcell_laplacian(ids), rtol=1e-12, atol=1e-14)
    ga = EmbeddedBoundary(gd)
    fids, _ = oc.select_boundary(prog)
    tot, fvals = integrate_measure(Measure(ga, 2), return_values=True)
    assert np.array_equal(fvals, oc.subfacet_measures(fids))
    assert abs(tot - integrate_measure(Measure(ga, 2))) <= 1e-12 * abs(tot)


def test_host_function_geometry_popcorn():
    geo = popcorn()
    model = box_model(geo, (14, 14, 14))
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    compare_layout(gd, oc)


def test_node_evaluation_kernel_bit_exact():
    """K1 alone: discretize(geo,model) on the device equals the oracle's leaf formulas bit for bit."""
    model = CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (13, 11, 7))
    geo = csg_geometry()
    leaves, _, _ = gecut.cutters.flatten_geometry(model, geo)
    vals = gecut.cutters.evaluate_leaves_at_nodes(model, leaves, device_output=False)
    for lf, v in zip(leaves, vals):
        ref = ob.node_values(model, ob.make_leaf(lf.kind, lf.x0, lf.v, lf.R, lf.r))
        assert np.array_equal(v, ref)
    lf = gecut.doughnut(0.6, 0.25).get_tree().data[0]
    v = gecut.cutters.evaluate_leaves_at_nodes(model, [lf], device_output=False)[0]
    assert np.array_equal(v, ob.node_values(model, ob.make_leaf(lf.kind, lf.x0, lf.v, lf.R, lf.r)))


@pytest.mark.parametrize("case", ["sphere_hex_12", "sphere_tet_12", "csg_hex_10", "disk_quad_50"])
def test_slab_cut_is_a_slice_of_the_whole_cut(case):
    """z-slab partition (src/Distributed/DistributedDiscretizations.jl:29-40): cutting layers [k0,k1) gives exactly the
    part of the whole cut that lives in those layers (ids shifted by the slab offsets)."""
    model, geo = CASES[case]()
    whole = cut(model, geo)
    hw = whole.to_host()
    nlast = model.partition[-1]
    k0, k1 = nlast // 3, nlast - nlast // 4
    part = LevelSetCutter(slab=(k0, k1)).cut(model, geo)
    hp = part.to_host()
    off = int(part.sizes.cell_offset)
    nc = int(part.sizes.num_bgcells)
    assert np.array_equal(hp["ls_to_bgcell_to_inoutcut"], hw["ls_to_bgcell_to_inoutcut"][:, off:off + nc])
    cw = hw["cutcell_to_cell"]
    mask = (cw > off) & (cw <= off + nc)
    assert np.array_equal(hp["cutcell_to_cell"] + off, cw[mask])
    scm = (hw["sc_cell_to_bgcell"] > off) & (hw["sc_cell_to_bgcell"] <= off + nc)
    assert np.array_equal(hp["sc_cell_to_bgcell"] + off, hw["sc_cell_to_bgcell"][scm])
    first = np.flatnonzero(scm)[0] if scm.any() else 0
    p0 = hw["sc_cell_to_points"][first].min() - hp["sc_cell_to_points"][0].min() if scm.any() else 0
    assert np.array_equal(hp["sc_cell_to_points"] + p0, hw["sc_cell_to_points"][scm])
    npts = hp["sc_point_to_coords"].shape[0]
    pstart = hw["sc_cell_to_points"][scm].min() - 1 if scm.any() else 0
    assert np.array_equal(hp["sc_point_to_coords"], hw["sc_point_to_coords"][pstart:pstart + npts])
    assert np.array_equal(hp["ls_to_subcell_to_inout"], hw["ls_to_subcell_to_inout"][:, scm])


def test_error_paths():
    model = CartesianDiscreteModel((0, 0), (1, 1), (4, 4))
    with pytest.raises(ValueError):
        cut(model, sphere(0.5))  # 3-D leaf on a 2-D model
    with pytest.raises(Exception):
        CartesianDiscreteModel((0, 0), (1, 1), (4, 0))
    with pytest.raises(NotImplementedError):
        Measure(Triangulation(cut(model, disk(0.3, x0=(0.5, 0.5))), PHYSICAL), 4)
    geo = disk(0.3, x0=(0.5, 0.5))
    cg = cut(model, geo)
    with pytest.raises(KeyError):
        Triangulation(cg, PHYSICAL, "nope")


# ---- BASELINE.json full sizes: size-independent properties (the oracle would take minutes to hours there) -----------

def _measures(cg, name=None):
    om, oo, ga = Triangulation(cg, PHYSICAL_IN, name), Triangulation(cg, PHYSICAL_OUT, name), EmbeddedBoundary(cg, name)
    out = (integrate_measure(Measure(om, 2)), integrate_measure(Measure(oo, 2)), integrate_measure(Measure(ga, 2)))
    for t in (om, oo, ga):
        t.close()
    return out


def test_full_size_disk_4096():
    """BASELINE configs[1]: disk on 4096^2.  Area and perimeter converge at O(h^2) (the reference asserts 1e-3 at 50^2,
    test/InterfacesTests/EmbeddedDiscretizationsTests.jl:55-56); IN + OUT partition the box to round-off."""
    R = 0.7
    geo = disk(R)
    model = box_model(geo, (4096, 4096))
    cg = cut(model, geo)
    vin, vout, surf = _measures(cg)
    box = (2 * 1.01 * R) ** 2
    assert abs(vin + vout - box) <= 1e-12 * box
    assert abs(vin - np.pi * R ** 2) < 2e-7 and abs(surf - 2 * np.pi * R) < 2e-7
    s = cg.sizes
    assert s.num_subcells == 3 * (2 * s.num_cutcells) - 2 * (2 * s.num_cutcells - s.num_subfacets)  # 1 or 3 per triangle
    # sortedness / idempotence of the layout
    h = cg.to_host()
    assert np.all(np.diff(h["cutcell_to_cell"]) > 0) and np.all(np.diff(h["sc_cell_to_bgcell"]) >= 0)
    cg2 = cut(model, geo)
    h2 = cg2.to_host()
    assert all(np.array_equal(h[k], h2[k]) for k in h)
    cg.close(); cg2.close()


@pytest.mark.parametrize("simplex", [False, True])
def test_full_size_sphere_256(simplex):
    """BASELINE configs[2] at 256^3 (the bench runs 512^3): volume/surface convergence, partition of the box, every
    cut cell's subcells tile the cell exactly, checksums of the connectivity."""
    R = 0.7
    geo = sphere(R)
    model = box_model(geo, (256, 256, 256), simplex)
    cg = cut(model, geo)
    vin, vout, surf = _measures(cg)
    box = (2 * 1.01 * R) ** 3
    assert abs(vin + vout - box) <= 1e-12 * box
    assert abs(vin - 4 / 3 * np.pi * R ** 3) < 2e-4 and abs(surf - 4 * np.pi * R ** 2) < 1e-3  # O(h^2), h = 5.5e-3
    h = cg.to_host()
    c2p, c2bg = h["sc_cell_to_points"], h["sc_cell_to_bgcell"]
    assert c2p.min() == 1 and c2p.max() == h["sc_point_to_coords"].shape[0]
    assert np.all(np.diff(c2bg) >= 0) and np.array_equal(np.unique(c2bg), h["cutcell_to_cell"])
    # per cut cell: sum of |subcell volumes| == bg cell volume (subcells tile the cut cell)
    X = h["sc_point_to_coords"][c2p - 1]
    vol = np.abs(np.linalg.det(X[:, 1:, :] - X[:, :1, :])) / 6.0
    idx = np.searchsorted(h["cutcell_to_cell"], c2bg)
    per_cell = np.bincount(idx, weights=vol, minlength=len(h["cutcell_to_cell"]))
    cellvol = np.prod(model.sizes) / (6.0 if simplex else 1.0)
    assert np.abs(per_cell - cellvol).max() <= 1e-12 * cellvol * 8
    # unit normals point from IN to OUT: away from the sphere centre
    f2p = h["sf_facet_to_points"]
    cen = h["sf_point_to_coords"][f2p - 1].mean(1)
    nrm = h["sf_facet_to_normal"]
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-12
    assert np.all((nrm * cen).sum(1) > 0)
    cg.close()


def test_max_levelsets_and_limits():
    """GECUT_MAX_LEVELSETS = 16 unique leaves work (generic walk, LM = 16); 17 are rejected like @notimplementedif."""
    model = CartesianDiscreteModel((-1, -1), (1, 1), (24, 24))
    geos = [disk(0.25 + 0.03 * i, x0=(-0.5 + 0.06 * i, 0.1 * np.sin(i)), name=f"d{i}") for i in range(17)]
    g16 = geos[0]
    for g in geos[1:16]:
        g16 = union(g16, g)
    oc = ob.oracle_cut(model, g16)
    gd = cut(model, g16)
    assert gd.L == 16
    compare_layout(gd, oc)
    with pytest.raises(NotImplementedError):
        cut(model, union(g16, geos[16]))


# ---- widening: aggregation input, re-cut loop, slab-local nodal values, classification-only verb, raw ABI ----

@pytest.mark.parametrize("case", ["sphere_hex_12", "sphere_tet_12", "csg_hex_10", "disk_quad_50", "bimaterial"])
def test_cell_measure_matches_oracle(case):
    """get_cell_measure(trian, bgtrian) (src/AgFEM/CellAggregation.jl:111-113): per-bg-cell measure of the physical
    triangulation = whole cell for IN cells, sum of the selected subcells (subcell order) for cut cells."""
    from gecut.measures import get_cell_measure
    model, geo = CASES[case]()
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    name = named_programs(gd)[0]
    prog = gd.program(name)
    for side, sel in ((IN, PHYSICAL_IN), (OUT, PHYSICAL_OUT)):
        t = Triangulation(gd, sel, name)
        got = get_cell_measure(t)
        want = np.zeros(model.num_cells)
        for c in oc.select_bgcells(prog, side):
            want[c - 1] = oc.bgcell_measure(int(c))
        ids = oc.select_subcells(prog, side)
        if len(ids):
            bg = oc.array("subcells.cell_to_bgcell")[ids - 1]
            np.add.at(want, bg - 1, oc.subcell_measures(ids))  # unbuffered, in subcell order
        assert_float_equal(got, want, "cell measures")
        dev = get_cell_measure(t, device_output=True)
        assert np.array_equal(dev.cpu().numpy(), got)
    # the unit cut measures of IN + OUT partition every cell
    tin, tout = Triangulation(gd, PHYSICAL_IN, name), Triangulation(gd, PHYSICAL_OUT, name)
    both = get_cell_measure(tin) + get_cell_measure(tout)
    cellvol = np.array([oc.bgcell_measure(c + 1) for c in range(min(model.num_cells, 12))])
    assert np.allclose(both[:len(cellvol)], cellvol, rtol=1e-10, atol=0)


def test_recut_loop_with_device_resident_levelset():
    """Re-cut on an updated level set (src/LevelSetCutters/DifferentiableTriangulations.jl:55-59): the nodal
    values live on the device, every iteration cuts again in the same context; each result equals the oracle's and the
    context stops asking the driver for memory after the first iteration."""
    import torch
    model = CartesianDiscreteModel((-1, -1, -1), (1, 1, 1), (14, 12, 10))
    X = ob.node_coords(model)
    ctx = gecut._lib.default_context()
    launches = []
    for it, R in enumerate((0.55, 0.6, 0.65, 0.7)):
        vals = (X ** 2).sum(1) - R * R
        dgeo = DiscreteGeometry(torch.from_numpy(vals).cuda(), model, name="ball")
        cg = cut(model, dgeo)
        oc = ob.OracleCut(model, [ob.make_leaf(4)], [vals])
        compare_layout(cg, oc)
        vol = integrate_measure(Measure(Triangulation(cg, PHYSICAL, "ball"), 2))
        assert abs(vol - 4 / 3 * np.pi * R ** 3) < 0.1 * (4 / 3 * np.pi * R ** 3)  # coarse grid, quadratic level set
        launches.append(ctx.launch_count)
        cg.close()
    assert launches[-1] > launches[0]  # the native kernels ran every iteration


def test_slab_local_nodal_values():
    """A rank that owns only the node planes of its slab (+ the ghost plane) cuts the same as a rank holding the whole
    vector (src/Distributed/DistributedDiscretizations.jl:164-170: ghost values come from the neighbour)."""
    model = CartesianDiscreteModel((-1, -1, -1), (1, 1, 1), (10, 9, 12))
    vals = ob.node_values(model, ob.make_leaf(0, R=0.7))
    k0, k1 = 3, 9
    layer_nodes = model.num_nodes // (model.partition[-1] + 1)
    whole = LevelSetCutter(slab=(k0, k1)).cut(model, DiscreteGeometry(vals, model, name="ball"))
    local_vals = np.ascontiguousarray(vals[k0 * layer_nodes:(k1 + 1) * layer_nodes])
    part = LevelSetCutter(slab=(k0, k1), nodal_slab_local=True).cut(model, DiscreteGeometry(local_vals, model, name="ball"))
    hw, hp = whole.to_host(), part.to_host()
    for k in hw:
        assert np.array_equal(hw[k], hp[k]), k


@pytest.mark.parametrize("case", ["sphere_hex_12", "csg_hex_10", "stokes_tet", "disk_tri_50"])
def test_classification_only_verb(case):
    """compute_bgcell_to_inoutcut(cutter, model, geo) (LevelSetCutters.jl:154-188) without building the sub-triangulation."""
    model, geo = CASES[case]()
    oc = ob.oracle_cut(model, geo)
    got = LevelSetCutter().compute_bgcell_to_inoutcut(model, geo)
    tree = geo.get_tree()
    from gecut.csg import find_unique_leaves
    _, oid_to_ls = find_unique_leaves(tree)
    prog = compile_postfix(tree, oid_to_ls)
    assert np.array_equal(got, oc.bgcell_to_inoutcut(prog))


def test_raw_abi_error_codes():
    """Status codes and gecut_last_error through ctypes, the way a foreign binding sees them."""
    import ctypes as C
    from gecut import _lib
    L = _lib.lib()
    ctx = _lib.default_context()
    out = C.c_void_p()
    assert L.gecut_cut(None, None, None, 0, None, 0, C.byref(out)) == _lib.GECUT_ERR_INVALID
    g = _lib.Grid()
    g.D = 4
    rec = (_lib.LeafRec * 1)()
    assert L.gecut_cut(ctx.handle, C.byref(g), rec, 1, None, 0, C.byref(out)) != 0
    assert len(ctx.last_error()) > 0
    g.D = 2
    g.partition[0], g.partition[1] = 4, 4
    g.pmax[0] = g.pmax[1] = 1.0
    assert L.gecut_cut(ctx.handle, C.byref(g), rec, 17, None, 0, C.byref(out)) != 0  # more level sets than GC_LMAX
    rec[0].kind = 4  # NODAL without values
    ptrs = (C.c_void_p * 1)()
    assert L.gecut_cut(ctx.handle, C.byref(g), rec, 1, ptrs, 0, C.byref(out)) == _lib.GECUT_ERR_INVALID
    # a valid cut, then malformed selector programs
    rec[0].kind = 0
    rec[0].x0[0] = rec[0].x0[1] = 0.5
    rec[0].R = 0.3
    assert L.gecut_cut(ctx.handle, C.byref(g), rec, 1, None, 0, C.byref(out)) == 0
    sel = C.c_void_p()
    bad = (C.c_int32 * 2)(0, -2)  # binary operator with one operand
    assert L.gecut_select_cells(out, 0, bad, 2, -1, C.byref(sel)) == _lib.GECUT_ERR_INVALID
    bad2 = (C.c_int32 * 1)(3)    # level set that does not exist
    assert L.gecut_select_cells(out, 0, bad2, 1, -1, C.byref(sel)) == _lib.GECUT_ERR_INVALID
    ok = (C.c_int32 * 1)(0)
    assert L.gecut_select_cells(out, 0, ok, 1, -1, C.byref(sel)) == 0
    tot = C.c_double()
    assert L.gecut_integrate_measure(sel, None, C.byref(tot)) == 0
    assert 0.15 < tot.value < 0.35  # disk of radius 0.3 on a 4 x 4 grid: pi*0.09 = 0.283 up to the coarse polygon
    assert L.gecut_cell_measure(sel, None, 0) == _lib.GECUT_ERR_INVALID
    L.gecut_selection_free(sel)
    L.gecut_result_free(out)


@pytest.mark.parametrize("case", ["bimaterial_160", "csg_96", "stokes_tet_64"])
def test_multi_levelset_properties_at_scale(case):
    """BASELINE configs 1, 4, 5 at sizes the oracle would need minutes for: IN + OUT partition the box for every named
    geometry, the subcells of every cut cell tile it, the facets of a level set are INTERFACE for it, and the
    two-launch cut (no-tree kernel + 8-lane tree walk) agrees with the slab-by-slab cut."""
    if case == "bimaterial_160":
        geo = bimaterial()
        model = CartesianDiscreteModel((0, 0, 0), (3, 2, 3), (160, 96, 160))
        names = ["steel", "concrete"]
    elif case == "csg_96":
        geo = csg_geometry()
        model = CartesianDiscreteModel((-0.8, -0.8, -0.8), (0.8, 0.8, 0.8), (96, 96, 96))
        names = ["csg"]
    else:
        geo, pmin, pmax = stokes_geometry()
        model = simplexify(CartesianDiscreteModel(pmin, pmax, (128, 40, 40)))
        names = ["fluid"]
    cg = cut(model, geo)
    box = float(np.prod(np.asarray(model.pmax) - np.asarray(model.pmin)))
    for nm in names:
        vin, vout, surf = _measures(cg, nm)
        assert abs(vin + vout - box) <= 1e-11 * box, (nm, vin, vout)
        assert surf > 0
    h = cg.to_host()
    c2p, c2bg = h["sc_cell_to_points"], h["sc_cell_to_bgcell"]
    assert c2p.min() == 1 and c2p.max() == h["sc_point_to_coords"].shape[0]
    assert np.all(np.diff(c2bg) >= 0) and np.array_equal(np.unique(c2bg), h["cutcell_to_cell"])
    X = h["sc_point_to_coords"][c2p - 1]
    vol = np.abs(np.linalg.det(X[:, 1:, :] - X[:, :1, :])) / 6.0
    idx = np.searchsorted(h["cutcell_to_cell"], c2bg)
    per_cell = np.bincount(idx, weights=vol, minlength=len(h["cutcell_to_cell"]))
    cellvol = np.prod(model.sizes) / (6.0 if model.simplexified else 1.0)
    assert np.abs(per_cell - cellvol).max() <= 1e-11 * cellvol * 8
    # ls-major facet order, INTERFACE row
    fst = h["ls_to_subfacet_to_inout"]
    own = np.argmax(fst == 0, axis=0)
    assert np.all(fst[own, np.arange(fst.shape[1])] == 0) and np.all(np.diff(own) >= 0)
    # unit normals, except on degenerate facets (a grid node exactly on a surface gives a cut point at coefficient 0): the
    # reference returns the zero vector when the cross product is shorter than eps (LookupTables.jl:309-360)
    nn = np.linalg.norm(h["sf_facet_to_normal"], axis=1)
    P = h["sf_point_to_coords"][h["sf_facet_to_points"] - 1]
    area = 0.5 * np.linalg.norm(np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]), axis=1)
    assert np.all((np.abs(nn - 1) < 1e-12) | ((nn == 0) & (area < 1e-15)))
    # slab by slab == whole (the recorded-simplex lists differ per slab, the layout must not)
    nlast = model.partition[-1]
    k0 = nlast // 2
    parts = [LevelSetCutter(slab=(0, k0)).cut(model, geo), LevelSetCutter(slab=(k0, nlast)).cut(model, geo)]
    hp = [p.to_host() for p in parts]
    assert hp[0]["sc_cell_to_points"].shape[0] + hp[1]["sc_cell_to_points"].shape[0] == c2p.shape[0]
    assert np.array_equal(np.concatenate([hp[0]["sc_point_to_coords"], hp[1]["sc_point_to_coords"]]), h["sc_point_to_coords"])
    assert np.array_equal(np.concatenate([hp[0]["ls_to_subcell_to_inout"], hp[1]["ls_to_subcell_to_inout"]], axis=1),
                          h["ls_to_subcell_to_inout"])
    assert np.array_equal(np.concatenate([hp[0]["sf_facet_to_normal"][hp[0]["ls_to_subfacet_to_inout"][0] == 0],
                                          hp[1]["sf_facet_to_normal"][hp[1]["ls_to_subfacet_to_inout"][0] == 0]]),
                          h["sf_facet_to_normal"][fst[0] == 0])
    for p in parts:
        p.close()
    cg.close()


# ---- BASELINE.json configurations at their own sizes: the whole layout against the oracle ------------------------------
# (one-ulp sign flips of nodes lying on a surface are what grows with n, so the sizes matter)

BASELINE_CASES = {
    # configs[0]: PoissonCSGCutFEM.jl:9-23, test/LevelSetCuttersTests/AnalyticalGeometriesTests.jl:147-166 (40^3, L = 10)
    "c1_csg_40": lambda: (CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (40, 40, 40)), csg_geometry()),
    # configs[1]: disk on 4096^2 quads
    "c2_disk_4096": lambda: (box_model(disk(0.7), (4096, 4096)), disk(0.7)),
    # configs[2] at 256^3 (the oracle needs ~6 s and 2.3 GB there; 512^3 = the bench size would need 20 GB of host memory)
    "c3_sphere_hex_256": lambda: (box_model(sphere(0.7), (256, 256, 256)), sphere(0.7)),
    "c3_sphere_tet_256": lambda: (box_model(sphere(0.7), (256, 256, 256), True), sphere(0.7)),
    # configs[4] geometry (two cylinders, union(steel, !steel)) at 256^3 per GPU
    "c5_bimaterial_256": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (256, 256, 256)), bimaterial()),
}


@pytest.mark.parametrize("case", sorted(BASELINE_CASES))
def test_baseline_size_layout_matches_oracle(case):
    model, geo = BASELINE_CASES[case]()
    oc = ob.oracle_cut(model, geo)
    gd = cut(model, geo)
    compare_layout(gd, oc)
    # the integrals of the step the benchmark times, on the same result
    for name in named_programs(gd)[:2]:
        prog = gd.program(name)
        for side, sel in ((IN, PHYSICAL_IN), (OUT, PHYSICAL_OUT)):
            t = Triangulation(gd, sel, name)
            ids = oc.select_subcells(prog, side)
            assert np.array_equal(t.sub_ids, ids)
            bg = oc.select_bgcells(prog, side)
            assert t.num_bg == len(bg)
            total = integrate_measure(Measure(t, 2))
            ntypes = 1 if not model.simplexified else (2 if model.D == 2 else 6)  # cell type = (cell - 1) mod ntypes
            ototal = oc.subcell_measures(ids).sum() + sum(
                np.count_nonzero((bg - 1) % ntypes == ty) * oc.bgcell_measure(ty + 1) for ty in range(ntypes))
            assert abs(total - ototal) <= 1e-12 * abs(ototal)
            K_cut, _ = integrate_laplacian(Measure(t, 2))
            assert_float_equal(K_cut, oc.cutcell_laplacian(ids), "cut-cell laplacian")
            t.close()
        tb = EmbeddedBoundary(gd, name)
        fids, ori = oc.select_boundary(prog)
        assert np.array_equal(tb.sub_ids, fids) and np.array_equal(tb.orientation, ori)
        surf = integrate_measure(Measure(tb, 2))
        osurf = oc.subfacet_measures(fids).sum()
        assert abs(surf - osurf) <= 1e-12 * abs(osurf)
        tb.close()
    gd.close()


def test_baseline_c4_stokes_discrete_geometry_256():
    """configs[3]: StokesTubeWithObstacleCutFEM tube + obstacle as a DiscreteGeometry (Q1 nodal values, L = 4) on the
    simplexified 256^3 model: PHYSICAL "fluid" subcells, EmbeddedBoundary("fluid","inlet") / ("fluid","solid") with
    normals -- the whole layout and the selections against the oracle (StokesTubeWithObstacleCutFEM.jl:17-28,48-60)."""
    g, pmin, pmax = stokes_geometry()
    model = simplexify(CartesianDiscreteModel(pmin, pmax, (256, 256, 256)))
    dgeo = gecut.discretize(g, model)  # nodal values evaluated on the device (bit-exact, test_node_evaluation_kernel_bit_exact)
    gd = cut(model, dgeo)
    oc = ob.oracle_cut(model, g)
    compare_layout(gd, oc)
    prog = gd.program("fluid")
    t = Triangulation(gd, PHYSICAL, "fluid")
    assert np.array_equal(t.sub_ids, oc.select_subcells(prog, IN))
    assert np.array_equal(t.bg_ids, oc.select_bgcells(prog, IN))
    for other in ("inlet", "solid"):
        tb = EmbeddedBoundary(gd, "fluid", other)
        fids, ori = oc.select_boundary(prog, gd.program(other))
        assert np.array_equal(tb.sub_ids, fids) and np.array_equal(tb.orientation, ori)
        if len(fids):
            f2n = oc.array("subfacets.facet_to_normal").reshape(-1, 3)
            assert_float_equal(tb.gather().facet_to_normal, f2n[fids - 1] * ori[:, None], "oriented normals")
        tb.close()
    t.close()
    gd.close()


@pytest.mark.gpu
def test_zero_copy_pinned_destinations():
    """gecut_result_copy straight into page-locked memory from gecut_host_alloc (what the Julia shim wraps with
    unsafe_wrap into SubCellData / SubFacetData) gives the same arrays as the pageable path; gecut_host_register pins memory
    the host language owns."""
    import ctypes as C
    from gecut import _lib
    model, geo = CASES["sphere_tet_12"]()
    gd = cut(model, geo)
    ref = gd.to_host()
    s = gd.sizes
    D, L = int(s.D), int(s.num_levelsets)
    shapes = {"ls_to_bgcell_to_inoutcut": ((L, s.num_bgcells), np.int8), "cutcell_to_cell": ((s.num_cutcells,), np.int32),
              "sc_cell_to_points": ((s.num_subcells, D + 1), np.int32), "sc_cell_to_bgcell": ((s.num_subcells,), np.int32),
              "sc_point_to_coords": ((s.num_subcell_points, D), np.float64),
              "sc_point_to_rcoords": ((s.num_subcell_points, D), np.float64),
              "ls_to_subcell_to_inout": ((L, s.num_subcells), np.int8),
              "sf_facet_to_points": ((s.num_subfacets, D), np.int32), "sf_facet_to_bgcell": ((s.num_subfacets,), np.int32),
              "sf_facet_to_normal": ((s.num_subfacets, D), np.float64),
              "ls_to_subfacet_to_inout": ((L, s.num_subfacets), np.int8)}
    bufs = {k: _lib.pinned_empty(tuple(int(x) for x in shp), dt) for k, (shp, dt) in shapes.items()}
    got = gd.to_host(buffers=bufs)
    for k in shapes:
        assert got[k] is bufs[k] or np.shares_memory(got[k], bufs[k]), k
        assert np.array_equal(bufs[k], ref[k]), k
    # memory owned by the host language, pinned in place
    own = np.empty(int(s.num_subcell_points) * D, dtype=np.float64)
    lib = _lib.lib()
    assert lib.gecut_host_register(C.c_void_p(own.ctypes.data), C.c_size_t(own.nbytes)) == 0
    try:
        got2 = gd.to_host(buffers={"sc_point_to_coords": own.reshape(-1, D)})
        assert np.array_equal(got2["sc_point_to_coords"], ref["sc_point_to_coords"])
    finally:
        assert lib.gecut_host_unregister(C.c_void_p(own.ctypes.data)) == 0
    assert lib.gecut_host_register(None, 0) != 0
    gd.close()


@pytest.mark.gpu
@pytest.mark.parametrize("simplex", [False, True])
def test_classification_fuzz_against_oracle(simplex):
    """Interval culling of the classification (k_classify_cull): random spheres, cylinders (axis-aligned and oblique) and
    planes, singly and in pairs, on models several tiles wide -- every state of every level set equals the oracle's, and so
    does the cut-cell list (the tiles the culling hands to the marching kernel are exactly the ones that need it)."""
    rng = np.random.default_rng(20261020 + int(simplex))
    base = CartesianDiscreteModel((-1.0, -0.5, -0.25), (1.0, 0.5, 0.25), (160, 40, 24))
    model = simplexify(base) if simplex else base

    def rand_leaf():
        k = rng.integers(0, 4)
        c = tuple(rng.uniform(lo, hi) for lo, hi in ((-1.1, 1.1), (-0.6, 0.6), (-0.3, 0.3)))
        if k == 0:
            return sphere(float(rng.uniform(0.05, 0.8)), x0=c)
        if k == 1:
            v = [(1, 0, 0), (0, 1, 0), (0, 0, 1)][rng.integers(0, 3)]
            return cylinder(float(rng.uniform(0.05, 0.5)), x0=c, v=v)
        if k == 2:
            return cylinder(float(rng.uniform(0.05, 0.5)), x0=c, v=tuple(rng.normal(size=3)))
        return plane(x0=c, v=tuple(rng.normal(size=3)))

    for trial in range(12):
        geo = rand_leaf() if trial % 2 == 0 else union(rand_leaf(), rand_leaf())
        oc = ob.oracle_cut(model, geo)
        gd = LevelSetCutter().cut(model, geo)
        h = gd.to_host()
        assert np.array_equal(h["ls_to_bgcell_to_inoutcut"], oc.rows("ls_to_bgcell_to_inoutcut")), trial
        assert np.array_equal(h["cutcell_to_cell"], oc.array("cutcell_to_cell")), trial
        assert np.array_equal(h["sc_cell_to_points"].ravel(), oc.array("subcells.cell_to_points")), trial
        gd.close()


@pytest.mark.gpu
def test_closed_form_cell_measures_partition_the_cell_at_scale():
    """k_cut1_fill on simplex cells computes the (IN, OUT) measure of every cut cell in closed form from the cut fractions:
    at 192^3 (1.1 M cut TETs) the two P1 element matrices of every cut cell add up to the matrix of the whole TET
    (K = G V with the same gradient products) to 1e-12, i.e. V_in + V_out = |cell| cell by cell, and the volume totals
    equal the sums of the per-entity measures computed from the stored coordinates (the reference's arithmetic)."""
    from gecut.measures import integrate_measure, integrate_laplacian
    geo = sphere(0.7)
    model = box_model(geo, (192, 192, 192), True)
    gd = cut(model, geo)
    tin, tout = Triangulation(gd, PHYSICAL_IN), Triangulation(gd, PHYSICAL_OUT)
    Kin, Kbg = integrate_laplacian(Measure(tin, 2))
    Kout, _ = integrate_laplacian(Measure(tout, 2))
    cells = gd.to_host()["cutcell_to_cell"]
    whole = Kbg[(cells - 1) % 6]  # matrices of the uncut TET types
    assert np.allclose(Kin + Kout, whole, rtol=1e-12, atol=1e-12 * np.abs(whole).max())
    for t in (tin, tout):
        total_fast = integrate_measure(Measure(t, 2))
        vals = integrate_measure(Measure(t, 2), return_values=True)[1]
        sub = float(np.sum(vals))
        nbg = t.num_bg
        cellvol = np.prod([(b - a) / n for a, b, n in zip(model.pmin, model.pmax, model.partition)]) / 6.0
        assert abs((sub + nbg * cellvol) - total_fast) <= 1e-11 * abs(total_fast)
    for o in (tin, tout, gd):
        o.close()
