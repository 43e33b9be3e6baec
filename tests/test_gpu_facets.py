"""GPU parity of `cut_facets` (EmbeddedFacetDiscretization): the CUDA path through the C ABI
against the CPU oracle on the same inputs -- bit-exact states, facet numbering, connectivity and ordering; coordinates
compared for equality first and within 1e-12 relative otherwise -- plus size-independent properties at scale."""
import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, disk, sphere, cube, cylinder, plane, union, intersect, setdiff, complement,
                   get_geometry, DiscreteGeometry, cut, cut_facets, compute_bgfacet_to_inoutcut,
                   compute_subfacet_to_inout, BoundaryTriangulation, SkeletonTriangulation, Triangulation,
                   EmbeddedBoundary, Measure, LevelSetCutter, PHYSICAL, PHYSICAL_IN, PHYSICAL_OUT, ACTIVE, ACTIVE_OUT,
                   CUT_IN, CUT_OUT, IN, OUT)
from gecut import _lib
from gecut.measures import integrate_measure
import oracle_binding as ob
from test_gpu_parity import csg_geometry, stokes_geometry, three_disks, bimaterial

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def sx(model):
    return gecut.simplexify(model)


def two_disks_plane():
    return setdiff(intersect(disk(0.45, x0=(0.5, 0.5)), plane(x0=(0.5, 0.7), v=(0.2, 1.0))), disk(0.2, x0=(0.4, 0.4)))


def stokes_model(n):
    geo, pmin, pmax = stokes_geometry()
    return CartesianDiscreteModel(pmin, pmax, n), geo


CASES = {
    "disk_corner_10": lambda: (CartesianDiscreteModel((0, 1, 0, 1), (10, 10)), disk(0.72, x0=(1, 1))),
    "disk_odd": lambda: (CartesianDiscreteModel((-0.8, 0.9, -0.7, 0.75), (37, 45)), disk(0.7)),
    "disk_1xN": lambda: (CartesianDiscreteModel((-1, 1, -1, 1), (1, 7)), disk(0.7)),
    "three_disks": lambda: (CartesianDiscreteModel((-1, 2, -1, 2), (23, 19)), three_disks()),
    "csg2d": lambda: (CartesianDiscreteModel((0, 1, 0, 1), (33, 31)), two_disks_plane()),
    "sphere_face": lambda: (CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (8, 8, 8)), sphere(0.37, x0=(0.5, 0.5, 1.0))),
    "sphere_odd": lambda: (CartesianDiscreteModel((-0.8, 0.9, -0.7, 0.75, -0.9, 0.8), (13, 9, 11)), sphere(0.7)),
    "sphere_1x1xN": lambda: (CartesianDiscreteModel((-0.2, 0.3, -0.2, 0.3, -1, 1), (1, 1, 9)), sphere(0.7)),
    "csg_10": lambda: (CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (10, 10, 10)), csg_geometry()),
    "stokes": lambda: stokes_model((20, 8, 8)),
    "bimaterial": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (9, 6, 9)), bimaterial()),
    "empty_cut": lambda: (CartesianDiscreteModel((2, 3, 2, 3), (5, 5)), disk(0.7)),
    "all_in": lambda: (CartesianDiscreteModel((-0.1, 0.1, -0.1, 0.1, -0.1, 0.1), (3, 3, 3)), sphere(0.7)),
    # simplexified models: the facet grid of the TRI / TET mesh (SEGMENT / TRI facets)
    "sx_disk_odd": lambda: (sx(CartesianDiscreteModel((-0.8, 0.9, -0.7, 0.75), (17, 13))), disk(0.7)),
    "sx_disk_1xN": lambda: (sx(CartesianDiscreteModel((-1, 1, -1, 1), (1, 7))), disk(0.7)),
    "sx_three_disks": lambda: (sx(CartesianDiscreteModel((-1, 2, -1, 2), (23, 19))), three_disks()),
    "sx_csg2d": lambda: (sx(CartesianDiscreteModel((0, 1, 0, 1), (33, 31))), two_disks_plane()),
    "sx_sphere_odd": lambda: (sx(CartesianDiscreteModel((-0.8, 0.9, -0.7, 0.75, -0.9, 0.8), (7, 5, 6))), sphere(0.7)),
    "sx_sphere_face": lambda: (sx(CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (6, 6, 6))), sphere(0.37, x0=(0.5, 0.5, 1.0))),
    "sx_sphere_1x1xN": lambda: (sx(CartesianDiscreteModel((-0.2, 0.3, -0.2, 0.3, -1, 1), (1, 1, 9))), sphere(0.7)),
    "sx_csg_8": lambda: (sx(CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (8, 8, 8))), csg_geometry()),
    "sx_empty_cut": lambda: (sx(CartesianDiscreteModel((2, 3, 2, 3), (5, 5))), disk(0.7)),
}


def assert_close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    if not np.array_equal(a, b):
        assert np.allclose(a, b, rtol=RTOL, atol=1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_facet_layout_parity(name):
    model, geo = CASES[name]()
    fc = cut_facets(model, geo)
    oc = ob.oracle_cut_facets(model, geo)
    h = fc.to_host(topology=True)
    s = fc.sizes
    assert (s.num_bgfacets, s.num_cutfacets, s.num_subfacets, s.num_subfacet_points, s.num_boundary_facets) == \
        (oc.Nf, oc.Ncutf, oc.Nsf, oc.Nsfp, oc.Nboundary)
    D = model.D
    assert np.array_equal(h["facet_to_nodes"], oc.array("facet_nodes").reshape(-1, oc.nvf) + 1)
    assert np.array_equal(h["facet_is_boundary"], oc.array("facet_is_boundary"))
    assert np.array_equal(h["ls_to_facet_to_inoutcut"], oc.rows("ls_to_facet_to_inoutcut"))
    assert np.array_equal(h["cutfacet_to_facet"], oc.array("cutfacet_to_facet"))
    assert np.array_equal(h["sf_cell_to_points"].ravel(), oc.array("subfacets.cell_to_points"))
    assert np.array_equal(h["sf_cell_to_bgcell"], oc.array("subfacets.cell_to_bgcell"))
    assert np.array_equal(h["ls_to_subfacet_to_inout"], oc.rows("ls_to_subfacet_to_inout").reshape(oc.L, -1))
    assert_close(h["sf_point_to_coords"].ravel(), oc.array("subfacets.point_to_coords"))
    assert_close(h["sf_point_to_rcoords"].ravel(), oc.array("subfacets.point_to_rcoords"))
    fc.close()


@pytest.mark.parametrize("name", ["disk_corner_10", "three_disks", "csg2d", "sphere_face", "csg_10", "stokes", "bimaterial",
                                  "sx_three_disks", "sx_csg2d", "sx_sphere_face", "sx_csg_8"])
def test_facet_selectors_and_measures(name):
    model, geo = CASES[name]()
    fc = cut_facets(model, geo)
    oc = ob.oracle_cut_facets(model, geo)
    names = [None]
    if name in ("csg_10", "sx_csg_8"):
        names += ["csg", "source"]
    if name == "stokes":
        names += ["fluid", "solid"]
    if name == "bimaterial":
        names += ["steel", "concrete"]
    if name in ("three_disks", "sx_three_disks"):
        names += ["disk1", "disk3"]
    for nm in names:
        prog = fc.program(nm)
        assert np.array_equal(compute_bgfacet_to_inoutcut(fc, nm), oc.bgfacet_to_inoutcut(prog))
        assert np.array_equal(compute_subfacet_to_inout(fc, nm), oc.subfacet_to_inout(prog))
        for ctor, which in ((BoundaryTriangulation, ob.BOUNDARY_FACETS), (SkeletonTriangulation, ob.SKELETON_FACETS)):
            for sel, side in ((PHYSICAL_IN, IN), (PHYSICAL_OUT, OUT)):
                t = ctor(fc, sel, nm)
                sub = oc.select(prog, which, ob.SEL_CUT, side)
                bg = oc.select(prog, which, ob.SEL_BG, side)
                assert np.array_equal(t.sub_ids, sub)
                assert np.array_equal(t.bg_ids, bg)
                tot, vals = t.measure(return_values=True)
                ovals = oc.subfacet_measures(sub)
                assert_close(vals, ovals)
                ref = ovals.sum() + sum(oc.bgfacet_measure(int(f)) for f in bg)
                assert abs(tot - ref) <= 1e-12 * max(abs(ref), 1e-300)
                t.close()
            t = ctor(fc, ACTIVE, nm)
            assert t.num_sub == 0 and np.array_equal(t.bg_ids, oc.select(prog, which, ob.SEL_ACTIVE, IN))
            t.close()
            t = ctor(fc, CUT_OUT, nm)
            assert t.num_bg == 0 and np.array_equal(t.sub_ids, oc.select(prog, which, ob.SEL_CUT, OUT))
            t.close()
            t = ctor(fc, OUT, nm)
            assert t.num_sub == 0 and np.array_equal(t.bg_ids, oc.select(prog, which, ob.SEL_BG, OUT))
            t.close()
    fc.close()


def test_facets_discrete_geometry_and_classify_only():
    model, geo = stokes_model((12, 6, 6))
    dgeo = gecut.discretize(geo, model)
    fc = cut_facets(model, dgeo)
    fa = cut_facets(model, geo)
    ha, hd = fa.to_host(), fc.to_host()
    for k in ha:
        assert np.array_equal(ha[k], hd[k]), k
    st = compute_bgfacet_to_inoutcut(model, geo)  # classification-only entry point
    assert np.array_equal(st, compute_bgfacet_to_inoutcut(fa, geo))
    # cut_facets(cut::EmbeddedDiscretization, geo) and cut_facets(cut_facets) forwards (LevelSetCutters.jl:122-128)
    cg = cut(model, geo)
    f2 = cut_facets(cg, geo)
    assert np.array_equal(f2.to_host()["sf_cell_to_points"], ha["sf_cell_to_points"])
    assert cut_facets(f2) is f2
    for x in (fc, fa, f2, cg):
        x.close()


def test_facets_unsupported_inputs():
    m = CartesianDiscreteModel((0, 1, 0, 1), (4, 4))
    with pytest.raises(_lib.GecutError):  # slabs are supported; an empty / out-of-range one is an invalid argument
        LevelSetCutter(slab=(3, 9)).cut_facets(m, disk(0.3, x0=(0.5, 0.5)))


@pytest.mark.parametrize("case", ["disk_2048", "sphere_160", "csg_64", "sx_disk_1024", "sx_sphere_96"])
def test_facets_properties_at_scale(case):
    """Size-independent properties where the oracle is too slow: IN + OUT parts of every cut facet add up to the facet,
    boundary/skeleton split, agreement of the facet states with the cell states of cut(), boundary IN measure against
    the analytic value."""
    if case == "disk_2048":
        model, geo = CartesianDiscreteModel((0, 1, 0, 1), (2048, 2048)), disk(0.72, x0=(1, 1))
    elif case == "sphere_160":
        model, geo = CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (160, 160, 160)), sphere(0.37, x0=(0.5, 0.5, 1.0))
    elif case == "sx_disk_1024":
        model, geo = sx(CartesianDiscreteModel((0, 1, 0, 1), (1024, 1024))), disk(0.72, x0=(1, 1))
    elif case == "sx_sphere_96":
        model, geo = sx(CartesianDiscreteModel((0, 1, 0, 1, 0, 1), (96, 96, 96))), sphere(0.37, x0=(0.5, 0.5, 1.0))
    else:
        model, geo = CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (64, 64, 64)), csg_geometry()
    D = model.D
    fc = cut_facets(model, geo)
    nm = "csg" if case == "csg_64" else None
    h = np.asarray(model.sizes)
    n = np.asarray(model.partition)
    tot_b = {}
    for ctor, key in ((BoundaryTriangulation, "b"), (SkeletonTriangulation, "s")):
        tin = ctor(fc, PHYSICAL_IN, nm)
        tout = ctor(fc, PHYSICAL_OUT, nm)
        m_in, m_out = tin.measure(), tout.measure()
        # total measure of the boundary / interior facets of the box
        if key == "b":
            whole = 2 * sum(np.prod(np.delete(h * n, d)) for d in range(D))
        else:
            whole = sum((n[d] - 1) * np.prod(np.delete(h * n, d)) for d in range(D))
            if model.simplexified:  # + the facets inside every cube: the diagonal (2-D); 6 triangles around the edge 1-6 (3-D)
                if D == 2:
                    whole += np.prod(n) * np.hypot(h[0], h[1])
                else:
                    e = np.array([-h[0], h[1], h[2]])  # cube vertex 1 -> vertex 6
                    others = [np.array(v) * h for v in ((-1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 1, 0), (-1, 0, 1), (0, 1, 1))]
                    whole += np.prod(n) * sum(0.5 * np.linalg.norm(np.cross(e, o)) for o in others)
        assert abs(m_in + m_out - whole) <= 1e-11 * whole
        assert tin.num_sub + tout.num_sub > 0 or (key == "b" and case == "csg_64")  # the CSG body stays inside the box
        tot_b[key] = m_in
        tin.close()
        tout.close()
    if case == "sx_disk_1024":
        assert abs(tot_b["b"] - 2 * 0.72) < 1e-6
    if case == "sx_sphere_96":
        assert abs(tot_b["b"] - np.pi * 0.37 ** 2) / (np.pi * 0.37 ** 2) < 5e-3
    if case == "disk_2048":  # quarter disk of radius 0.72 centred on the corner: two boundary segments of length R
        assert abs(tot_b["b"] - 2 * 0.72) < 1e-6
    if case == "sphere_160":  # disk of radius R on the z-max face
        assert abs(tot_b["b"] - np.pi * 0.37 ** 2) / (np.pi * 0.37 ** 2) < 2e-3
    # sub-facets are grouped by ascending cut facet; ids within range; states +-1
    hh = fc.to_host()
    c2bg = hh["sf_cell_to_bgcell"]
    assert np.all(np.diff(c2bg) >= 0)
    assert np.array_equal(np.unique(c2bg), hh["cutfacet_to_facet"])
    assert np.all(np.diff(hh["cutfacet_to_facet"]) > 0)
    assert hh["sf_cell_to_points"].min() == 1 and hh["sf_cell_to_points"].max() == fc.sizes.num_subfacet_points
    assert set(np.unique(hh["ls_to_subfacet_to_inout"]).tolist()) <= {-1, 1}
    # every cut facet belongs to a facet the classification calls CUT for some level set
    anycut = (hh["ls_to_facet_to_inoutcut"] == 0).any(0)
    assert np.array_equal(np.nonzero(anycut)[0] + 1, hh["cutfacet_to_facet"])
    fc.close()


@pytest.mark.parametrize("name,nslab", [("disk_odd", 3), ("three_disks", 2), ("sphere_odd", 3), ("csg_10", 2), ("bimaterial", 3),
                                        ("sx_csg2d", 3), ("sx_sphere_odd", 2), ("sx_csg_8", 3)])
def test_facet_slab_is_the_local_model_of_the_whole_cut(name, nslab):
    """z-slabs (multi-GPU partition): cut_facets with `slab=(k0, k1)` is the facet discretization of the rank's LOCAL model
    (local numbering) on GLOBAL node coordinates: every array equals the whole-model result on the facets of the slab, bit
    for bit, through the facet -> node-tuple map."""
    model, geo = CASES[name]()
    D = model.D
    ctx = _lib.default_context()
    whole = cut_facets(model, geo)
    hw = whole.to_host(topology=True)
    gnodes = {tuple(r): f for f, r in enumerate(hw["facet_to_nodes"])}
    # sub-facets of every cut facet of the whole model (contiguous, in cut-facet order)
    wsf = {}
    for i, f in enumerate(hw["sf_cell_to_bgcell"]):
        wsf.setdefault(int(f) - 1, []).append(i)
    nz = model.partition[-1]
    layer_nodes = model.num_nodes // (nz + 1)
    cuts = [round(q * nz / nslab) for q in range(nslab + 1)]
    total_cutf = 0
    for k0, k1 in zip(cuts[:-1], cuts[1:]):
        loc = LevelSetCutter(ctx, slab=(k0, k1)).cut_facets(model, geo)
        hl = loc.to_host(topology=True)
        assert loc.bgmodel.partition[-1] == k1 - k0
        # local facet -> global facet through the (ordered) node tuple
        l2g = np.array([gnodes[tuple(int(v) + k0 * layer_nodes for v in r)] for r in hl["facet_to_nodes"]], dtype=np.int64)
        assert len(set(l2g.tolist())) == len(l2g)
        assert np.array_equal(hl["ls_to_facet_to_inoutcut"], hw["ls_to_facet_to_inoutcut"][:, l2g])
        lcut = hl["cutfacet_to_facet"] - 1
        assert np.all(np.diff(lcut) > 0)
        gcut = set((hw["cutfacet_to_facet"] - 1).tolist())
        assert set(l2g[lcut].tolist()) == {f for f in gcut if f in set(l2g.tolist())}
        # sub-facets: per cut facet the same simplices, points, reference points and states
        lsf = {}
        for i, f in enumerate(hl["sf_cell_to_bgcell"]):
            lsf.setdefault(int(f) - 1, []).append(i)
        assert sorted(lsf) == sorted(lcut.tolist())
        for f, ids in lsf.items():
            gids = wsf[int(l2g[f])]
            assert len(ids) == len(gids)
            a, b = np.asarray(ids), np.asarray(gids)
            assert np.array_equal(hl["ls_to_subfacet_to_inout"][:, a], hw["ls_to_subfacet_to_inout"][:, b])
            pa, pb = hl["sf_cell_to_points"][a] - 1, hw["sf_cell_to_points"][b] - 1
            assert np.array_equal(pa - pa.min(), pb - pb.min())
            assert np.array_equal(hl["sf_point_to_coords"][pa.ravel()], hw["sf_point_to_coords"][pb.ravel()])
            assert np.array_equal(hl["sf_point_to_rcoords"][pa.ravel()], hw["sf_point_to_rcoords"][pb.ravel()])
        total_cutf += len(lcut)
        # the slab's lower / upper planes are boundary facets of the local model
        assert int(loc.sizes.num_boundary_facets) == int(hl["facet_is_boundary"].sum())
        loc.close()
    # every cut facet of the model shows up in the slabs; those on the planes shared by two slabs in both
    shared = 0
    for kk in cuts[1:-1]:
        on_plane = [f for f in (hw["cutfacet_to_facet"] - 1)
                    if all((int(v) - 1) // layer_nodes == kk for v in hw["facet_to_nodes"][f])]
        shared += len(on_plane)
    assert total_cutf == len(hw["cutfacet_to_facet"]) + shared
    whole.close()


def test_facet_slab_with_slab_local_nodal_values():
    """DiscreteGeometry given as slab-local nodal vectors (node planes k0..k1 only): same facets as the closed-form slab."""
    from gecut.cutters import evaluate_leaves_at_nodes
    from gecut.csg import find_unique_leaves, replace_data
    model, geo = CASES["sphere_odd"]()
    ctx = _lib.default_context()
    k0, k1 = 3, 8
    ref = LevelSetCutter(ctx, slab=(k0, k1)).cut_facets(model, geo)
    hr = ref.to_host(topology=True)
    tree = geo.get_tree()
    payloads, oid_to_j = find_unique_leaves(tree)
    bufs = [evaluate_leaves_at_nodes(model, [pl], ctx=ctx, device_output=True, slab=(k0, k1))[0] for pl in payloads]
    dgeo = DiscreteGeometry(replace_data(lambda d: d, lambda d: (bufs[oid_to_j[id(d[0])]], d[1], d[2]), tree), model)
    loc = LevelSetCutter(ctx, slab=(k0, k1), nodal_slab_local=True).cut_facets(model, dgeo)
    hl = loc.to_host(topology=True)
    for k in ("ls_to_facet_to_inoutcut", "cutfacet_to_facet", "sf_cell_to_points", "sf_cell_to_bgcell", "sf_point_to_coords",
              "sf_point_to_rcoords", "ls_to_subfacet_to_inout"):
        assert np.array_equal(hl[k], hr[k]), k
    ref.close()
    loc.close()
