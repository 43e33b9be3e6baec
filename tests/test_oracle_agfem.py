"""Oracle checks of the AgFEM aggregation restatement (`src/AgFEM/CellAggregation.jl:104-241`): the reference's own test
(`test/AgFEMTests/CellAggregationTests.jl`) only asserts that the call paths agree, so the invariants of the algorithm
are asserted here as well."""
import numpy as np
import pytest

import gecut
from gecut import CartesianDiscreteModel, disk, sphere, union, get_geometry
from gecut.csg import compile_postfix
import oracle_binding as ob
import facet_helpers as FH
from test_oracle_kat import csg_geometry

IN, OUT, CUT = -1, 1, 0


def setup(model, geo):
    oc = ob.oracle_cut(model, geo)
    fc = ob.oracle_cut_facets(model, geo)
    return oc, fc


def check_invariants(model, oc, fc, p, loc, thr, agg):
    st = oc.bgcell_to_inoutcut(p)
    n = model.num_cells
    cells = np.arange(1, n + 1)
    # cells on the other side are not aggregated (threshold > 0); cells on `loc` are their own roots
    assert np.all(agg[st == -loc] == 0)
    assert np.all(agg[st == loc] == cells[st == loc])
    # every CUT cell has a root, and a root is its own root
    assert np.all(agg[st == CUT] > 0)
    roots = np.unique(agg[agg > 0])
    assert np.all(agg[roots - 1] == roots)
    # roots that are CUT cells have a unit cut measure >= threshold; CUT cells below it are slaves
    ids = oc.select_subcells(p, loc)
    vol = np.zeros(n)
    np.add.at(vol, oc.array("subcells.cell_to_bgcell")[ids - 1] - 1, oc.subcell_measures(ids))
    unit = vol / np.prod(model.sizes)
    cut_roots = roots[st[roots - 1] == CUT]
    assert np.all(unit[cut_roots - 1] >= thr - 1e-12)
    slaves = cells[(st == CUT) & (agg != cells)]
    assert np.all(unit[slaves - 1] < thr + 1e-12)
    # a slave is connected to its root through face neighbours with the same root (or the root itself)
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    d, pos, minus, plus = FH.facet_cells(model, fn)
    m = (minus >= 0) & (plus >= 0)
    nbr = {}
    for a, b in zip(minus[m], plus[m]):
        nbr.setdefault(a, []).append(b)
        nbr.setdefault(b, []).append(a)
    for c in slaves:
        assert any(agg[x] == agg[c - 1] for x in nbr[c - 1])


def test_cell_aggregation_reference_case():
    # test/AgFEMTests/CellAggregationTests.jl:9-43
    model = CartesianDiscreteModel((0, 1, 0, 1), (30, 30))
    geo = disk(0.4, x0=(0.5, 0.5))
    oc, fc = setup(model, geo)
    p = compile_postfix(geo.get_tree(), oc.oid_to_ls)
    agg, ok = ob.oracle_aggregate(oc, fc, p, IN, 1.0)
    assert ok
    check_invariants(model, oc, fc, p, IN, 1.0, agg)
    agg_out, ok = ob.oracle_aggregate(oc, fc, p, OUT, 1.0)
    assert ok
    check_invariants(model, oc, fc, p, OUT, 1.0, agg_out)
    assert not np.array_equal(agg, agg_out)
    # AggregateAllCutCells: no cut cell is its own root
    st = oc.bgcell_to_inoutcut(p)
    cells = np.arange(1, model.num_cells + 1)
    assert np.all(agg[st == CUT] != cells[st == CUT])


@pytest.mark.parametrize("thr", [0.25, 0.5, 1.0])
def test_cell_aggregation_thresholds_3d(thr):
    model = CartesianDiscreteModel((-1, 1, -1, 1, -1, 1), (9, 8, 7))
    geo = sphere(0.7)
    oc, fc = setup(model, geo)
    p = compile_postfix(geo.get_tree(), oc.oid_to_ls)
    agg, ok = ob.oracle_aggregate(oc, fc, p, IN, thr)
    assert ok
    check_invariants(model, oc, fc, p, IN, thr, agg)


def test_cell_aggregation_csg_named_geometry():
    model = CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (10, 10, 10))
    geo = csg_geometry()
    oc, fc = setup(model, geo)
    p = compile_postfix(get_geometry(geo, "csg").get_tree(), oc.oid_to_ls)
    agg, ok = ob.oracle_aggregate(oc, fc, p, IN, 1.0)
    assert ok
    check_invariants(model, oc, fc, p, IN, 1.0, agg)


def test_ghost_skeleton_mask():
    # EmbeddedDiscretizations.jl:504-543: interior facets next to a CUT cell whose two cells are active
    model = CartesianDiscreteModel((0, 1, 0, 1), (30, 30))
    geo = disk(0.4, x0=(0.5, 0.5))
    oc, fc = setup(model, geo)
    p = compile_postfix(geo.get_tree(), oc.oid_to_ls)
    st = oc.bgcell_to_inoutcut(p)
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    d, pos, minus, plus = FH.facet_cells(model, fn)
    for side in (IN, OUT, CUT):
        mask = ob.oracle_ghost_skeleton(oc, fc, p, side)
        interior = (minus >= 0) & (plus >= 0)
        assert not mask[~interior].any()
        a, b = st[minus[interior]], st[plus[interior]]
        act = lambda s: (s == CUT) | (s == side)
        expect = ((a == CUT) | (b == CUT)) & act(a) & act(b)
        assert np.array_equal(mask[interior], expect)
        assert mask.sum() > 0
    # ACTIVE_IN ghosts contain the CUT-only ghosts
    assert np.all(ob.oracle_ghost_skeleton(oc, fc, p, IN)[ob.oracle_ghost_skeleton(oc, fc, p, CUT)])


# ---- simplexified models: cells = the TRI / TET of simplexify(model), topology through the facet grid of the simplices ----
def _simplex_adjacency(model, fc):
    D = model.D
    cells = model.cell_node_ids() - 1
    fn = fc.array("facet_nodes").reshape(-1, fc.nvf)
    facet_of = {tuple(sorted(r)): f for f, r in enumerate(fn.tolist())}
    lfacets = [(0, 1), (0, 2), (1, 2)] if D == 2 else [(0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)]
    f2c = [[] for _ in range(fc.Nf)]
    for c, nodes in enumerate(cells.tolist()):
        for lf in lfacets:
            f2c[facet_of[tuple(sorted(nodes[i] for i in lf))]].append(c)
    return f2c


@pytest.mark.parametrize("case", ["disk", "sphere", "csg"])
@pytest.mark.parametrize("thr", [0.4, 1.0])
def test_cell_aggregation_simplexified(case, thr):
    from gecut import simplexify
    if case == "disk":
        model, geo, nm = simplexify(CartesianDiscreteModel((0, 1, 0, 1), (30, 30))), disk(0.4, x0=(0.5, 0.5)), None
    elif case == "sphere":
        model, geo, nm = simplexify(CartesianDiscreteModel((-1, 1, -1, 1, -1, 1), (9, 8, 7))), sphere(0.7), None
    else:
        model, geo, nm = simplexify(CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (8, 8, 8))), csg_geometry(), "csg"
    oc, fc = setup(model, geo)
    g = geo if nm is None else get_geometry(geo, nm)
    p = compile_postfix(g.get_tree(), oc.oid_to_ls)
    f2c = _simplex_adjacency(model, fc)
    assert all(len(c) in (1, 2) for c in f2c)
    assert sum(len(c) == 1 for c in f2c) == fc.Nboundary
    st = oc.bgcell_to_inoutcut(p)
    n = model.num_cells
    cells = np.arange(1, n + 1)
    nbr = {}
    for c in f2c:
        if len(c) == 2:
            nbr.setdefault(c[0], []).append(c[1])
            nbr.setdefault(c[1], []).append(c[0])
    for loc in (IN, OUT):
        agg, ok = ob.oracle_aggregate(oc, fc, p, loc, thr)
        assert ok
        assert np.all(agg[st == -loc] == 0)
        assert np.all(agg[st == loc] == cells[st == loc])
        assert np.all(agg[st == CUT] > 0)
        roots = np.unique(agg[agg > 0])
        assert np.all(agg[roots - 1] == roots)
        ids = oc.select_subcells(p, loc)
        vol = np.zeros(n)
        np.add.at(vol, oc.array("subcells.cell_to_bgcell")[ids - 1] - 1, oc.subcell_measures(ids))
        meas = np.array([oc.bgcell_measure(int(c)) for c in cells[st == CUT]])
        unit = np.zeros(n)
        unit[st == CUT] = vol[st == CUT] / meas
        cut_roots = roots[st[roots - 1] == CUT]
        assert np.all(unit[cut_roots - 1] >= thr - 1e-12)
        slaves = cells[(st == CUT) & (agg != cells)]
        assert np.all(unit[slaves - 1] < thr + 1e-12)
        for c in slaves:  # connected to the root through face neighbours with the same root
            assert any(agg[x] == agg[c - 1] for x in nbr[c - 1])
    # ghost skeleton: interior facets next to a CUT cell whose two cells are active
    for side in (IN, OUT, CUT):
        mask = ob.oracle_ghost_skeleton(oc, fc, p, side)
        act = lambda s: s == CUT or s == side
        expect = np.array([len(c) == 2 and (st[c[0]] == CUT or st[c[1]] == CUT) and act(st[c[0]]) and act(st[c[1]])
                           for c in f2c])
        assert np.array_equal(mask, expect)
