"""GPU parity of the AgFEM aggregation (`gecut_aggregate`) against the CPU oracle: `cell_to_cellin` must be identical
(Int32, bit-exact), through every call path of the reference's own test (`test/AgFEMTests/CellAggregationTests.jl`)."""
import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, disk, sphere, cut, cut_facets, aggregate, AggregateAllCutCells,
                   AggregateCutCellsByThreshold, get_geometry, simplexify, IN, OUT)
import oracle_binding as ob
from test_gpu_parity import csg_geometry, stokes_geometry, bimaterial, three_disks

pytestmark = pytest.mark.gpu

CASES = {
    "disk_30": lambda: (CartesianDiscreteModel((0, 1, 0, 1), (30, 30)), disk(0.4, x0=(0.5, 0.5)), None),
    "disk_corner": lambda: (CartesianDiscreteModel((0, 1, 0, 1), (17, 23)), disk(0.72, x0=(1, 1)), None),
    "three_disks": lambda: (CartesianDiscreteModel((-1, 2, -1, 2), (23, 19)), three_disks(), None),
    "sphere": lambda: (CartesianDiscreteModel((-1, 1, -1, 1, -1, 1), (9, 8, 7)), sphere(0.7), None),
    "csg": lambda: (CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (12, 12, 12)), csg_geometry(), "csg"),
    "bimaterial": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (18, 12, 18)), bimaterial(), "steel"),
    # too coarse for an IN root inside the cylinders: the reference's @assert all_aggregated fires
    "bimaterial_coarse": lambda: (CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (9, 6, 9)), bimaterial(), "steel"),
    # simplexified models (cells = the TRI / TET of simplexify(model), topology through the facet grid of the simplices)
    "sx_disk_30": lambda: (simplexify(CartesianDiscreteModel((0, 1, 0, 1), (30, 30))), disk(0.4, x0=(0.5, 0.5)), None),
    "sx_disk_corner": lambda: (simplexify(CartesianDiscreteModel((0, 1, 0, 1), (17, 23))), disk(0.72, x0=(1, 1)), None),
    "sx_three_disks": lambda: (simplexify(CartesianDiscreteModel((-1, 2, -1, 2), (23, 19))), three_disks(), None),
    "sx_sphere": lambda: (simplexify(CartesianDiscreteModel((-1, 1, -1, 1, -1, 1), (9, 8, 7))), sphere(0.7), None),
    "sx_csg": lambda: (simplexify(CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (10, 10, 10))), csg_geometry(), "csg"),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("thr", [1.0, 0.4])
def test_aggregate_parity(name, thr):
    model, geo, nm = CASES[name]()
    cg = cut(model, geo)
    fc = cut_facets(model, geo)
    oc = ob.oracle_cut(model, geo)
    ofc = ob.oracle_cut_facets(model, geo)
    prog = cg.program(nm)
    strategy = AggregateCutCellsByThreshold(thr)
    for loc in (IN, OUT):
        ref, ok = ob.oracle_aggregate(oc, ofc, prog, loc, thr)
        if not ok:  # the reference asserts; the library reports the same condition as an error
            with pytest.raises(gecut._lib.GecutError, match="not all cut cells"):
                aggregate(strategy, cg, fc, nm, loc)
            continue
        a1 = aggregate(strategy, cg, fc, nm, loc)          # with an EmbeddedFacetDiscretization
        a2 = aggregate(strategy, cg, nm, loc)               # compute_bgfacet_to_inoutcut(cut.bgmodel, geo) inside
        assert a1.dtype == np.int32 and np.array_equal(a1, ref)
        assert np.array_equal(a2, ref)
    if ob.oracle_aggregate(oc, ofc, prog, IN, thr)[1]:
        a3 = aggregate(strategy, cg, fc, nm, IN, device_output=True)
        assert np.array_equal(a3.cpu().numpy(), aggregate(strategy, cg, fc, nm, IN))
    cg.close()
    fc.close()


def test_aggregate_defaults_and_errors():
    model = CartesianDiscreteModel((0, 1, 0, 1), (30, 30))
    geo = disk(0.4, x0=(0.5, 0.5))
    cg = cut(model, geo)
    fc = cut_facets(model, geo)
    a = aggregate(AggregateAllCutCells(), cg)               # aggregate(strategy, cut) (CellAggregation.jl:1-3)
    assert np.array_equal(a, aggregate(AggregateAllCutCells(), cg, fc))
    assert np.array_equal(a, aggregate(AggregateAllCutCells(), cg, geo, IN))
    with pytest.raises(ValueError):
        AggregateCutCellsByThreshold(1.5)
    # simplexified models go through the same entry points (facet grid of the simplices)
    ms = simplexify(model)
    cs = cut(ms, geo)
    fs = cut_facets(ms, geo)
    a = aggregate(AggregateAllCutCells(), cs)
    assert a.shape == (ms.num_cells,) and np.array_equal(a, aggregate(AggregateAllCutCells(), cs, fs))
    with pytest.raises(gecut._lib.GecutError):  # a facet discretization of another model
        aggregate(AggregateAllCutCells(), cs, fc)
    for x in (cg, fc, cs, fs):
        x.close()


def test_aggregate_at_scale():
    """256^3 sphere: invariants without the oracle (roots are fixed points, every CUT cell has an IN root at most a few
    cells away, OUT cells are 0) and the time of the whole aggregation."""
    n = 256
    g = sphere(0.7)
    b = gecut.get_metadata(g)
    model = CartesianDiscreteModel(b.pmin, b.pmax, (n, n, n))
    cg = cut(model, g)
    fc = cut_facets(model, g)
    agg = aggregate(AggregateAllCutCells(), cg, fc)
    st = gecut.compute_bgcell_to_inoutcut(cg, g)
    cells = np.arange(1, model.num_cells + 1, dtype=np.int64)
    assert np.all(agg[st == OUT] == 0)
    assert np.all(agg[st == IN] == cells[st == IN])
    cutm = st == 0
    assert np.all(agg[cutm] > 0) and np.all(st[agg[cutm] - 1] == IN)
    # distance (in cells) between a cut cell and its root
    def ijk(c):
        c = c - 1
        return np.stack([c % n, (c // n) % n, c // (n * n)], 1)
    dist = np.abs(ijk(cells[cutm]) - ijk(agg[cutm].astype(np.int64))).max(1)
    assert dist.max() <= 3
    cg.close()
    fc.close()


@pytest.mark.parametrize("name", ["disk_30", "disk_corner", "three_disks", "sphere", "csg", "bimaterial",
                                  "sx_disk_30", "sx_disk_corner", "sx_three_disks", "sx_sphere", "sx_csg"])
def test_ghost_skeleton_parity(name):
    from gecut import GhostSkeleton, ACTIVE_IN, ACTIVE_OUT, PHYSICAL_IN
    model, geo, nm = CASES[name]()
    cg = cut(model, geo)
    oc = ob.oracle_cut(model, geo)
    ofc = ob.oracle_cut_facets(model, geo)
    prog = cg.program(nm)
    for sel, side in ((ACTIVE_IN, IN), (ACTIVE_OUT, OUT), (0, 0)):
        ref = ob.oracle_ghost_skeleton(oc, ofc, prog, side)
        mask = GhostSkeleton(cg, sel, nm)
        assert mask.dtype == bool and np.array_equal(mask, ref)
        ids = GhostSkeleton(cg, sel, nm, ids=True)
        assert np.array_equal(ids, np.nonzero(ref)[0] + 1)
    assert np.array_equal(GhostSkeleton(cg) if nm is None else GhostSkeleton(cg, nm), ob.oracle_ghost_skeleton(oc, ofc, prog, IN))
    with pytest.raises(NotImplementedError):
        GhostSkeleton(cg, PHYSICAL_IN, nm)
    cg.close()


@pytest.mark.parametrize("name,nslab", [("disk_30", 3), ("three_disks", 2), ("sphere", 3), ("csg", 2), ("bimaterial", 4),
                                        ("sx_three_disks", 3), ("sx_sphere", 2), ("sx_csg", 3)])
@pytest.mark.parametrize("threshold", [1.0, 0.4])
def test_aggregation_over_slabs_equals_the_whole_model(name, nslab, threshold):
    """Distributed aggregation (`src/Distributed/DistributedAgFEM.jl:52-144`) on z-slabs: every slab sweeps its own cut cells
    with the ghost cell layers refreshed after every sweep; the global roots equal the whole-model aggregation (and hence the
    oracle's).  The slabs live on one GPU here (ghost layers by device copies); `gecut_comm_aggregate` runs the same sweeps
    with NCCL transport (tests/test_gpu_multigpu.py)."""
    from gecut import aggregate_slabs, LevelSetCutter, _lib
    model, geo, nm = CASES[name]()
    ctx = _lib.default_context()
    strategy = AggregateCutCellsByThreshold(threshold)
    nz = model.partition[-1]
    cutsz = [round(q * nz / nslab) for q in range(nslab + 1)]
    for side in (IN, OUT):
        whole_cut = cut(model, geo)
        try:
            want = aggregate(strategy, whole_cut, nm, side)
        except _lib.GecutError:
            whole_cut.close()
            continue  # the reference's @assert all_aggregated fires for this case: nothing to compare
        cuts = [LevelSetCutter(ctx, slab=(a, b)).cut(model, geo) for a, b in zip(cutsz[:-1], cutsz[1:])]
        got = aggregate_slabs(strategy, cuts, None, nm, side)
        assert np.array_equal(np.concatenate(got), want)
        # with user-supplied slab facet discretizations (full cut_facets, not only the classification)
        facets = [LevelSetCutter(ctx, slab=(a, b)).cut_facets(model, geo) for a, b in zip(cutsz[:-1], cutsz[1:])]
        got2 = aggregate_slabs(strategy, cuts, facets, nm, side)
        assert np.array_equal(np.concatenate(got2), want)
        for o in cuts + facets + [whole_cut]:
            o.close()


@pytest.mark.parametrize("name,nslab", [("disk_30", 3), ("three_disks", 2), ("sphere", 3), ("csg", 2), ("bimaterial", 4),
                                        ("sx_three_disks", 3), ("sx_sphere", 2), ("sx_csg", 3)])
def test_ghost_skeleton_over_slabs_equals_the_whole_model(name, nslab):
    """GhostSkeleton on z-slabs: the mask of every slab's local model equals the whole-model mask on the slab's facets
    (through the facet -> node-tuple map), including the facets on the planes two slabs share."""
    from gecut import GhostSkeleton, GhostSkeleton_slabs, LevelSetCutter, _lib, ACTIVE_IN, ACTIVE_OUT
    model, geo, nm = CASES[name]()
    ctx = _lib.default_context()
    whole_cut = cut(model, geo)
    wf = cut_facets(model, geo)
    hw = wf.to_host(topology=True)
    gnodes = {tuple(r): f for f, r in enumerate(hw["facet_to_nodes"])}
    nz = model.partition[-1]
    layer_nodes = model.num_nodes // (nz + 1)
    cutsz = [round(q * nz / nslab) for q in range(nslab + 1)]
    slabs = list(zip(cutsz[:-1], cutsz[1:]))
    cuts = [LevelSetCutter(ctx, slab=s).cut(model, geo) for s in slabs]
    l2g = []
    for s in slabs:
        lf = LevelSetCutter(ctx, slab=s).cut_facets(model, geo)
        hl = lf.to_host(topology=True)
        l2g.append(np.array([gnodes[tuple(int(v) + s[0] * layer_nodes for v in r)] for r in hl["facet_to_nodes"]]))
        lf.close()
    for sel in (ACTIVE_IN, ACTIVE_OUT):
        want = GhostSkeleton(whole_cut, sel, nm)
        got = GhostSkeleton_slabs(cuts, sel, nm)
        seen = np.zeros(len(want), bool)
        for m, g in zip(got, l2g):
            assert np.array_equal(m, want[g])
            seen[g[m]] = True
        assert np.array_equal(seen, want)  # every ghost facet of the model is marked in some slab
    for o in cuts + [whole_cut, wf]:
        o.close()
