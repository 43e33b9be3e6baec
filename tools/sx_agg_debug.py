"""Debug helper (GPU): simplexified aggregation / ghost skeleton against the oracle, printing the first mismatches."""
import sys
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import gecut
from gecut import (CartesianDiscreteModel, simplexify, disk, sphere, cut, cut_facets, aggregate, AggregateCutCellsByThreshold,
                   GhostSkeleton, ACTIVE_IN, IN, OUT)
import oracle_binding as ob

for model, geo in [(simplexify(CartesianDiscreteModel((0, 1, 0, 1), (30, 30))), disk(0.4, x0=(0.5, 0.5))),
                   (simplexify(CartesianDiscreteModel((-1, 1, -1, 1, -1, 1), (4, 3, 3))), sphere(0.7))]:
    cg, fc = cut(model, geo), cut_facets(model, geo)
    oc, ofc = ob.oracle_cut(model, geo), ob.oracle_cut_facets(model, geo)
    prog = cg.program(None)
    nt = 2 if model.D == 2 else 6
    st = oc.rows("ls_to_bgcell_to_inoutcut")[0] if hasattr(oc, "rows") else None
    for thr in (1.0, 0.4):
        for loc in (IN, OUT):
            ref, ok = ob.oracle_aggregate(oc, ofc, prog, loc, thr)
            got = aggregate(AggregateCutCellsByThreshold(thr), cg, fc, None, loc)
            bad = np.nonzero(got != ref)[0]
            print("D", model.D, "thr", thr, "loc", loc, "ok", ok, "mismatches", len(bad), "of", len(ref))
            for c in bad[:12]:
                print("   cell", c, "cube", c // nt, "t", c % nt, "got", got[c], "ref", ref[c])
    for side, sel in ((IN, ACTIVE_IN),):
        refm = ob.oracle_ghost_skeleton(oc, ofc, prog, side)
        m = GhostSkeleton(cg, sel, None)
        bad = np.nonzero(m != refm)[0]
        print("ghost mismatches", len(bad), "of", len(refm), "marked", int(m.sum()), int(refm.sum()))
        fn = ofc.array("facet_nodes").reshape(-1, model.D)
        for f in bad[:12]:
            print("   facet", f, fn[f], "got", m[f], "ref", refm[f])
