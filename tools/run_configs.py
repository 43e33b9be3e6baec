#!/usr/bin/env python
"""Runs the BASELINE.json configs other than the bench workload once on the GPU: timing, counts and the
size-independent checks (IN + OUT volume = box volume, known volumes), plus exact parity against the oracle where the
oracle finishes in seconds (config 1 at its full 40^3)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import gecut
from gecut import *
from gecut.measures import integrate_measure, integrate_laplacian
from test_gpu_parity import csg_geometry, stokes_geometry, bimaterial, compare_layout

def timed(f, n=3):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3, r

def report(name, model, geo, omega_names, check_oracle=False):
    def step(keep=False):
        cg = cut(model, geo)
        out = {}
        for nm in omega_names:
            om = Triangulation(cg, PHYSICAL_IN, nm); out[("vol", nm)] = integrate_measure(Measure(om, 2))
            ga = EmbeddedBoundary(cg, nm); out[("surf", nm)] = integrate_measure(Measure(ga, 2))
            cg.ctx.check(cg.ctx._lib.gecut_integrate_laplacian(om.handle, None, None), "lap")
            oo = Triangulation(cg, PHYSICAL_OUT, nm); out[("vol_out", nm)] = integrate_measure(Measure(oo, 2))
            om.close(); ga.close(); oo.close()
        out["sizes"] = cg.sizes
        if not keep:
            cg.close()
        return out, cg
    ms, _ = timed(step)
    ctx = gecut._lib.default_context()
    ctx.profile(True); step(); rep = ctx.profile_report(); ctx.profile(False)
    top = sorted(rep.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOPK", "4"))]
    out, cg = step(keep=True)
    s = out["sizes"]
    box = float(np.prod(np.asarray(model.pmax) - np.asarray(model.pmin)))
    line = {"config": name, "ms_per_step": round(ms, 3), "cells": model.num_cubes, "cells_per_s": model.num_cubes / ms * 1e3,
            "L": int(s.num_levelsets), "cut_cells": int(s.num_cutcells), "subcells": int(s.num_subcells), "subfacets": int(s.num_subfacets)}
    for nm in omega_names:
        v, vo, sf = out[("vol", nm)], out[("vol_out", nm)], out[("surf", nm)]
        line[f"{nm}:vol"] = v; line[f"{nm}:surf"] = sf
        line[f"{nm}:partition_err"] = abs(v + vo - box) / box
        assert abs(v + vo - box) <= 1e-11 * box, (nm, v, vo, box)
    if check_oracle:
        import oracle_binding as ob
        t = time.perf_counter(); oc = ob.oracle_cut(model, geo); line["oracle_cut_s"] = round(time.perf_counter() - t, 2)
        compare_layout(cg, oc); line["oracle_parity"] = "layout identical"
    line["top_kernels_ms"] = {k: round(v[1], 3) for k, v in top}
    print(json.dumps(line), flush=True)
    cg.close()
    return line

which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
if "c1" in which:
    n1 = int(os.environ.get("C1_N", "40"))
    report(f"C1 PoissonCSGCutFEM {n1}^3 L=10", CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (n1, n1, n1)), csg_geometry(), ["csg"], check_oracle=(n1 <= 48))
if "c2" in which:
    g = disk(0.7); b = get_metadata(g)
    l = report("C2 disk 4096^2", CartesianDiscreteModel(b.pmin, b.pmax, (4096, 4096)), g, [None])
    assert abs(l["None:vol"] - np.pi * 0.49) < 1e-6 and abs(l["None:surf"] - 2 * np.pi * 0.7) < 1e-6
if "c4" in which:
    geo, pmin, pmax = stokes_geometry()
    n = int(os.environ.get("C4_N", "256"))
    model = simplexify(CartesianDiscreteModel(pmin, pmax, (n, n, n)))
    dgeo = discretize(geo, model)   # DiscreteGeometry: nodal values on the device (NODAL leaves)
    report(f"C4 StokesTube DiscreteGeometry {n}^3 simplexified L=4", model, dgeo, ["fluid"])
if "c5" in which:
    n = int(os.environ.get("C5_N", "256"))
    report(f"C5 bimaterial {n}^3 L=2 (single-GPU size)", CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (n, n, n)), bimaterial(), ["steel", "concrete"])
