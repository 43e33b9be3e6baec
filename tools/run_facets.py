#!/usr/bin/env python
"""Times `cut_facets` + the physical boundary / skeleton measures on the GPU on the BASELINE
geometries: per-step time, per-kernel CUDA-event times and the algorithmic bytes of the classification pass."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import gecut
from gecut import *
from test_gpu_parity import csg_geometry, bimaterial

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)


def report(name, model, geo, nm=None, steps=5):
    ctx = gecut._lib.default_context()

    def step(select=True):
        fc = cut_facets(model, geo)
        out = {"sizes": fc.sizes}
        if select:
            for ctor, key in ((BoundaryTriangulation, "boundary"), (SkeletonTriangulation, "skeleton")):
                t = ctor(fc, PHYSICAL_IN, nm)
                out[key] = t.measure()
                t.close()
        fc.close()
        return out

    def timed(f):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            r = f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, r

    ms_cut, out = timed(lambda: step(False))
    ms_all, out = timed(step)
    ctx.profile(True); step(); rep = ctx.profile_report(); ctx.profile(False)
    s = out["sizes"]
    L = int(s.num_levelsets)
    D = int(s.D)
    cls_bytes = L * s.num_bgfacets + s.num_bgfacets / 8 + 4 * s.num_cutfacets
    lay_bytes = s.num_subfacets * (4 * D + 4 + L + 8) + s.num_subfacet_points * 8 * (2 * D - 1)
    line = {"config": name, "facets": int(s.num_bgfacets), "cut_facets": int(s.num_cutfacets), "subfacets": int(s.num_subfacets),
            "L": L, "ms_cut_facets": round(ms_cut, 3), "ms_with_selectors": round(ms_all, 3),
            "facets_per_s": s.num_bgfacets / ms_cut * 1e3,
            "classify_alg_bytes": int(cls_bytes), "layout_bytes": int(lay_bytes),
            "boundary_in": out["boundary"], "skeleton_in": out["skeleton"],
            "kernels_ms": {k: round(v[1], 3) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]}}
    kc = rep.get("k_facet_classify")
    if kc:
        line["classify_gbs"] = round(cls_bytes / (kc[1] / kc[0] * 1e-3) / 1e9, 1)
        line["classify_frac_of_peak"] = round(line["classify_gbs"] / PEAK, 3)
    print(json.dumps(line), flush=True)


which = sys.argv[1:] or ["sphere", "disk", "csg", "bimat"]
n3 = int(os.environ.get("N3", "512"))
if "sphere" in which:
    g = sphere(0.7); b = get_metadata(g)
    report(f"sphere {n3}^3 L=1", CartesianDiscreteModel(b.pmin, b.pmax, (n3,) * 3), g)
if "disk" in which:
    g = disk(0.7); b = get_metadata(g)
    report("disk 4096^2 L=1", CartesianDiscreteModel(b.pmin, b.pmax, (4096, 4096)), g)
if "csg" in which:
    n1 = int(os.environ.get("C1_N", "256"))
    report(f"CSG {n1}^3 L=10", CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, (n1,) * 3), csg_geometry(), "csg")
if "bimat" in which:
    n5 = int(os.environ.get("C5_N", "512"))
    report(f"bimaterial {n5}^3 L=2", CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (n5,) * 3), bimaterial(), "steel")
