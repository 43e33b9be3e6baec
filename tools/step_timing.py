#!/usr/bin/env python
"""Wall-clock per API call of one bench step (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gecut
from gecut import *
from gecut import _lib
from gecut.measures import integrate_measure
sys.argv += [""] * 3
wl = sys.argv[1] or "c3_tet"
n = int(sys.argv[2] or 512)
import bench
model, geo, _ = bench.make_workload(wl, n, 1)
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
cutter = LevelSetCutter(ctx)
def T(f, name, acc):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + (time.perf_counter() - t) * 1e3; return r
for it in range(6):
    acc = {}
    cg = T(lambda: cutter.cut(model, geo), "cut", acc)
    om = T(lambda: Triangulation(cg, PHYSICAL), "select_physical", acc)
    ga = T(lambda: EmbeddedBoundary(cg), "select_boundary", acc)
    T(lambda: integrate_measure(Measure(om, 2)), "measure_vol", acc)
    T(lambda: integrate_measure(Measure(ga, 2)), "measure_surf", acc)
    T(lambda: ctx.check(ctx._lib.gecut_integrate_laplacian(om.handle, None, None), "lap"), "laplacian", acc)
    T(lambda: (om.close(), ga.close(), cg.close()), "free", acc)
    print(it, " ".join(f"{k}={v:.2f}" for k, v in acc.items()), "total=%.2f" % sum(acc.values()), flush=True)
