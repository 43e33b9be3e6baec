#!/usr/bin/env python
"""Single-GPU replay of what rank 0 does in bench.py's multi-GPU check: its own slab several times, then the other slabs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gecut
from gecut import *
from gecut import _lib
from gecut.measures import integrate_measure
import bench
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
wl = sys.argv[3] if len(sys.argv) > 3 else "c3_tet"
model, geo, _ = bench.make_workload(wl, n, world)
from gecut.distributed import balanced_slab_ranges
ranges = balanced_slab_ranges(model, geo, world, cut_weight=300.0)
print("ranges", ranges, flush=True)
ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
def one(slab, lap=True):
    t = time.perf_counter()
    cg = LevelSetCutter(ctx, slab=slab).cut(model, geo)
    om = Triangulation(cg, gecut.PHYSICAL_IN)
    if lap:
        ctx.check(ctx._lib.gecut_integrate_laplacian(om.handle, None, None), "lap")
    ga = EmbeddedBoundary(cg)
    v, s = integrate_measure(Measure(om, 2)), integrate_measure(Measure(ga, 2))
    out = (v, s, cg.num_cutcells, cg.sizes.num_subcells)
    om.close(); ga.close(); cg.close()
    torch.cuda.synchronize()
    print("slab", slab, "lap", lap, out, "%.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    return out
for _ in range(3):
    one(tuple(ranges[0]))
for sl in ranges:
    one(tuple(sl), lap=False)
print("done", flush=True)
