#!/usr/bin/env python
"""bench.py -- background cells/s through cut() + ∫dΩ on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic background model:
    cutgeo = cut(LevelSetCutter(), bgmodel, geo)             (classification + subtriangulation, all outputs in HBM)
    Ω = Triangulation(cutgeo, PHYSICAL); Γ = EmbeddedBoundary(cutgeo)
    sum(∫1 dΩ), sum(∫1 dΓ), cut-cell Laplacian element matrices of ∫∇v·∇u dΩ
Workload (config.workload): BASELINE.json configs[2] -- sphere R=0.7 on the n^3 Cartesian box ±0.707 (n=512 per GPU),
simplexified (6 TET per HEX) by default, z-slab sharded over the ranks with the global model growing with N (weak
scaling: N=8 -> 1024^3).  `value` counts n^3 n-cube cells per second (NOT the 6x larger number of simplices).

Lines printed: rank 0 prints ONE JSON line (see the keys required by the driver contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bg cells/s through cut()+∫dΩ at n³"
UNIT = "cells/s"


# --------------------------------------------------------------------------------------------------------------------
def make_workload(name, n, world):
    import gecut
    from gecut import CartesianDiscreteModel, simplexify, sphere, disk
    if name in ("c3_tet", "c3_hex"):
        geo = sphere(0.7)
        box = gecut.get_metadata(geo)
        mult = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
        if mult is None:
            raise SystemExit("--gpus must be 1, 2, 4 or 8")
        part = tuple(n * m for m in mult)
        model = CartesianDiscreteModel(box.pmin, box.pmax, part)
        if name == "c3_tet":
            model = simplexify(model)
        desc = (f"BASELINE configs[2]: sphere(R=0.7) on {part[0]}x{part[1]}x{part[2]} Cartesian box ±0.707"
                f"{', simplexified (6 TET/HEX)' if name == 'c3_tet' else ' (HEX, Q1)'}; cut + PHYSICAL volume + "
                f"EmbeddedBoundary area + cut-cell {'P1 4x4' if name == 'c3_tet' else 'Q1 8x8'} Laplacian matrices")
        return model, geo, desc
    if name == "c2_disk":
        geo = disk(0.7)
        box = gecut.get_metadata(geo)
        mult = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}[world]
        part = tuple(n * m for m in mult)
        model = CartesianDiscreteModel(box.pmin, box.pmax, part)
        return model, geo, f"BASELINE configs[1]: disk(R=0.7) on {part[0]}x{part[1]} quads: cut + area + perimeter + Q1 4x4 matrices"
    if name == "c1_csg":
        # PoissonCSGCutFEM.jl:9-23 (BASELINE configs[0], the north-star "512^3 3-D CSG cut"): box ∩ sphere − 3 cylinders,
        # ten unique level sets (6 planes of the cube, the sphere, 3 cylinders) on [-0.8, 0.8]^3
        from gecut import cylinder, cube, union, intersect, setdiff
        R = 0.5
        src = union(union(cylinder(R, v=(1, 0, 0)), cylinder(R, v=(0, 1, 0))), cylinder(R, v=(0, 0, 1)), name="source")
        geo = setdiff(intersect(cube(L=1.5), sphere(1)), src, name="csg")
        mult = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
        if mult is None:
            raise SystemExit("--gpus must be 1, 2, 4 or 8")
        part = tuple(n * m for m in mult)
        model = CartesianDiscreteModel((-0.8,) * 3, (0.8,) * 3, part)
        desc = (f"BASELINE configs[0] at the north-star size: PoissonCSGCutFEM geometry (box ∩ sphere − 3 cylinders, L=10) on "
                f"{part[0]}x{part[1]}x{part[2]} HEX box [-0.8,0.8]^3; cut + PHYSICAL \"csg\" volume + EmbeddedBoundary(\"csg\") "
                f"area + cut-cell Q1 8x8 Laplacian matrices")
        return model, geo, desc
    if name == "c5_bimaterial":
        # BimaterialLinElastCutFEM.jl:16-37,56: two steel cylinders (axis x) in a concrete block, tree union(steel, !steel)
        from gecut import cylinder, union, complement
        geo1 = cylinder(0.4, x0=(0.0, 1.0, 0.7))
        geo2 = cylinder(0.4, x0=(0.0, 1.0, 2.3))
        steel = union(geo1, geo2, name="steel")
        geo = union(steel, complement(steel, name="concrete"))
        mult = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world)
        if mult is None:
            raise SystemExit("--gpus must be 1, 2, 4 or 8")
        part = tuple(n * m for m in mult)
        model = CartesianDiscreteModel((0.0, 0.0, 0.0), (3.0, 2.0, 3.0), part)
        desc = (f"BASELINE configs[4]: BimaterialLinElastCutFEM two-phase interface (2 cylinders, L=2) on "
                f"{part[0]}x{part[1]}x{part[2]} HEX box [0,3]x[0,2]x[0,3]; cut + steel (IN) and concrete (OUT) PHYSICAL "
                f"volumes + interface area + cut-cell Q1 8x8 Laplacian matrices of both phases")
        return model, geo, desc
    raise SystemExit(f"unknown workload {name}")


def workload_plan(name):
    """(cell selections [(sub-geometry name, side)], boundary names) integrated by one step; side -1 = IN, +1 = OUT."""
    if name == "c5_bimaterial":
        return [("steel", -1), ("steel", 1)], ["steel"]
    if name == "c1_csg":
        return [("csg", -1)], ["csg"]
    return [(None, -1)], [None]


def host_program(model, geo, name):
    """postfix program of the (named sub-)geometry over the unique leaves, compiled on the host (CPU arms)"""
    import gecut
    from gecut.csg import compile_postfix, get_geometry
    _, _, oid_to_ls = gecut.cutters.flatten_geometry(model, geo, host_only=True)
    g = geo if name is None else get_geometry(geo, name)
    return [int(x) for x in compile_postfix(g.get_tree(), oid_to_ls)]


def oracle_step(oc, model, geo, workload):
    """one step of the workload on an oracle cut: (volumes per cell selection, areas per boundary, matrices)"""
    cells, bnds = workload_plan(workload)
    vols, surfs, nK = [], [], 0
    for name, side in cells:
        prog = host_program(model, geo, name)
        ids = oc.select_subcells(prog, side)
        bg = oc.select_bgcells(prog, side)
        vols.append(float(oc.subcell_measures(ids).sum() + len(bg) * oc.bgcell_measure(1)))
        nK += int(oc.cutcell_laplacian(ids).shape[0])
    for name in bnds:
        fids, _ = oc.select_boundary(host_program(model, geo, name))
        surfs.append(float(oc.subfacet_measures(fids).sum()))
    return vols, surfs, nK


def slab_of(model, rank, world):
    nl = model.partition[-1]
    assert nl % world == 0
    per = nl // world
    return (rank * per, (rank + 1) * per)


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region (B200_PROFILING.md recipe), read through NVML in-process
    (pynvml / nvidia-ml-py) every 5 ms: the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints, without the heavy per-sample cost of the nvidia-smi loop (which was measured to stretch 5.5 ms steps to
    10-16 ms).  Falls back to one `nvidia-smi` query after the timed region if NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.smax = None
        self.stop_flag = False
        self.period = 0.02  # seconds; 5 ms during the short device-resident timed region (NVML calls every 5 ms slowed the
        # pinned D2H copies of the end-to-end leg by 20 %: 76.6 -> 93.4 ms)
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
        except Exception:
            self.nvml = None

    def _loop(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def clear(self):
        self.samples.clear()
        self.reasons.clear()

    def stop(self):
        if self.nvml is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=1)
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": "NVML in-process, 5 ms period in the timed region, 20 ms after"}


# --------------------------------------------------------------------------------------------------------------------
def layout_bytes(s):
    """bytes of every output array of the layout (B_cut + states)"""
    D, L = int(s.D), int(s.num_levelsets)
    b = {}
    b["bg_state"] = L * s.num_bgcells
    b["cutcells"] = 4 * s.num_cutcells
    b["subcells"] = s.num_subcells * (4 * (D + 1) + 4 + L) + s.num_subcell_points * 16 * D
    # one level set: the subfacet point arrays alias the subcell point arrays on the device (the reference aliases them
    # too until merge_sub_face_data copies them), so they are not moved a second time
    sfp = 0 if L == 1 else s.num_subfacet_points
    b["subfacets"] = s.num_subfacets * (4 * D + 4 + 8 * D + L) + sfp * 16 * D
    return {k: int(v) for k, v in b.items()}


def algorithmic_bytes(s, model, nsel, nfsel, nodal=False, ncellsel=1):
    """ALGORITHMIC bytes per kernel and per step (DESIGN.md section 'Roofline accounting')."""
    D, L = int(s.D), int(s.num_levelsets)
    nv = int(s.num_vertices_per_bgcell)
    Nc, Nn, Ncut = int(s.num_bgcells), int(s.num_bgnodes), int(s.num_cutcells)
    lb = layout_bytes(s)
    nt0 = 1 if model.simplexified else (2 if D == 2 else 6)
    nsimp = Ncut * nt0
    narr = 2 + L + (L if L > 1 else 0)
    k = {}
    # NODAL leaves (DiscreteGeometry): k_nodal_signs streams the values once (8 B per node in, 1 bit out), the marching
    # kernel then reads the bits
    k["k_classify"] = L * Nc + Nc // 8 + (L * Nn // 8 if nodal else 0)
    k["k_nodal_signs"] = (8 * L * Nn + L * Nn // 8) if nodal else 0
    ntile32 = (nsimp + 31) // 32
    # L > 1: one info word per stage-0 simplex and narr totals per 32-simplex tile (no per-simplex counters)
    k["k_cutm_count"] = 4 * Ncut + 4 * nsimp + 4 * narr * ntile32
    k["k_cutm_fill"] = 4 * Ncut + 4 * nsimp + 4 * narr * ntile32 + lb["subcells"] + lb["subfacets"] + 8 * Ncut
    ntiles = (nsimp + 127) // 128
    k["k_cut1_count"] = 4 * Ncut + 12 * ntiles
    meas = 8 * (int(s.num_subcells) + int(s.num_subfacets))  # per-entity measures written next to the layout
    if model.simplexified and L == 1:
        meas = 0  # simplex cells, one level set: summed on chip (tile statistics, per-cell IN / OUT), per entity on demand
    k["k_cut1_fill"] = 4 * Ncut + 12 * ntiles + lb["subcells"] + lb["subfacets"] + 4 * Ncut + meas
    if L == 1:  # tile statistics (4 doubles per 32-simplex tile) and, for simplex bg cells, (IN, OUT) measure per cut cell
        k["k_cut1_fill"] += 32 * ((nsimp + 31) // 32) + (16 * Ncut if model.simplexified else 0)
    # L > 1: k_cutm_fill writes the layout of the simplices cut by at most one level set (the bulk), k_cut_walk the rest;
    # the whole layout is attributed to the tile kernel (the walk's share is not subtracted: an upper bound of its GB/s)
    k["k_cutm_fill"] += meas
    k["k_cut_two_fill"] = None
    k["k_count_bgcells"] = L * Nc * ncellsel
    k["k_select_subcells"] = int(s.num_subcells) * L * 2 * ncellsel + 4 * nsel
    k["k_flag_subcells"] = int(s.num_subcells) * (4 + 2 * L + 4)
    k["k_flag_boundary"] = int(s.num_subfacets) * (L + 4 + 1)
    k["k_measures"] = (nsel + nfsel) * (4 + 8 + 8)  # id + gathered measure + value
    nsc = int(s.num_subcells)
    if model.simplexified and L == 1:  # P1, one level set: (IN, OUT) measure + id of every cut cell in, matrices out
        k["k_cutcell_laplacian"] = (16 + 4) * Ncut + 8 * nv * nv * Ncut
    elif model.simplexified:  # P1: states + measures of the cut cells' subcells, matrices out
        k["k_cutcell_laplacian"] = ncellsel * (nsc * L + 8 * nv * nv * Ncut + 8 * Ncut) + nsel * 8
    else:  # Q1: + connectivity and reference coordinates of the selected subcells
        k["k_cutcell_laplacian"] = (ncellsel * (nsc * L + 8 * nv * nv * Ncut + 8 * Ncut) +
                                    nsel * (8 + 4 * (D + 1) + 8 * D * (D + 1)))
    # whole step, SURVEY.md section 8(d): B_sweep + B_cut + B_int (node values are never materialised: 16*L*Nn of the
    # survey's B_sweep is not moved at all, so it is NOT counted here)
    step = (L * Nc + Nc // 8) + (8 * L * Nn if nodal else 0) + lb["cutcells"] + lb["subcells"] + lb["subfacets"] + meas + 8 * (nsel + nfsel) + 8 * nv * nv * Ncut * ncellsel
    return k, int(step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3_tet", choices=["c3_tet", "c3_hex", "c2_disk", "c5_bimaterial", "c1_csg"])
    ap.add_argument("--n", type=int, default=None, help="cells per direction per GPU (default 512; 4096 for c2_disk)")
    ap.add_argument("--cpu-n", type=int, default=None, help="size of the bounded CPU sample (default: the GPU arm's n)")
    ap.add_argument("--cpu-stride", type=int, default=None,
                    help="reference arm: a step covers every m-th thin z-slab of the model (default: chosen so that a step "
                         "takes a few seconds; 1 = the whole model)")
    ap.add_argument("--nodal", action="store_true",
                    help="feed the level sets as a DiscreteGeometry: slab-local nodal vectors resident on each GPU, ghost "
                         "node plane from the next rank by NCCL send/recv every step (SURVEY 8e)")
    ap.add_argument("--uniform-slabs", action="store_true", help="even z-slabs instead of the work-balanced partition")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-csg", action="store_true", help="skip the north-star CSG leg of the default single-GPU run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    n = args.n or (4096 if args.workload == "c2_disk" else 512)
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, n)
        return
    run_b200(args, n, rank, world, local_rank)


# --------------------------------------------------------------------------------------------------------------------
def cpu_sample(workload, n_cpu, steps=1):
    """The CPU restatement of the reference (oracle/, 1 thread = the reference's execution model) on a bounded sample of
    the same workload.  Returns (cells/s, description)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_binding as ob
    model, geo, _ = make_workload(workload, n_cpu, 1)
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        oc = ob.oracle_cut(model, geo)
        oracle_step(oc, model, geo, workload)
        dt = time.perf_counter() - t0
        del oc
        best = dt if best is None else min(best, dt)
    cubes = model.num_cubes
    return cubes / best, best, (f"same geometry at {'x'.join(str(p) for p in model.partition)} "
                                f"({'simplexified' if model.simplexified else 'n-cubes'}), full cut + selectors + integrals, "
                                f"oracle/gecut_oracle.cpp -O2 single thread (the reference is serial Julia; it cannot run here)")


def _cpu_slab_job(job):
    """One thin z-slab of the model through the oracle (a worker process = one rank of the reference's distributed mode,
    src/Distributed/DistributedDiscretizations.jl:29-40: every rank cuts its own part of the Cartesian model)."""
    workload, n, world, k0, k1 = job
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, ROOT)
    import oracle_binding as ob
    import gecut
    model, geo, _ = make_workload(workload, n, world)
    D = model.D
    dz = model.sizes[D - 1]
    pmin, pmax, part = list(model.pmin), list(model.pmax), list(model.partition)
    pmax[D - 1] = pmin[D - 1] + k1 * dz
    pmin[D - 1] = pmin[D - 1] + k0 * dz
    part[D - 1] = k1 - k0
    sub = gecut.CartesianDiscreteModel(tuple(pmin), tuple(pmax), tuple(part), simplexified=model.simplexified)
    oc = ob.oracle_cut(sub, geo)
    vols, surfs, nK = oracle_step(oc, sub, geo, workload)
    return float(vols[0]), float(surfs[0]), nK


def run_reference(args, n):
    """Reference arm: the CPU restatement on ALL host cores, on the SAME model as the GPU arm (same n, same geometry, same
    --gpus multiplier): thin z-slabs of the model are dealt to one worker process per core, the way the reference itself
    uses many cores (one rank per part, src/Distributed; the library has no threading).  When a whole pass would take more
    than a few seconds (N > 1: up to 1024^3 cells) a step covers every m-th thin slab -- a bounded sample of the same
    model at the same resolution, spread evenly over the slab direction -- and the value is sample cells / time."""
    import concurrent.futures as cf
    cores = max(1, min(os.cpu_count() or 1, 64))
    n_cpu = args.cpu_n or n
    model, geo, desc = make_workload(args.workload, n_cpu, args.gpus)
    nz = model.partition[model.D - 1]
    layer_cubes = model.num_cubes // nz
    chunk = max(1, nz // (cores * 4))
    chunks = [(k, min(k + chunk, nz)) for k in range(0, nz, chunk)]
    # predicted seconds of a whole pass at ~6e7 cells/s per 16 cores (measured round 1: 4.1e7 at 192^3, growing with n)
    pred = model.num_cubes / (6.0e7 * cores / 16.0)
    m = args.cpu_stride or max(1, int(-(-pred // 4.0)))
    picked = chunks[::m]
    cores = min(cores, len(picked))
    jobs = [(args.workload, n_cpu, args.gpus, a, b) for a, b in picked]
    sample_cubes = sum(b - a for a, b in picked) * layer_cubes
    with cf.ProcessPoolExecutor(max_workers=cores) as pool:
        def one_step():
            res = list(pool.map(_cpu_slab_job, jobs))
            return sum(r[0] for r in res), sum(r[1] for r in res)
        for _ in range(max(args.warmup, 1)):
            one_step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            vol, surf = one_step()
        elapsed = time.perf_counter() - t0
    value = sample_cubes * args.steps / elapsed
    whole = m == 1
    sample = (("the whole model" if whole else f"every {m}-th thin z-slab ({len(picked)} of {len(chunks)} slabs of {chunk} "
               f"layers = {sample_cubes / model.num_cubes:.3f} of the cells)") +
              f" of the same model as the GPU arm ({'x'.join(str(p) for p in model.partition)}, "
              f"{'simplexified' if model.simplexified else 'n-cubes'}), full cut + selectors + integrals per step, "
              f"oracle/gecut_oracle.cpp -O2 on {cores} worker processes (one thin z-slab at a time each, as the reference's "
              f"distributed mode would use the cores; the reference is Julia and cannot run here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "same_size_as_gpu_arm": True, "partition": list(model.partition),
                   "sample_fraction": sample_cubes / model.num_cubes,
                   "note": "same size as the GPU arm; rank 0 only, host cores",
                   "integrals": {"volume": vol, "surface": surf}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
def multi_gpu_legs(args, n, rank, world, local_rank, ctx, model, geo, slab, state, step_bytes, ms_step, peak, barrier, stream,
                   host_barrier):
    """N > 1 only.  (1) Every number the ranks all-reduced through the library's NCCL communicator is recomputed by rank 0
    alone, slab by slab on its own GPU, and must agree (counts exactly, integrals to 1e-12).  (2) iso-work efficiency: the
    mean per-GPU fraction of the HBM roofline at N ranks over the fraction one GPU reaches on the N = 1 model, measured in
    this same process (per-GPU algorithmic bytes shrink with N because the surface work is proportional to n^2 / N, so the
    driver's cells/s ratio overstates the scaling).  (3) DiscreteGeometry leg: the same model with the level sets given as
    slab-local nodal vectors -- ghost node planes by ncclSend/ncclRecv and the consistency guard inside every step."""
    import torch
    import torch.distributed as dist
    import gecut
    from gecut import LevelSetCutter, Triangulation, EmbeddedBoundary, Measure
    from gecut.measures import integrate_measure
    SIDE = {-1: gecut.PHYSICAL_IN, 1: gecut.PHYSICAL_OUT}
    cells_plan, bnd_plan = workload_plan(args.workload)
    dev = f"cuda:{local_rank}"

    def one(cutter, mdl, g, lap=True):
        cg = cutter.cut(mdl, g)
        oms = [Triangulation(cg, SIDE[side], name) for name, side in cells_plan]
        if lap:
            for om in oms:
                ctx.check(ctx._lib.gecut_integrate_laplacian(om.handle, None, None), "gecut_integrate_laplacian")
        gas = [EmbeddedBoundary(cg, name) for name in bnd_plan]
        vols = [integrate_measure(Measure(om, 2)) for om in oms]
        surfs = [integrate_measure(Measure(ga, 2)) for ga in gas]
        out = [vols[0], surfs[0], float(cg.num_cutcells), float(cg.sizes.num_subcells), vols[-1]]
        for o in oms + gas:
            o.close()
        cg.close()
        return out

    # ---- (1) slab-by-slab on one GPU
    slabs = [None] * world
    dist.all_gather_object(slabs, tuple(slab))
    host_barrier()
    check = None
    if rank == 0:
        tot = [0.0] * 5
        for sl in slabs:
            part = one(LevelSetCutter(ctx, slab=tuple(sl)), model, geo, lap=False)
            tot = [a + b for a, b in zip(tot, part)]
        glob = [state["gvol"], state["gsurf"], float(state["global_counts"]["cut_cells"]),
                float(state["global_counts"]["subcells"]), state["gvols"][-1]]
        ok = (tot[2] == glob[2] and tot[3] == glob[3] and
              all(abs(a - b) <= 1e-12 * abs(a) for a, b in ((tot[0], glob[0]), (tot[1], glob[1]), (tot[4], glob[4]))))
        check = {"ok": bool(ok), "nccl": glob, "single_gpu_slab_by_slab": tot}
        if not ok:
            sys.stderr.write(f"the {world}-rank NCCL result differs from the single-GPU slab-by-slab result: {check}\n")
            sys.stderr.flush()
            os._exit(3)  # the other ranks wait on a barrier: leave without the interpreter's tear-down
    host_barrier()

    # ---- (2) iso-work efficiency
    bytes_t = torch.tensor([float(step_bytes)], dtype=torch.float64, device=dev)
    dist.all_reduce(bytes_t)
    frac_n = float(bytes_t[0]) / world / (ms_step * 1e-3) / 1e9 / peak
    iso = None
    model1, geo1, _ = make_workload(args.workload, n, 1)
    host_barrier()
    if rank == 0:
        c1 = LevelSetCutter(ctx)
        for _ in range(3):
            one(c1, model1, geo1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = max(5, min(args.steps, 20))
        a.record(stream)
        for _ in range(K):
            one(c1, model1, geo1)
        b.record(stream)
        torch.cuda.synchronize()
        ms1 = a.elapsed_time(b) / K
        # algorithmic bytes of the N = 1 model from its own counts
        cg = c1.cut(model1, geo1)
        oms = [Triangulation(cg, SIDE[side], name) for name, side in cells_plan]
        gas = [EmbeddedBoundary(cg, name) for name in bnd_plan]
        _, b1 = algorithmic_bytes(cg.sizes, model1, sum(o.num_sub for o in oms), sum(g_.num_sub for g_ in gas),
                                  ncellsel=len(cells_plan))
        for o in oms + gas:
            o.close()
        cg.close()
        frac_1 = b1 / (ms1 * 1e-3) / 1e9 / peak
        iso = {"frac_per_gpu_at_n": round(frac_n, 4), "frac_single_gpu_model": round(frac_1, 4),
               "iso_work_efficiency": round(frac_n / frac_1, 4), "single_gpu_ms_per_step": round(ms1, 4),
               "note": "mean per-GPU step_roofline.frac at N ranks / the same fraction of the N = 1 model timed on rank 0 in "
                       "this run (device-resident laplacian, same kernels)"}
    host_barrier()

    # ---- (3) DiscreteGeometry leg (skipped when the main leg already ran --nodal)
    nodal = None
    if not args.nodal:
        from gecut import DiscreteGeometry
        from gecut.cutters import evaluate_leaves_at_nodes
        from gecut.csg import find_unique_leaves, replace_data
        tree = geo.get_tree()
        payloads, oid_to_j = find_unique_leaves(tree)
        layer_nodes = model.num_nodes // (model.partition[-1] + 1)
        # DiscreteGeometry data cannot be re-balanced from the geometry: even slabs
        k0, k1 = slab_of(model, rank, world)
        bufs = []
        for pl in payloads:
            (buf,) = evaluate_leaves_at_nodes(model, [pl], ctx=ctx, device_output=True, slab=(k0, k1))
            if rank < world - 1:
                buf[-layer_nodes:] = float("nan")  # the ghost plane must come from the neighbour
            bufs.append(buf)
        dgeo = DiscreteGeometry(replace_data(lambda d: d, lambda d: (bufs[oid_to_j[id(d[0])]], d[1], d[2]), tree), model)
        ncut = LevelSetCutter(ctx, slab=(k0, k1), nodal_slab_local=True)
        red2 = torch.zeros(16, dtype=torch.float64, device=dev)

        def nstep():
            ctx.comm_exchange_ghost_planes(bufs, layer_nodes)
            if not ctx.comm_check_consistency(bufs, layer_nodes):
                raise SystemExit("inconsistent ghost planes")
            loc = one(ncut, model, dgeo)
            ctx.comm_allreduce_sum(loc, red2.data_ptr())
        for _ in range(3):
            nstep()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(args.steps):
            nstep()
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        g2 = ctx.comm_fetch(red2.data_ptr(), 5)
        same = (g2[2] == float(state["global_counts"]["cut_cells"]) and g2[3] == float(state["global_counts"]["subcells"]) and
                abs(g2[0] - state["gvol"]) <= 1e-12 * abs(state["gvol"]) and abs(g2[1] - state["gsurf"]) <= 1e-12 * abs(state["gsurf"]))
        if not same:
            raise SystemExit(f"DiscreteGeometry leg: global numbers {g2} differ from the closed-form leg")
        msn = float(t[0]) / args.steps
        nodal = {"ms_per_step": msn, "value": model.num_cubes / (msn * 1e-3), "unit": UNIT, "slabs": "even layers",
                 "ghost_plane_bytes_per_step_per_boundary": int(8 * layer_nodes * len(bufs)),
                 "what": "same model, level sets as a DiscreteGeometry: slab-local device-resident nodal vectors (ghost plane "
                         "poisoned with NaN at start), per step: grouped ncclSend/ncclRecv of the ghost node planes + the "
                         "consistency guard (sign compare + ncclAllReduce(min)) + cut + integrals + ncclAllReduce(sum); "
                         "global integrals and counts equal the closed-form leg's"}
        del bufs
    return {"check_vs_single_gpu": check, "iso_work": iso, "discrete_geometry_leg": nodal,
            "collectives": "inside libgecut.so (gecut_comm_*): NCCL on the context's stream, no host round trip per step"}


# --------------------------------------------------------------------------------------------------------------------
def north_star_csg_leg(ctx, n, stream, peak, steps=5, warmup=2):
    """N = 1 only: the literal north-star configuration -- the PoissonCSGCutFEM geometry (box ∩ sphere − 3 cylinders, ten level
    sets) on the n^3 HEX model: cut + PHYSICAL "csg" volume + EmbeddedBoundary("csg") area + Q1 8x8 Laplacian matrices, device
    resident, timed like the main leg (CUDA events on the context's stream).  Reported next to the main line so that the
    driver-run record carries the CSG number too; `python bench.py --workload c1_csg` is the full line for it."""
    import torch
    import gecut
    from gecut import LevelSetCutter, Triangulation, EmbeddedBoundary, Measure
    from gecut.measures import integrate_measure
    model, geo, desc = make_workload("c1_csg", n, 1)
    cutter = LevelSetCutter(ctx)
    out = {}

    def one():
        cg = cutter.cut(model, geo)
        om = Triangulation(cg, gecut.PHYSICAL_IN, "csg")
        ctx.check(ctx._lib.gecut_integrate_laplacian(om.handle, None, None), "gecut_integrate_laplacian")
        ga = EmbeddedBoundary(cg, "csg")
        vol = integrate_measure(Measure(om, 2))
        surf = integrate_measure(Measure(ga, 2))
        out.update(sizes=cg.sizes, nsel=om.num_sub, nfsel=ga.num_sub, vol=vol, surf=surf)
        om.close()
        ga.close()
        cg.close()

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    l0 = ctx.launch_count
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        one()
    b.record(stream)
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    launches = (ctx.launch_count - l0) // steps
    s = out["sizes"]
    _, step_bytes = algorithmic_bytes(s, model, out["nsel"], out["nfsel"])
    return {"workload": desc, "ms_per_step": ms, "value": model.num_cubes / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "warmup": warmup, "gpu_launches_per_step": int(launches),
            "step_roofline": {"alg_bytes_per_step": int(step_bytes), "achieved": round(step_bytes / (ms * 1e-3) / 1e9, 1),
                              "peak": peak, "unit": "GB/s", "frac": round(step_bytes / (ms * 1e-3) / 1e9 / peak, 4)},
            "counts": {"cut_cells": int(s.num_cutcells), "subcells": int(s.num_subcells),
                       "subcell_points": int(s.num_subcell_points), "subfacets": int(s.num_subfacets)},
            "integrals": {"volume": out["vol"], "surface": out["surf"]}}


# --------------------------------------------------------------------------------------------------------------------
def run_b200(args, n, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    import gecut
    from gecut import LevelSetCutter, Triangulation, EmbeddedBoundary, Measure, PHYSICAL
    from gecut import _lib
    from gecut.measures import integrate_measure
    import ctypes as C

    model, geo, desc = make_workload(args.workload, n, world)
    # z-slabs: work-balanced by default (even layers of classification + cut cells weighted by their measured cost; every
    # rank derives the same partition from the geometry, no communication), --uniform-slabs for the even split
    slab, slab_kind = None, "single GPU"
    if world > 1:
        if args.uniform_slabs:
            slab, slab_kind = slab_of(model, rank, world), f"z-slab x{world} (even layers)"
        else:
            from gecut.distributed import balanced_slab_ranges
            # cost of a cut cube in units of a classified cube (calibrated in round 1; re-deriving the weights from the round-2
            # kernel times -- 370 for the tets -- moved two layers between ranks and measured WORSE at 8 GPUs: 0.90 against
            # 0.81 ms per step, so the measured-good partition stays)
            cw = {"c3_tet": 300.0, "c3_hex": 760.0, "c5_bimaterial": 1400.0, "c1_csg": 1400.0}.get(args.workload, 600.0)
            ranges = balanced_slab_ranges(model, geo, world, cut_weight=cw)
            slab = ranges[rank]
            slab_kind = (f"z-slab x{world}, work-balanced (layers per rank {[b - a for a, b in ranges]}; cut cell = "
                         f"{cw:g} cells of classification)")
    stream = torch.cuda.current_stream()
    ctx = _lib.Context(local_rank, stream.cuda_stream)
    cutter = LevelSetCutter(ctx, slab=slab)
    # global integrals + counts: all-reduced INSIDE the library (gecut_comm_allreduce_sum: the local numbers travel as kernel
    # arguments, ncclAllReduce on the context's stream, nothing waits on the host); read back once after the loop
    red = torch.zeros(16, dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        from gecut.distributed import init_library_communicator
        init_library_communicator(ctx)

    host_pg = None
    if world > 1:
        # host-side (gloo) group for the sections only rank 0 works in: an NCCL barrier would leave a kernel spinning on
        # every other GPU, and a driver call that synchronises peer devices on rank 0 (cudaMalloc of a new block size with
        # peer access enabled) then never returns
        host_pg = dist.new_group(backend="gloo")
        import faulthandler
        faulthandler.dump_traceback_later(420, exit=True)  # a hang prints the Python stacks and ends the run

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=host_pg)

    state = {}

    dbg = []

    # ---- DiscreteGeometry input: every rank holds the nodal values of its own node planes; the ghost plane (first plane of
    # the next rank) arrives by NCCL send/recv inside every step, as in a re-cut loop on an evolving level set
    nodal_bufs, layer_nodes = [], 0
    if args.nodal:
        from gecut import DiscreteGeometry
        from gecut.cutters import evaluate_leaves_at_nodes
        from gecut.csg import find_unique_leaves, replace_data
        tree = geo.get_tree()
        payloads, oid_to_j = find_unique_leaves(tree)
        layer_nodes = model.num_nodes // (model.partition[-1] + 1)
        k0, k1 = slab if slab is not None else (0, model.partition[-1])
        for pl in payloads:
            (buf,) = evaluate_leaves_at_nodes(model, [pl], ctx=ctx, device_output=True, slab=(k0, k1))
            if world > 1 and rank < world - 1:
                buf[-layer_nodes:] = float("nan")  # the ghost plane is NOT known locally: it must come from the neighbour
            nodal_bufs.append(buf)
        geo = DiscreteGeometry(replace_data(lambda d: d, lambda d: (nodal_bufs[oid_to_j[id(d[0])]], d[1], d[2]), tree), model)
        cutter = LevelSetCutter(ctx, slab=slab, nodal_slab_local=slab is not None)
        desc += "; level sets given as a DiscreteGeometry (device-resident slab-local nodal values, NCCL ghost-plane exchange per step)"

    cells_plan, bnd_plan = workload_plan(args.workload)
    SIDE = {-1: gecut.PHYSICAL_IN, 1: gecut.PHYSICAL_OUT}

    def step(e2e=False, host=None):
        t = [time.perf_counter()]
        if nodal_bufs and world > 1:
            # ghost node planes of the slab-local level sets: grouped ncclSend/ncclRecv inside the library, stream-ordered
            # before the cut; then the reference's guard (isconsistent_bgcell_to_inoutcut) on the shared planes
            ctx.comm_exchange_ghost_planes(nodal_bufs, layer_nodes)
            if not ctx.comm_check_consistency(nodal_bufs, layer_nodes):
                raise SystemExit("inconsistent ghost planes")
        cg = cutter.cut(model, geo)
        t.append(time.perf_counter())
        oms = [Triangulation(cg, SIDE[side], name) for name, side in cells_plan]
        t.append(time.perf_counter())
        # device-resident matrices are stream-ordered (the call returns once the kernel is queued), so the host side of
        # the boundary selection and of the integrals below overlaps the Laplacian kernels
        for i, om in enumerate(oms):
            ctx.check(ctx._lib.gecut_integrate_laplacian(om.handle, host["K_cut"][i].ctypes.data if e2e else None, None),
                      "gecut_integrate_laplacian")
        t.append(time.perf_counter())
        gas = [EmbeddedBoundary(cg, name) for name in bnd_plan]
        t.append(time.perf_counter())
        vols = [integrate_measure(Measure(om, 2)) for om in oms]
        surfs = [integrate_measure(Measure(ga, 2)) for ga in gas]
        t.append(time.perf_counter())
        dbg.append([round((b - a) * 1e3, 2) for a, b in zip(t[:-1], t[1:])])
        if e2e:
            t0 = time.perf_counter()
            cg.to_host(buffers=host["layout"])  # blocking: returns when the last array has landed
            state.setdefault("d2h_s", []).append(time.perf_counter() - t0)
        vol, surf = vols[0], surfs[0]
        local = [vol, surf, float(cg.num_cutcells), float(cg.sizes.num_subcells), vols[-1]]
        if world > 1:  # global integrals + counts: the only data-path collective, device resident
            ctx.comm_allreduce_sum(local, red.data_ptr())
        state.update(sizes=cg.sizes, nsel=sum(om.num_sub for om in oms), nfsel=sum(ga.num_sub for ga in gas), vol=vol,
                     surf=surf, vols=[float(v) for v in vols], local=local)
        for o in oms + gas:
            o.close()
        cg.close()
        return vol, surf

    # ---- device-resident throughput ----------------------------------------------------------------------------
    # the clock sampler is started BEFORE the warm-up: nvidia-smi's start-up (NVML init) stalls the driver for ~100 ms,
    # which must not land inside the (few ms per step) timed region; it keeps sampling during the timed steps
    sampler = ClockSampler(local_rank) if (rank == 0 and not os.environ.get("GECUT_NO_SAMPLER")) else None
    if sampler:
        sampler.start()
    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    if sampler:
        sampler.clear()
        sampler.period = 0.005
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    wall = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        step()
        wall.append((time.perf_counter() - t0) * 1e3)
    ev1.record(stream)
    if os.environ.get("GECUT_BENCH_DEBUG"):
        sys.stderr.write("per-step wall ms: " + " ".join(f"{w:.1f}" for w in wall) + "\n")
        sys.stderr.write("per-call [cut, physical, laplacian, boundary, measures] ms of the timed steps:\n")
        for row in dbg[-args.steps:]:
            sys.stderr.write("   " + str(row) + "\n")
    barrier()
    if sampler:
        sampler.period = 0.02
    timed_clock_samples = len(sampler.samples) if sampler else 0
    launches = ctx.launch_count - l0
    glob = ctx.comm_fetch(red.data_ptr(), 5) if world > 1 else list(state["local"])
    # global numbers under their own keys: later steps (profiling, end to end) overwrite the per-rank entries of `state`
    state["gvol"], state["gsurf"] = glob[0], glob[1]
    state["gvols"] = [glob[0]] + ([glob[4]] if len(state["vols"]) > 1 else [])
    state["global_counts"] = {"cut_cells": int(glob[2]), "subcells": int(glob[3])}
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms[0])
    cubes_total = model.num_cubes  # all ranks together
    value = cubes_total * args.steps / (ms_total * 1e-3)
    s = state["sizes"]

    # ---- per-kernel timing for the roofline (CUDA events around every launch, same stream) -------------------------
    PROF_STEPS = 3
    ctx.profile(True)
    for _ in range(PROF_STEPS):
        step()
    rep = ctx.profile_report()
    ctx.profile(False)
    kbytes, step_bytes = algorithmic_bytes(s, model, state["nsel"], state["nfsel"], ncellsel=len(cells_plan), nodal=bool(nodal_bufs))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    kernels = []
    if "k_classify_cull" in rep:  # the two classification kernels share the algorithmic bytes of the classification
        c0, t0 = rep.pop("k_classify_cull")
        c1, t1 = rep.get("k_classify", (0, 0.0))
        rep["k_classify"] = (max(c0, c1), t0 + t1)
        state["classify_split_ms"] = {"k_classify_cull": round(t0 / PROF_STEPS, 4), "k_classify(march)": round(t1 / PROF_STEPS, 4)}
    for name, (cnt, tot) in rep.items():
        per_step_ms = tot / PROF_STEPS
        kb = kbytes.get(name)
        if name == "k_classify" and "classify_split_ms" in state:
            name = "k_classify (k_classify_cull + k_classify)"
        kernels.append({"name": name, "launches_per_step": cnt / PROF_STEPS, "ms_per_step": round(per_step_ms, 4),
                        "alg_bytes_per_step": kb,
                        "gbs": round(kb / (per_step_ms * 1e-3) / 1e9, 1) if kb and per_step_ms > 0 else None})
    kernels.sort(key=lambda k: -k["ms_per_step"])
    top = kernels[0]
    top_launches = max(1.0, top["launches_per_step"])
    achieved = (top["alg_bytes_per_step"] or 0) / (top["ms_per_step"] * 1e-3) / 1e9 if top["ms_per_step"] > 0 else 0.0
    traffic = None
    try:  # DRAM bytes per launch of this kernel from the committed ncu --set full capture (not measured live)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(f"{args.workload}:{top['name']}")
        if world > 1 or n != 512:
            traffic = None  # the capture was taken on the single-GPU 512^3 workload
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": top["name"], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_kind,
                "alg_bytes_per_launch": int((top["alg_bytes_per_step"] or 0) / top_launches),
                "ms_per_launch": round(top["ms_per_step"] / top_launches, 4),
                "kernel_share_of_step": round(top["ms_per_step"] / sum(k["ms_per_step"] for k in kernels), 3)}
    ms_step = ms_total / args.steps
    step_roofline = {"alg_bytes_per_step": step_bytes, "achieved": round(step_bytes / (ms_step * 1e-3) / 1e9, 1),
                     "peak": peak, "unit": "GB/s", "frac": round(step_bytes / (ms_step * 1e-3) / 1e9 / peak, 4)}

    # ---- multi-GPU: the N-rank result against a single-GPU slab-by-slab run, iso-work efficiency, DiscreteGeometry leg ----
    multi = None
    if world > 1:
        multi = multi_gpu_legs(args, n, rank, world, local_rank, ctx, model, geo, slab, state, step_bytes, ms_step, peak,
                               barrier, stream, host_barrier)

    # ---- end to end through the C ABI with HOST buffers ------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        D, L = int(s.D), int(s.num_levelsets)
        nv = int(s.num_vertices_per_bgcell)
        host = {"layout": {}, "K_cut": None}
        shapes = {
            "ls_to_bgcell_to_inoutcut": ((L, s.num_bgcells), np.int8), "cutcell_to_cell": ((s.num_cutcells,), np.int32),
            "sc_cell_to_points": ((s.num_subcells, D + 1), np.int32), "sc_cell_to_bgcell": ((s.num_subcells,), np.int32),
            "sc_point_to_coords": ((s.num_subcell_points, D), np.float64),
            "sc_point_to_rcoords": ((s.num_subcell_points, D), np.float64),
            "ls_to_subcell_to_inout": ((L, s.num_subcells), np.int8),
            "sf_facet_to_points": ((s.num_subfacets, D), np.int32), "sf_facet_to_bgcell": ((s.num_subfacets,), np.int32),
            "sf_facet_to_normal": ((s.num_subfacets, D), np.float64),
            "sf_point_to_coords": ((s.num_subfacet_points, D), np.float64),
            "sf_point_to_rcoords": ((s.num_subfacet_points, D), np.float64),
            "ls_to_subfacet_to_inout": ((L, s.num_subfacets), np.int8)}
        tdt = {np.int8: torch.int8, np.int32: torch.int32, np.float64: torch.float64}
        d2h = 0
        for k, (shape, dt) in shapes.items():
            if int(s.subfacet_points_alias) and k in ("sf_point_to_coords", "sf_point_to_rcoords"):
                continue  # one level set: these ARE the subcell point arrays; the binding shares one host array
            t = torch.empty(tuple(int(x) for x in shape), dtype=tdt[dt], pin_memory=True)
            host["layout"][k] = t.numpy()
            d2h += t.numel() * t.element_size()
        host["K_cut"] = []
        for _ in cells_plan:
            kt = torch.empty((int(s.num_cutcells), nv, nv), dtype=torch.float64, pin_memory=True)
            host["K_cut"].append(kt.numpy())
            d2h += kt.numel() * 8 + 16
        h2d = C.sizeof(_lib.Grid) + L * C.sizeof(_lib.LeafRec) + 3 * 4 * 8  # grid + leaf records + 3 boolean programs
        e2e_steps = max(1, min(args.steps, 3))
        step(True, host)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(e2e_steps):
            step(True, host)
        e1.record(stream)
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        # where the end-to-end time goes: the blocking copy-out of the layout (max over ranks), and the floor of this box for
        # it -- one plain cudaMemcpyAsync of the same number of bytes into pinned memory, all ranks at the same time
        # (one level set: the sub-facet point arrays are the sub-cell point arrays -- the dict holds them twice, they travel once)
        lay_bytes = sum(a.nbytes for a in {a.ctypes.data: a for a in host["layout"].values() if a.size}.values())
        d2h_ms = torch.tensor([1e3 * sum(state["d2h_s"][-e2e_steps:]) / e2e_steps], dtype=torch.float64, device=f"cuda:{local_rank}")
        src = torch.empty(lay_bytes, dtype=torch.uint8, device=f"cuda:{local_rank}")
        dst = torch.empty(lay_bytes, dtype=torch.uint8, pin_memory=True)
        dst.copy_(src, non_blocking=True)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(2):
            dst.copy_(src, non_blocking=True)
        f1.record(stream)
        barrier()
        floor_ms = torch.tensor([f0.elapsed_time(f1) / 2], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(d2h_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(floor_ms, op=dist.ReduceOp.MAX)
        del src, dst
        d2h_detail = {"layout_bytes_per_gpu": int(lay_bytes), "copy_out_ms_per_step": round(float(d2h_ms[0]), 3),
                      "copy_out_gbs_per_gpu": round(lay_bytes / (float(d2h_ms[0]) * 1e-3) / 1e9, 2),
                      "floor_ms": round(float(floor_ms[0]), 3),
                      "floor_gbs_per_gpu": round(lay_bytes / (float(floor_ms[0]) * 1e-3) / 1e9, 2),
                      "copy_out_over_floor": round(float(floor_ms[0]) / float(d2h_ms[0]), 3),
                      "note": "floor = one cudaMemcpyAsync of the same bytes device -> pinned host, every rank at the same time "
                              "(max over ranks); copy_out = gecut_result_copy's 11 per-array copies"}
        e2e = {"value": cubes_total * e2e_steps / (float(ems[0]) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h": d2h_detail,
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": float(ems[0]) / e2e_steps,
               "what": "gecut_cut + selectors + integrals through the C ABI, whole EmbeddedDiscretization layout and the "
                       "element matrices copied into pinned HOST buffers every step; the input is an AnalyticalGeometry "
                       "(a few hundred bytes of leaf records), exactly what the reference API receives"}

    # ---- north-star CSG configuration (N = 1, default workload only): same context, the main leg's memory handed back first
    csg = None
    if world == 1 and args.workload == "c3_tet" and not args.no_csg and n >= 256:
        ctx.trim()
        try:  # a side leg: whatever happens here must not take the main line down
            csg = north_star_csg_leg(ctx, n, stream, peak)
        except Exception as exc:  # noqa: BLE001
            csg = {"error": f"{type(exc).__name__}: {exc}"}
        ctx.trim()

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_cpu = args.cpu_n or {"c2_disk": 4096, "c5_bimaterial": 128, "c1_csg": 64}.get(args.workload, 320)
        v, dt, sample = cpu_sample(args.workload, n_cpu)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample, "seconds": round(dt, 2)}

    clocks = None
    if sampler:
        clocks = sampler.stop()
        clocks["samples_timed_region"] = timed_clock_samples
        clocks["note"] = ("sampled from the first timed step to the end of the end-to-end leg (the device-resident timed "
                          "region alone lasts tens of ms = a handful of samples at the 5 ms period)")
    if rank == 0:
        lb = layout_bytes(s)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": desc, "cells_per_gpu": int(model.num_cubes // world), "cells_rank0": int(model.num_cubes // model.partition[-1] * ((slab[1] - slab[0]) if slab else model.partition[-1])),
                       "bg_simplices_per_gpu": int(s.num_bgcells), "partition": list(model.partition),
                       "parallelism": slab_kind,
                       "l2": "no L2 flush needed: every step rewrites its whole working set "
                             f"({(sum(lb.values())) / 1e9:.2f} GB of outputs per GPU >> 126 MB L2), nothing is reused across steps",
                       "per_gpu_counts": {"cut_cells": int(s.num_cutcells), "subcells": int(s.num_subcells),
                                          "subcell_points": int(s.num_subcell_points), "subfacets": int(s.num_subfacets)},
                       "global_counts": state["global_counts"],
                       "integrals": {"volume": state["gvol"], "surface": state["gsurf"], "volumes": state["gvols"]}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "multi_gpu": multi,
            "roofline": roofline, "step_roofline": step_roofline, "kernels": kernels[:12], "cpu_baseline": cpu,
            "north_star_csg": csg, "classify_split_ms": state.get("classify_split_ms"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
