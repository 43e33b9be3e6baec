"""Two ranks, two GPUs, NCCL: the multi-GPU path against the single-GPU result.

Each rank cuts its z-slab on its own GPU; the collectives run inside libgecut.so (gecut_comm_*: ghost node planes by
grouped ncclSend/ncclRecv, global sums by ncclAllReduce, sizes by ncclAllGather, the consistency flag by
ncclAllReduce(min)) and, for comparison, through torch.distributed.  Asserted: every rank's layout is the slice of the
whole-model cut; the all-reduced integrals and counts equal the whole-model numbers; a NaN-poisoned ghost plane is
replaced by the neighbour's values; a sign flip in one rank's copy of the shared plane lowers the consistency flag on
both ranks (isconsistent_bgcell_to_inoutcut, src/Distributed/DistributedDiscretizations.jl:136-174).
Skipped below two GPUs (NCCL refuses two ranks on one device); run with `gpurun --gpus 2`.
"""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models(G):
    out = []
    geo = G.sphere(0.7)
    box = G.get_metadata(geo)
    out.append(("sphere_tet", G.simplexify(G.CartesianDiscreteModel(box.pmin, box.pmax, (40, 36, 34))), geo, None))
    out.append(("sphere_hex", G.CartesianDiscreteModel(box.pmin, box.pmax, (33, 40, 37)), geo, None))
    g1 = G.cylinder(0.4, x0=(0.0, 1.0, 0.7))
    g2 = G.cylinder(0.4, x0=(0.0, 1.0, 2.3))
    steel = G.union(g1, g2, name="steel")
    out.append(("bimaterial", G.CartesianDiscreteModel((0., 0., 0.), (3., 2., 3.), (30, 24, 40)),
                G.union(steel, G.complement(steel, name="concrete")), "steel"))
    return out


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
        import gecut as G
        from gecut import _lib
        from gecut.distributed import (DistributedLevelSetCutter, init_library_communicator, exchange_ghost_plane_,
                                       slab_range)
        from gecut.measures import integrate_measure
        from gecut.cutters import evaluate_leaves_at_nodes
        from gecut.csg import find_unique_leaves, replace_data
        ctx = _lib.Context(rank, torch.cuda.current_stream().cuda_stream)
        assert init_library_communicator(ctx) == (rank, world)
        assert ctx.comm_rank_world == (rank, world)
        red = torch.zeros(16, dtype=torch.float64, device=f"cuda:{rank}")
        for label, model, geo, name in _models(G):
            # ---- closed-form leaves: slab cut == slice of the whole cut, global numbers == whole-model numbers
            whole = G.LevelSetCutter(ctx).cut(model, geo)
            hw = whole.to_host()
            dcut = DistributedLevelSetCutter(ctx)
            part = dcut.cut(model, geo)
            hp = part.to_host()
            off, nc = int(part.sizes.cell_offset), int(part.sizes.num_bgcells)
            assert np.array_equal(hp["ls_to_bgcell_to_inoutcut"], hw["ls_to_bgcell_to_inoutcut"][:, off:off + nc]), label
            cw = hw["cutcell_to_cell"]
            assert np.array_equal(hp["cutcell_to_cell"] + off, cw[(cw > off) & (cw <= off + nc)]), label
            scm = (hw["sc_cell_to_bgcell"] > off) & (hw["sc_cell_to_bgcell"] <= off + nc)
            assert np.array_equal(hp["sc_cell_to_bgcell"] + off, hw["sc_cell_to_bgcell"][scm]), label
            assert np.array_equal(hp["ls_to_subcell_to_inout"], hw["ls_to_subcell_to_inout"][:, scm]), label
            # rank-major global layout: offsets from the all-gathered sizes (NCCL inside the library)
            sizes = ctx.comm_allgather_sizes(part.handle)
            assert len(sizes) == world and int(sizes[rank].num_subcells) == int(part.sizes.num_subcells)
            sc_off = sum(int(s.num_subcells) for s in sizes[:rank])
            pt_off = sum(int(s.num_subcell_points) for s in sizes[:rank])
            assert sum(int(s.num_subcells) for s in sizes) == int(whole.sizes.num_subcells), label
            assert sum(int(s.num_cutcells) for s in sizes) == int(whole.sizes.num_cutcells), label
            assert sc_off == int(np.flatnonzero(scm)[0]) if scm.any() else True
            nsc, npt = hp["sc_cell_to_points"].shape[0], hp["sc_point_to_coords"].shape[0]
            assert np.array_equal(hp["sc_cell_to_points"] + pt_off, hw["sc_cell_to_points"][sc_off:sc_off + nsc]), label
            assert np.array_equal(hp["sc_point_to_coords"], hw["sc_point_to_coords"][pt_off:pt_off + npt]), label
            assert np.array_equal(hp["sc_point_to_rcoords"], hw["sc_point_to_rcoords"][pt_off:pt_off + npt]), label
            sf_off = [sum(int(s.num_subfacets) for s in sizes[:r]) for r in range(world + 1)]
            assert sf_off[-1] == int(whole.sizes.num_subfacets), label

            def integrals(cg):
                om, ga = G.Triangulation(cg, G.PHYSICAL, name), G.EmbeddedBoundary(cg, name)
                v = (integrate_measure(G.Measure(om, 2)), integrate_measure(G.Measure(ga, 2)))
                om.close(); ga.close()
                return v
            vol_w, surf_w = integrals(whole)
            vol_l, surf_l = integrals(part)
            ctx.comm_allreduce_sum([vol_l, surf_l, float(part.sizes.num_cutcells), float(part.sizes.num_subcells)],
                                   red.data_ptr())
            gv = ctx.comm_fetch(red.data_ptr(), 4)
            assert abs(gv[0] - vol_w) <= 1e-12 * abs(vol_w) and abs(gv[1] - surf_w) <= 1e-12 * abs(surf_w), (label, gv)
            assert gv[2] == float(whole.sizes.num_cutcells) and gv[3] == float(whole.sizes.num_subcells), (label, gv)
            assert ctx.comm_check_consistency() is True
            # ---- DiscreteGeometry: slab-local nodal vectors, ghost plane poisoned, then exchanged over NCCL
            k0, k1 = slab_range(model.partition[-1], rank, world)
            layer_nodes = model.num_nodes // (model.partition[-1] + 1)
            tree = geo.get_tree()
            payloads, oid_to_j = find_unique_leaves(tree)
            bufs = evaluate_leaves_at_nodes(model, payloads, ctx=ctx, device_output=True, slab=(k0, k1))
            full = evaluate_leaves_at_nodes(model, payloads, ctx=ctx, device_output=True)
            for b, f in zip(bufs, full):
                assert torch.equal(b, f[k0 * layer_nodes:(k1 + 1) * layer_nodes]), label  # slab-only evaluation
            dgeo = G.DiscreteGeometry(replace_data(lambda d: d, lambda d: (bufs[oid_to_j[id(d[0])]], d[1], d[2]), tree), model)
            lcut = G.LevelSetCutter(ctx, slab=(k0, k1), nodal_slab_local=True)
            for how in ("library", "torch"):
                if rank < world - 1:
                    for b in bufs:
                        b[-layer_nodes:] = float("nan")
                    assert ctx.comm_check_consistency(bufs, layer_nodes) is False  # NaN never matches the owner's plane
                else:
                    assert ctx.comm_check_consistency(bufs, layer_nodes) is False  # the flag is global
                if how == "library":
                    ctx.comm_exchange_ghost_planes(bufs, layer_nodes)
                else:
                    for b in bufs:
                        exchange_ghost_plane_(b, layer_nodes)
                for b, f in zip(bufs, full):
                    assert torch.equal(b, f[k0 * layer_nodes:(k1 + 1) * layer_nodes]), (label, how)
                assert ctx.comm_check_consistency(bufs, layer_nodes) is True
                nod = lcut.cut(model, dgeo).to_host()
                for k in hp:
                    assert np.array_equal(nod[k], hp[k]), (label, how, k)
            # ---- distributed AgFEM aggregation (n-cube models): slab facets + ghost cell layers over NCCL == whole model
            if not model.simplexified:
                from gecut import aggregate, aggregate_slabs, AggregateCutCellsByThreshold
                for thr in (1.0, 0.5):
                    strategy = AggregateCutCellsByThreshold(thr)
                    for side in (G.IN, G.OUT):
                        try:
                            want = aggregate(strategy, whole, name, side)
                        except _lib.GecutError:
                            want = None  # the reference's @assert all_aggregated: every rank must see the same failure
                        try:
                            got = aggregate_slabs(strategy, part, None, name, side, comm=True)
                        except _lib.GecutError:
                            got = None
                        assert (want is None) == (got is None), (label, thr, side)
                        if want is not None:
                            assert np.array_equal(got, want[off:off + nc]), (label, thr, side)
                # GhostSkeleton of the rank's local model: cells across the shared plane come over NCCL
                from gecut import GhostSkeleton, GhostSkeleton_slabs
                wf = G.LevelSetCutter(ctx).cut_facets(model, geo)
                gn = {tuple(r): f for f, r in enumerate(wf.to_host(topology=True)["facet_to_nodes"])}
                lf = G.LevelSetCutter(ctx, slab=part.slab).cut_facets(model, geo) if part.slab else wf
                k0s = part.slab[0] if part.slab else 0
                l2g = np.array([gn[tuple(int(v) + k0s * layer_nodes for v in r)]
                                for r in lf.to_host(topology=True)["facet_to_nodes"]])
                want_mask = GhostSkeleton(whole, None, name)
                got_mask = GhostSkeleton_slabs(part, None, name, comm=True)
                assert np.array_equal(got_mask, want_mask[l2g]), label
            # a sign flip in ONE rank's copy of the shared plane is seen by every rank
            if rank == 0:
                i = int(torch.argmax(torch.abs(bufs[0][-layer_nodes:])))
                bufs[0][-layer_nodes + i] = -bufs[0][-layer_nodes + i]
            assert ctx.comm_check_consistency(bufs, layer_nodes) is False
            whole.close(); part.close()
        ctx.comm_free()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:
        q.put((rank, traceback.format_exc()))


def test_two_rank_nccl_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL does not put two ranks on one device)")
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    try:
        for _ in range(2):
            r, msg = q.get(timeout=600)
            res[r] = msg
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.terminate()
    assert res.get(0) == "ok" and res.get(1) == "ok", res


def test_library_communicator_single_rank():
    """One rank: the communicator entry points work on a world of 1 (the path a single-GPU host takes with the same code)."""
    import torch
    import gecut as G
    from gecut import _lib
    ctx = _lib.Context(0, torch.cuda.current_stream().cuda_stream)
    with pytest.raises(_lib.GecutError):
        ctx.comm_check_consistency()  # no communicator yet
    ctx.comm_init(ctx.comm_unique_id(), 0, 1)
    assert ctx.comm_rank_world == (0, 1)
    red = torch.zeros(4, dtype=torch.float64, device="cuda:0")
    ctx.comm_allreduce_sum([1.5, -2.0, 3.0], red.data_ptr())
    assert ctx.comm_fetch(red.data_ptr(), 3) == [1.5, -2.0, 3.0]
    assert ctx.comm_check_consistency() is True
    model = G.CartesianDiscreteModel((-1, -1, -1), (1, 1, 1), (8, 8, 8))
    cg = G.LevelSetCutter(ctx).cut(model, G.sphere(0.7))
    sizes = ctx.comm_allgather_sizes(cg.handle)
    assert len(sizes) == 1 and int(sizes[0].num_subcells) == int(cg.sizes.num_subcells)
    v = torch.arange(3 * 81, dtype=torch.float64, device="cuda:0")
    w = v.clone()
    ctx.comm_exchange_ghost_planes([v], 81)  # world of 1: nothing moves
    assert torch.equal(v, w)
    with pytest.raises(_lib.GecutError):
        ctx.comm_init(ctx.comm_unique_id(), 0, 1)  # already has one
    # the distributed aggregation on a world of 1 is the whole-model aggregation
    from gecut import aggregate, aggregate_slabs, AggregateAllCutCells
    assert np.array_equal(aggregate_slabs(AggregateAllCutCells(), cg, comm=True), aggregate(AggregateAllCutCells(), cg))
    cg.close()
    ctx.comm_free()
    ctx.close()


def test_slab_local_closure_leaves():
    """host-callable leaves (popcorn, closures) in slab-local NODAL mode are evaluated on the slab's own
    node planes, so a slab with k0 > 0 cuts like the same layers of the whole model."""
    import gecut as G
    geo = G.popcorn()
    box = G.get_metadata(geo)
    model = G.CartesianDiscreteModel(box.pmin, box.pmax, (14, 12, 16))
    k0, k1 = 5, 13
    ref = G.LevelSetCutter(slab=(k0, k1)).cut(model, geo).to_host()
    got = G.LevelSetCutter(slab=(k0, k1), nodal_slab_local=True).cut(model, geo).to_host()
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k
    # the slab's own node planes, in global order
    layer_nodes = model.num_nodes // (model.partition[-1] + 1)
    assert np.array_equal(model.node_coordinates((k0, k1)), model.node_coordinates()[k0 * layer_nodes:(k1 + 1) * layer_nodes])
