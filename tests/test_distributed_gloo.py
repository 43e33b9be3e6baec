"""N>1 host logic on CPU with gloo (world_size 2): slab partition, global offsets (all_gather), global integrals
(all_reduce) and the ghost node-plane exchange for slab-local NODAL level sets.  The local "cut" numbers come from the
oracle run on the whole (small) model and filtered per slab -- the CUDA path itself is covered by the -m gpu tests
(test_slab_cut_is_a_slice_of_the_whole_cut)."""
import os
import socket
import sys

import numpy as np
import pytest

import gecut
from gecut.distributed import slab_range, slab_cell_offset, slab_node_offset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_range_partitions_the_layers():
    for nlast in (1, 7, 8, 128, 513):
        for world in (1, 2, 3, 4, 8):
            if world > nlast:
                with pytest.raises(ValueError):
                    slab_range(nlast, 0, world)
                continue
            r = [slab_range(nlast, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nlast
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    m = gecut.simplexify(gecut.CartesianDiscreteModel((0, 0, 0), (1, 1, 1), (4, 3, 8)))
    assert slab_cell_offset(m, 2) == 2 * 12 * 6 and slab_node_offset(m, 2) == 2 * 20


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import gecut as G
    from gecut.distributed import (global_offsets, global_sum, exchange_ghost_plane, slab_range, slab_cell_offset,
                                   DistributedLevelSetCutter)
    import oracle_binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geo = G.sphere(0.7)
        box = G.get_metadata(geo)
        model = G.CartesianDiscreteModel(box.pmin, box.pmax, (10, 9, 8))
        oc = ob.oracle_cut(model, geo)
        cutter = DistributedLevelSetCutter()
        assert (cutter.rank, cutter.world) == (rank, world)
        k0, k1 = cutter.slab(model)
        assert (k0, k1) == slab_range(8, rank, world)
        off = slab_cell_offset(model, k0)
        ncl = slab_cell_offset(model, k1) - off
        c2bg = oc.array("subcells.cell_to_bgcell")
        mine = (c2bg > off) & (c2bg <= off + ncl)
        ids = np.flatnonzero(mine) + 1
        prog = [0]
        sel = oc.select_subcells(prog, -1)
        sel_mine = sel[np.isin(sel, ids)]
        bg = oc.select_bgcells(prog, -1)
        bg_mine = bg[(bg > off) & (bg <= off + ncl)]
        vol_local = oc.subcell_measures(sel_mine).sum() + len(bg_mine) * oc.bgcell_measure(1)
        local = {"subcells": int(mine.sum()), "bgcells": ncl}
        offs, tot = global_offsets(local)
        assert tot["subcells"] == oc.Nsc and tot["bgcells"] == oc.Nc
        assert offs["bgcells"] == off
        assert offs["subcells"] == int((c2bg <= off).sum())
        (vol,) = global_sum([vol_local])
        vol_ref = oc.subcell_measures(sel).sum() + len(bg) * oc.bgcell_measure(1)
        assert abs(vol - vol_ref) <= 1e-12 * vol_ref
        # ghost plane exchange for slab-local NODAL values
        layer_nodes = 11 * 10
        vals = ob.node_values(model, ob.make_leaf(0, R=0.7))
        top = k1 + 1 if rank == world - 1 else k1
        owned = torch.from_numpy(vals[k0 * layer_nodes: top * layer_nodes].copy())
        full = exchange_ghost_plane(owned, layer_nodes)
        assert np.array_equal(full.numpy(), vals[k0 * layer_nodes: (k1 + 1) * layer_nodes])
        # in-place form (persistent slab buffer of a re-cut loop): the ghost plane is unknown (NaN) before the exchange
        from gecut.distributed import exchange_ghost_plane_
        buf = torch.from_numpy(vals[k0 * layer_nodes: (k1 + 1) * layer_nodes].copy())
        if rank < world - 1:
            buf[-layer_nodes:] = float("nan")
        exchange_ghost_plane_(buf, layer_nodes)
        assert np.array_equal(buf.numpy(), vals[k0 * layer_nodes: (k1 + 1) * layer_nodes])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"


def test_balanced_slabs_cover_the_layers_and_even_out_the_estimated_work():
    """Work-balanced z-slabs (the cut work is surface-proportional): contiguous, non-empty, covering; the
    estimated work of the most loaded rank drops for the two-cylinder model whose cylinders cross only some slabs."""
    import gecut as G
    from gecut.distributed import balanced_slab_ranges, estimate_cut_cells_per_layer, slab_range
    steel = G.union(G.cylinder(0.4, x0=(0.0, 1.0, 0.7)), G.cylinder(0.4, x0=(0.0, 1.0, 2.3)), name="steel")
    geo = G.union(steel, G.complement(steel, name="concrete"))
    model = G.CartesianDiscreteModel((0.0, 0.0, 0.0), (3.0, 2.0, 3.0), (96, 96, 96))
    est = estimate_cut_cells_per_layer(model, geo, coarse=32)
    assert est.shape == (96,) and est.min() == 0.0 and est.max() > 0.0
    cpl = model.num_cubes // 96

    def worst(ranges, cw):
        return max((b - a) * cpl + cw * est[a:b].sum() for a, b in ranges)

    for world in (2, 3, 8):
        rs = balanced_slab_ranges(model, geo, world, cut_weight=1400.0, coarse=32)
        assert rs[0][0] == 0 and rs[-1][1] == 96 and all(a < b for a, b in rs)
        assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
        even = [slab_range(96, r, world) for r in range(world)]
        assert worst(rs, 1400.0) <= worst(even, 1400.0) + 1e-9
    assert worst(balanced_slab_ranges(model, geo, 8, 1400.0, coarse=32), 1400.0) < 0.8 * worst([slab_range(96, r, 8) for r in range(8)], 1400.0)
    # value-vector leaves cannot be evaluated on the host: even split
    dg = G.DiscreteGeometry(np.zeros(model.num_nodes), model, name="d")
    assert balanced_slab_ranges(model, dg, 4) == [slab_range(96, r, 4) for r in range(4)]
    # more ranks than layers is an error, one rank owns everything
    assert balanced_slab_ranges(model, geo, 1) == [(0, 96)]
    with pytest.raises(ValueError):
        balanced_slab_ranges(model, geo, 97)
