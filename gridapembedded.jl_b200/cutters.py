"""`Cutter` plugin API and the B200 `LevelSetCutter` (reference: `src/Interfaces/Cutters.jl:22-68`,
`src/LevelSetCutters/LevelSetCutters.jl:77-205`).

`cut(cutter, background, geom)` flattens the CSG tree on the host (unique leaves in
`_find_unique_leaves` order -> leaf records / nodal value vectors), calls `gecut_cut` through the C ABI and
wraps the device-resident result in an `EmbeddedDiscretization`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .csg import Geometry, find_unique_leaves
from .geometries import LEAF_NODAL, AnalyticalGeometry, DiscreteGeometry, LeafFunction


class Cutter:
    """`abstract type Cutter` (`Cutters.jl:22`): implement `cut` and `compute_bgcell_to_inoutcut`."""

    def cut(self, background, geom):
        raise NotImplementedError("abstract method")

    def compute_bgcell_to_inoutcut(self, background, geom):
        raise NotImplementedError("abstract method")

    def cut_facets(self, background, geom):
        raise NotImplementedError("abstract method")

    def compute_bgfacet_to_inoutcut(self, background, geom):
        raise NotImplementedError("abstract method")


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


def flatten_geometry(model, geo: Geometry, host_only: bool = False, num_nodal: Optional[int] = None,
                     nodal_planes: Optional[Tuple[int, int]] = None):
    """Host half of `initial_sub_triangulation` (`CutTriangulations.jl:443-454`): the unique leaves of the tree in
    `_find_unique_leaves` order as leaf records, plus per-leaf nodal values for NODAL leaves.

    `nodal_planes=(k0,k1)` (slab-local NODAL mode): value vectors hold, and host-callable leaves (popcorn, arbitrary
    closures) are evaluated on, only the node planes k0..k1 of the last direction -- every NODAL vector handed to the
    library then has the same slab-local extent `num_nodal`.

    Returns (leaves: List[LeafFunction-like], nodal: List[array|None], oid_to_ls: dict id(payload)->ls (0-based))."""
    tree = geo.get_tree()
    payloads, oid_to_ls = find_unique_leaves(tree)
    if len(payloads) > _lib.MAX_LEVELSETS:
        raise NotImplementedError(f"more than {_lib.MAX_LEVELSETS} unique level sets")
    leaves, nodal = [], []
    for p in payloads:
        if isinstance(p, LeafFunction):
            if p.kind == LEAF_NODAL:
                vals = np.ascontiguousarray(p(model.node_coordinates(nodal_planes)), dtype=np.float64)
                if vals.shape != (model.num_nodes if num_nodal is None else num_nodal,):
                    raise ValueError("a host level-set function must return one value per node")
                leaves.append(p)
                nodal.append(vals)
            else:
                if p.dim != model.D:
                    raise ValueError(f"{p.dim}-D leaf used on a {model.D}-D background model")
                leaves.append(p)
                nodal.append(None)
        else:  # vector of nodal values (DiscreteGeometry leaf)
            n = p.shape[0]
            expected = model.num_nodes if num_nodal is None else num_nodal
            if n != expected:
                raise ValueError(f"expected {expected} nodal values, got {n}")
            if host_only and _is_torch_tensor(p):
                p = p.detach().cpu().numpy()
            leaves.append(LeafFunction(LEAF_NODAL, model.D))
            nodal.append(p)
    return leaves, nodal, oid_to_ls


def make_grid(model, slab: Optional[Tuple[int, int]] = None) -> _lib.Grid:
    g = _lib.Grid()
    g.D = model.D
    g.simplexified = 1 if model.simplexified else 0
    for d in range(model.D):
        g.pmin[d] = model.pmin[d]
        g.pmax[d] = model.pmax[d]
        g.partition[d] = model.partition[d]
    if slab is not None:
        g.slab_begin, g.slab_end = int(slab[0]), int(slab[1])
    return g


def make_leaf_records(leaves: Sequence[LeafFunction]):
    arr = (_lib.LeafRec * max(len(leaves), 1))()
    for i, l in enumerate(leaves):
        arr[i].kind = l.kind
        for d in range(3):
            arr[i].x0[d] = l.x0[d]
            arr[i].v[d] = l.v[d]
        arr[i].R = l.R
        arr[i].r = l.r
    return arr


def _nodal_pointers(ctx, model, nodal, slab=None):
    """Pointers for `gecut_cut`: all NODAL vectors must live on the same side (host or device)."""
    L = len(nodal)
    ptrs = (C.c_void_p * max(L, 1))()
    keep = []
    given = [v for v in nodal if v is not None]
    on_device = bool(given) and all(_is_torch_tensor(v) and v.is_cuda for v in given)
    for i, v in enumerate(nodal):
        if v is None:
            ptrs[i] = None
            continue
        if on_device:
            import torch
            t = v.contiguous()
            if t.dtype != torch.float64:
                raise TypeError("nodal values must be float64")
            keep.append(t)
            ptrs[i] = t.data_ptr()
        else:
            a = v.detach().cpu().numpy() if _is_torch_tensor(v) else v
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            ptrs[i] = a.ctypes.data
    return ptrs, on_device, keep


class LevelSetCutter(Cutter):
    """`struct LevelSetCutter <: Cutter` (`LevelSetCutters.jl:77`), B200 implementation.

    `ctx`: a `_lib.Context` (default: one per LOCAL_RANK device).  `slab=(k0,k1)`: cut only the n-cubes whose last
    index lies in [k0,k1) (multi-GPU z-slab partition)."""

    def __init__(self, ctx: Optional[_lib.Context] = None, slab: Optional[Tuple[int, int]] = None,
                 nodal_slab_local: bool = False):
        self._ctx = ctx
        self.slab = slab
        self.nodal_slab_local = bool(nodal_slab_local)  # NODAL vectors hold only the node planes slab[0]..slab[1]

    @property
    def ctx(self) -> _lib.Context:
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def _call(self, fn_name: str, background, geom):
        from .discretizations import EmbeddedDiscretization
        model = background
        num_nodal, planes = None, None
        if self.nodal_slab_local and self.slab is not None:
            layer_nodes = model.num_nodes // (model.partition[-1] + 1)
            num_nodal = layer_nodes * (self.slab[1] - self.slab[0] + 1)
            planes = (self.slab[0], self.slab[1])
        leaves, nodal, oid_to_ls = flatten_geometry(model, geom, num_nodal=num_nodal, nodal_planes=planes)
        grid = make_grid(model, self.slab)
        grid.nodal_slab_local = 1 if num_nodal is not None else 0
        recs = make_leaf_records(leaves)
        ptrs, on_device, keep = _nodal_pointers(self.ctx, model, nodal)
        out = C.c_void_p()
        fn = getattr(self.ctx._lib, fn_name)
        rc = fn(self.ctx.handle, C.byref(grid), recs, len(leaves), ptrs, 1 if on_device else 0, C.byref(out))
        self.ctx.check(rc, fn_name)
        del keep
        disc = EmbeddedDiscretization(self.ctx, out, model, geom, oid_to_ls)
        disc.slab = tuple(self.slab) if self.slab is not None else None  # z-slab of the model this result covers
        return disc

    def cut(self, background, geom):
        """`cut(cutter::LevelSetCutter, background, geom)` (`LevelSetCutters.jl:79-82`)."""
        return self._call("gecut_cut", background, geom)

    def compute_bgcell_to_inoutcut(self, background, geom) -> np.ndarray:
        """`compute_bgcell_to_inoutcut(::LevelSetCutter, model, geom)` (`LevelSetCutters.jl:154-188`)."""
        from .discretizations import compute_bgcell_to_inoutcut as _bg
        disc = self._call("gecut_classify", background, geom)
        return _bg(disc, geom)


def cut(*args):
    """`cut(background, geom)` (`LevelSetCutters.jl:84-92`) or `cut(cutter, background, geom)` (`Cutters.jl:30-32`)."""
    if len(args) == 2:
        return LevelSetCutter().cut(*args)
    cutter, background, geom = args
    return cutter.cut(background, geom)


def compute_bgcell_to_inoutcut(*args):
    """`compute_bgcell_to_inoutcut(model, geom)` / `(cutter, model, geom)` / `(cutgeo, geo)`."""
    from .discretizations import EmbeddedDiscretization, compute_bgcell_to_inoutcut as _bg
    if len(args) == 2 and isinstance(args[0], EmbeddedDiscretization):
        return _bg(*args)
    if len(args) == 2:
        return LevelSetCutter().compute_bgcell_to_inoutcut(*args)
    cutter, background, geom = args
    return cutter.compute_bgcell_to_inoutcut(background, geom)


def evaluate_leaves_at_nodes(model, payloads, ctx: Optional[_lib.Context] = None, device_output: bool = True,
                             slab: Optional[Tuple[int, int]] = None):
    """`j_to_ls = [fun.(point_to_coords) for fun in j_to_fun]` (`DiscreteGeometries.jl:51`) on the device.
    Returns one vector per payload: CUDA torch tensors (device_output) or numpy arrays.  `slab=(k0,k1)`: only the node
    planes k0..k1 of the last direction (the slab-local vectors of a rank of the z-slab partition)."""
    ctx = ctx or _lib.default_context()
    grid = make_grid(model, slab)
    num = model.num_nodes
    planes = None
    if slab is not None:
        grid.nodal_slab_local = 1
        planes = (int(slab[0]), int(slab[1]))
        num = (model.num_nodes // (model.partition[-1] + 1)) * (planes[1] - planes[0] + 1)
    out = []
    for p in payloads:
        if not isinstance(p, LeafFunction):
            out.append(p)
            continue
        if p.kind == LEAF_NODAL:
            out.append(np.ascontiguousarray(p(model.node_coordinates(planes)), dtype=np.float64))
            continue
        rec = make_leaf_records([p])
        if device_output:
            import torch
            t = torch.empty(num, dtype=torch.float64, device=f"cuda:{ctx.device}")
            rc = ctx._lib.gecut_eval_leaf_at_nodes(ctx.handle, C.byref(grid), rec, C.c_void_p(t.data_ptr()), 1)
            ctx.check(rc, "gecut_eval_leaf_at_nodes")
            out.append(t)
        else:
            a = np.empty(num, dtype=np.float64)
            rc = ctx._lib.gecut_eval_leaf_at_nodes(ctx.handle, C.byref(grid), rec, C.c_void_p(a.ctypes.data), 0)
            ctx.check(rc, "gecut_eval_leaf_at_nodes")
            out.append(a)
    return out
