"""Host-side mirror of the reference API (CSG trees, geometry factories, level-set numbering, postfix programs) and the
C-ABI library surface.  No GPU needed: nothing here computes on the device."""
import ctypes
import os
import re

import numpy as np
import pytest

import gecut
from gecut import (CartesianDiscreteModel, simplexify, sphere, disk, cube, cylinder, plane, square, quadrilateral, tube,
                   doughnut, olympic_rings, union, intersect, setdiff, complement, get_geometry, get_geometry_names)
from gecut.csg import (Node, Leaf, find_unique_leaves, compile_postfix, leaves, replace_data, OP_UNION, OP_INTERSECT,
                       OP_SETDIFF, OP_COMPLEMENT)
from gecut import _lib
from gecut.geometries import LEAF_PLANE, LEAF_SPHERE, LEAF_CYLINDER

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_csg_nodes():
    # test/CSGTests/NodesTests.jl:19-45
    a, b = Leaf(("a", "A", None)), Leaf(("b", "B", None))
    n = Node((":∪", "u", None), a, b)
    assert n.leftchild is a and n.rightchild is b and not n.is_leaf and a.is_leaf
    u = Node(("!", "n", None), a)
    assert u.is_unary and u.children() == (a,)
    t = replace_data(lambda d: d, lambda d: (d[0].upper(), d[1], d[2]), n)
    assert [l.data[0] for l in leaves(t)] == ["A", "B"]


def test_csg_geometry_names_and_lookup():
    # test/CSGTests/GeometriesTests.jl:8-54
    g1, g2, g3 = sphere(1, name="s1"), sphere(2, name="s2"), sphere(3, name="s3")
    g = setdiff(union(g1, g2, name="u"), intersect(g2, complement(g3, name="c"), name="i"), name="all")
    assert get_geometry_names(g) == ["all", "u", "s1", "s2", "i", "s2", "c", "s3"]
    assert gecut.get_name(get_geometry(g, "i")) == "i"
    assert get_geometry(g, "s2").get_tree() is g.get_tree().leftchild.rightchild
    with pytest.raises(KeyError):
        get_geometry(g, "missing")
    # operators
    h = (g1 | g2) & ~g3
    assert h.get_tree().data[0] == "∩" and h.get_tree().rightchild.data[0] == "!"


def test_unique_leaves_numbering():
    # _find_unique_leaves (DiscreteGeometries.jl:65-74): DFS order, de-duplicated by payload identity
    R, L = 0.3, 2.0
    geo2 = tube(R, L, x0=(0.0, 0.0, 0.0))
    geo3 = sphere(0.3 * R, x0=(L / 3, 0.0, 0.0))
    geo4 = complement(get_geometry(geo2, "walls"))
    geo5 = union(geo3, geo4, name="solid")
    geo1 = setdiff(geo2, geo3, name="fluid")
    g = union(geo1, geo5)
    payloads, oid = find_unique_leaves(g.get_tree())
    assert len(payloads) == 4  # inlet, outlet, walls, sphere; `!walls` and the second sphere re-use payloads
    kinds = [p.kind for p in payloads]
    assert kinds == [LEAF_PLANE, LEAF_PLANE, LEAF_CYLINDER, LEAF_SPHERE]
    prog = compile_postfix(g.get_tree(), oid)
    assert prog == [0, 1, OP_INTERSECT, 2, OP_INTERSECT, 3, OP_SETDIFF, 3, 2, OP_COMPLEMENT, OP_UNION, OP_UNION]
    assert compile_postfix(get_geometry(g, "solid").get_tree(), oid) == [3, 2, OP_COMPLEMENT, OP_UNION]
    # two spheres with equal parameters are ONE level set: Julia closures over isbits values are egal, `objectid` agrees
    # (ADVICE round 1); see test_unique_leaves_follow_julia_objectid_for_closed_forms
    p2, _ = find_unique_leaves(union(sphere(1), sphere(1)).get_tree())
    assert len(p2) == 1


def test_cube_and_square_planes():
    # cube (AnalyticalGeometries.jl:262-281): face1..face6 = z-,z+,y-,y+,x-,x+
    g = cube(L=1.5)
    payloads, _ = find_unique_leaves(g.get_tree())
    assert [l.data[1] for l in leaves(g.get_tree())] == [f"face{i}" for i in range(1, 7)]
    x0s = [p.x0 for p in payloads]
    vs = [p.v for p in payloads]
    assert x0s == [(0, 0, -0.75), (0, 0, 0.75), (0, -0.75, 0), (0, 0.75, 0), (-0.75, 0, 0), (0.75, 0, 0)]
    assert vs == [(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)]
    s = square(L=2, x0=(1, 1))
    ps, _ = find_unique_leaves(s.get_tree())
    assert [p.x0[:2] for p in ps] == [(1, 0), (1, 2), (0, 1), (2, 1)]
    assert gecut.get_name(s) == "square"


def test_factories_parameters():
    c = cylinder(0.5, v=(0, 3, 4)).get_tree().data[0]
    assert c.v == (0.0, 3 / 5.0, 4 / 5.0)
    b = gecut.get_metadata(sphere(0.7))
    assert b.pmin == (-(1.01 * 0.7),) * 3 and b.pmax == (1.01 * 0.7,) * 3
    t = tube(0.3, 2.0)
    names = [l.data[1] for l in leaves(t.get_tree())]
    assert names == ["inlet", "outlet", "walls"]
    inlet, outlet, walls = [l.data[0] for l in leaves(t.get_tree())]
    assert inlet.v == (-1.0, -0.0, -0.0) and outlet.x0 == (2.0, 0.0, 0.0) and walls.kind == LEAF_CYLINDER
    q = quadrilateral(x0=(0, 0), d1=(1, 0), d2=(0, 1))
    ps, _ = find_unique_leaves(q.get_tree())
    assert ps[0].v[:2] == (0.0, -1.0) and ps[1].v[:2] == (-0.0, 1.0)  # slope_n == -Inf branch
    o = olympic_rings(0.6, 0.1)
    assert len(find_unique_leaves(o.get_tree())[0]) == 5 and gecut.get_metadata(o) is not None
    d = doughnut(1.2, 0.4, x0=(1.0, 0.0, 0.0))
    assert gecut.get_metadata(d).pmax == ((1.2 + 0.4) + 0.1 * 0.4 + 1.0, (1.2 + 0.4) + 0.1 * 0.4, 0.4 + 0.1 * 0.4)


def test_leaf_host_evaluation_matches_oracle():
    import oracle_binding as ob
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1000, 3))
    for geo in (sphere(0.7, x0=(0.1, 0.2, 0.3)), cylinder(0.4, x0=(0, 1, 0.7), v=(1, 2, 3)), plane(x0=(0.1, 0, 0), v=(1, 2, -1)),
                doughnut(1.2, 0.4)):
        lf = geo.get_tree().data[0]
        out = np.zeros(len(x))
        leaf = ob.make_leaf(lf.kind, lf.x0, lf.v, lf.R, lf.r)
        ob.lib().gecut_oracle_eval_leaf(ctypes.byref(leaf), 3, x.ctypes.data, len(x), out.ctypes.data)
        assert np.array_equal(lf(x), out)


def test_cartesian_model_conventions():
    m = CartesianDiscreteModel((0, 0, 0), (1, 2, 3), (2, 3, 4))
    assert (m.num_cells, m.num_nodes, m.num_vertices_per_cell) == (24, 60, 8)
    X = m.node_coordinates()
    assert X.shape == (60, 3) and np.array_equal(X[1], [0.5, 0, 0]) and np.array_equal(X[3], [0, 2 / 3, 0])
    cn = m.cell_node_ids()
    assert cn[0].tolist() == [1, 2, 4, 5, 13, 14, 16, 17]
    t = simplexify(m)
    assert (t.num_cells, t.num_vertices_per_cell) == (144, 4)
    assert t.cell_node_ids()[:2].tolist() == [[1, 2, 4, 16], [1, 2, 13, 16]]  # simplexify(HEX)[1:2] on cell 1
    import oracle_binding as ob
    assert np.array_equal(ob.node_coords(m), X)
    m2 = CartesianDiscreteModel((-1, 2, 0, 3), (4, 6))
    assert m2.pmin == (-1.0, 0.0) and m2.pmax == (2.0, 3.0)
    with pytest.raises(ValueError):
        CartesianDiscreteModel((0, 0), (1, 1), (0, 4))


def test_flatten_geometry_and_errors():
    m = CartesianDiscreteModel((0, 0), (1, 1), (4, 4))
    with pytest.raises(ValueError):
        gecut.cutters.flatten_geometry(m, sphere(1))  # 3-D leaf on 2-D model
    f = gecut.AnalyticalGeometry(lambda x: x[:, 0] - 0.5, name="half", dim=2)
    leaves_, nodal, oid = gecut.cutters.flatten_geometry(m, union(disk(0.3), f))
    assert [l.kind for l in leaves_] == [LEAF_SPHERE, 4] and nodal[0] is None and nodal[1].shape == (25,)
    dg = gecut.DiscreteGeometry(np.zeros(24), m)
    with pytest.raises(ValueError):
        gecut.cutters.flatten_geometry(m, dg)


def test_c_abi_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports every symbol include/gecut.h declares (no compute call without a GPU)."""
    path = _lib.build_library()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "gecut.h")).read()
    declared = set(re.findall(r"\b(gecut_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    # struct layouts agree with the header
    assert ctypes.sizeof(_lib.Grid) == 8 + 48 + 12 + 12 and ctypes.sizeof(_lib.LeafRec) == 8 + 48 + 16
    assert ctypes.sizeof(_lib.Sizes) == 9 * 8 + 4 * 4  # incl. subfacet_points_alias and ctypes.sizeof(_lib.Buffers) == 13 * 8 + 3 * 8


def test_julia_shim_binds_only_exported_symbols():
    """The reference-side binding (julia/B200LevelSetCutter.jl, delivered as source: there is no Julia here) `ccall`s only
    entry points the library really exports, and the statuses / leaf kinds it hard-codes are the header's."""
    src = open(os.path.join(ROOT, "julia", "B200LevelSetCutter.jl")).read()
    used = set(re.findall(r":(gecut_[a-z0-9_]+)", src))
    assert len(used) >= 10 and used <= set(_lib.EXPORTED_SYMBOLS), used - set(_lib.EXPORTED_SYMBOLS)
    assert {"gecut_create", "gecut_cut", "gecut_classify", "gecut_result_sizes", "gecut_result_copy",
            "gecut_cut_facets"} <= used
    header = open(os.path.join(ROOT, "include", "gecut.h")).read()
    for name in re.findall(r"\b(GECUT_LEAF_[A-Z]+)\b", src):
        assert re.search(r"#define\s+%s\b" % name, header) or re.search(r"\b%s\b" % name, header), name


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.GecutError):
        gecut.cut(CartesianDiscreteModel((0, 0), (1, 1), (4, 4)), disk(0.3, x0=(0.5, 0.5)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gridapembedded.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_binding" not in txt and "gecut_oracle" not in txt, f


def test_unique_leaves_follow_julia_objectid_for_closed_forms():
    """Julia's `objectid` makes two closures over the same isbits values ONE level set
    (`DiscreteGeometries.jl:65-74`): `sphere(1.0)` built twice is deduplicated, different parameters or host callables are not."""
    from gecut import sphere, cylinder, union, intersect, AnalyticalGeometry
    from gecut.csg import find_unique_leaves, compile_postfix
    a, b, c = sphere(1.0), sphere(1.0), sphere(1.0, x0=(0.1, 0.0, 0.0))
    payloads, oid_to_j = find_unique_leaves(union(intersect(a, c), b).get_tree())
    assert len(payloads) == 2
    assert oid_to_j[id(a.get_tree().data[0])] == oid_to_j[id(b.get_tree().data[0])] == 0
    assert oid_to_j[id(c.get_tree().data[0])] == 1
    # a geometry built LATER from an equal closed form finds the level set too (same objectid in Julia)
    assert compile_postfix(sphere(1.0).get_tree(), oid_to_j) == [0]
    f = lambda x: x[:, 0]  # noqa: E731
    g1, g2 = AnalyticalGeometry(f, dim=3), AnalyticalGeometry(lambda x: x[:, 0], dim=3)
    payloads, _ = find_unique_leaves(union(g1, g2).get_tree())
    assert len(payloads) == 2  # arbitrary callables: identity
    assert len(find_unique_leaves(union(cylinder(0.3), cylinder(0.3, v=(0, 1, 0))).get_tree())[0]) == 2
