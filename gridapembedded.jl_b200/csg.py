"""CSG tree of the host-side mirror (reference: `src/CSG/Nodes.jl`, `src/CSG/Geometries.jl`).

A geometry is a binary/unary tree whose nodes carry `(payload, name, meta)`:
leaves hold a level-set payload (an analytical leaf record or a vector of nodal
values), inner nodes hold one of the operators `"∪" "∩" "-" "!"`
(`Geometries.jl:111-125`).  Level-set numbering follows `_find_unique_leaves`
(`src/LevelSetCutters/DiscreteGeometries.jl:65-74`): leaves in left-to-right DFS
order, de-duplicated by the identity of the payload object (Julia `objectid`
== Python `id`), first occurrence wins.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

UNION, INTERSECT, SETDIFF, COMPLEMENT = "∪", "∩", "-", "!"

# postfix-program operator tokens shared with the C ABI (include/gecut.h)
OP_UNION, OP_INTERSECT, OP_SETDIFF, OP_COMPLEMENT = -1, -2, -3, -4
_OP_TOKEN = {UNION: OP_UNION, INTERSECT: OP_INTERSECT, SETDIFF: OP_SETDIFF, COMPLEMENT: OP_COMPLEMENT}


class Node:
    """`Node{Td,Ta,Tb}` (`Nodes.jl:2-20`); `rightchild is None` = UnaryNode, both None = Leaf."""

    __slots__ = ("data", "leftchild", "rightchild")

    def __init__(self, data, leftchild: Optional["Node"] = None, rightchild: Optional["Node"] = None):
        self.data = data
        self.leftchild = leftchild
        self.rightchild = rightchild

    @property
    def is_leaf(self) -> bool:
        return self.leftchild is None and self.rightchild is None

    @property
    def is_unary(self) -> bool:
        return self.leftchild is not None and self.rightchild is None

    def children(self) -> Tuple["Node", ...]:
        if self.is_leaf:
            return ()
        if self.is_unary:
            return (self.leftchild,)
        return (self.leftchild, self.rightchild)

    def __repr__(self):
        if self.is_leaf:
            return f"Leaf({self.data!r})"
        return "Node(" + ",".join([repr(self.data)] + [repr(c) for c in self.children()]) + ")"


def Leaf(data) -> Node:
    return Node(data)


def leaves(tree: Node) -> Iterator[Node]:
    """`AbstractTrees.Leaves(tree)`: left-to-right DFS."""
    if tree.is_leaf:
        yield tree
    else:
        for c in tree.children():
            yield from leaves(c)


def pre_order_dfs(tree: Node) -> Iterator[Node]:
    yield tree
    for c in tree.children():
        yield from pre_order_dfs(c)


def replace_data(filter_node: Callable, filter_leaf: Callable, a: Node) -> Node:
    """`replace_data(filter_node,filter_leaf,tree)` (`Nodes.jl:83-102`)."""
    if a.is_leaf:
        return Leaf(filter_leaf(a.data))
    if a.is_unary:
        return Node(filter_node(a.data), replace_data(filter_node, filter_leaf, a.leftchild))
    return Node(filter_node(a.data), replace_data(filter_node, filter_leaf, a.leftchild),
                replace_data(filter_node, filter_leaf, a.rightchild))


class Geometry:
    """`abstract type Geometry` (`Geometries.jl:14-35`)."""

    def get_tree(self) -> Node:
        raise NotImplementedError("abstract method")

    def similar_geometry(self, tree: Node) -> "Geometry":
        raise NotImplementedError("abstract method")

    def compatible_geometries(self, other: "Geometry"):
        raise NotImplementedError("abstract method")

    # operators (`Geometries.jl:78-109`)
    def __invert__(self):
        return complement(self)

    def __or__(self, other):
        return union(self, other)

    def __and__(self, other):
        return intersect(self, other)

    def __sub__(self, other):
        return setdiff(self, other)


def get_tree(geo: Geometry) -> Node:
    return geo.get_tree()


def get_name(geo: Geometry) -> str:
    return geo.get_tree().data[1]


def get_metadata(geo: Geometry):
    return geo.get_tree().data[2]


def replace_metadata(geo: Geometry, meta) -> Geometry:
    t = geo.get_tree()
    g, name, _ = t.data
    return geo.similar_geometry(Node((g, name, meta), t.leftchild, t.rightchild))


def _binary(op: str, a: Geometry, b: Geometry, name: str, meta) -> Geometry:
    _a, _b = a.compatible_geometries(b)
    tree = Node((op, name, meta), _a.get_tree(), _b.get_tree())
    return _a.similar_geometry(tree)


def union(a: Geometry, b: Geometry, name: str = "", meta=None) -> Geometry:
    """`Base.union(a,b;name,meta)` (`Geometries.jl:78-82,111-113`)."""
    return _binary(UNION, a, b, name, meta)


def intersect(a: Geometry, b: Geometry, name: str = "", meta=None) -> Geometry:
    """`Base.intersect` (`Geometries.jl:87-91,115-117`)."""
    return _binary(INTERSECT, a, b, name, meta)


intersection = intersect


def setdiff(a: Geometry, b: Geometry, name: str = "", meta=None) -> Geometry:
    """`Base.setdiff` (`Geometries.jl:96-100,119-121`)."""
    return _binary(SETDIFF, a, b, name, meta)


def complement(a: Geometry, name: str = "", meta=None) -> Geometry:
    """`Base.:!` (`Geometries.jl:105-108,123-125`)."""
    return a.similar_geometry(Node((COMPLEMENT, name, meta), a.get_tree()))


def get_geometry_node(a: Node, name: str) -> Node:
    """`get_geometry_node` (`Geometries.jl:146-154`): first pre-order node with that name."""
    for ai in pre_order_dfs(a):
        if ai.data[1] == name:
            return ai
    raise KeyError(f"There is no entity called {name}")


def get_geometry(a: Geometry, name: str) -> Geometry:
    """`get_geometry(geo,name)` (`Geometries.jl:128-132`)."""
    return a.similar_geometry(get_geometry_node(a.get_tree(), name))


def get_geometry_names(geo: Geometry) -> List[str]:
    return [n.data[1] for n in pre_order_dfs(geo.get_tree()) if n.data[1] != ""]


class LevelSetMap(dict):
    """`oid_to_ls`: id(payload) -> level set, plus the egal keys (`by_key`) so that a geometry built later from equal
    closed-form leaves (Julia: the same `objectid`) finds its level sets too."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.by_key: Dict[Any, int] = {}

    def level_set_of(self, payload) -> int:
        ls = self.get(id(payload))
        if ls is None:
            ls = self.by_key.get(_egal_key(payload))
        if ls is None:
            raise KeyError("this leaf is not one of the level sets of the discretization")
        return ls


def _egal_key(payload):
    """What Julia's `objectid` sees (`DiscreteGeometries.jl:69`): closures over isbits values (the closed-form leaf functions of
    `AnalyticalGeometries.jl` capture a radius, a centre, a direction) are immutable, so two of them built from the same
    numbers are egal and hash alike -- `sphere(1.0)` twice is ONE level set in the reference.  Closed-form payloads compare
    by kind and parameter bits; everything else (arbitrary callables, value vectors) by identity."""
    kind = getattr(payload, "kind", None)
    if kind is None or getattr(payload, "fun", None) is not None or not hasattr(payload, "x0"):
        return ("id", id(payload))
    import struct
    bits = struct.pack("<9d", *payload.x0, *payload.v, payload.R, payload.r, 0.0)
    return ("leaf", int(kind), int(getattr(payload, "dim", 0)), bits)


def find_unique_leaves(tree: Node):
    """`_find_unique_leaves` (`DiscreteGeometries.jl:65-74`): payloads of the unique
    leaves (first DFS occurrence) and the map `id(payload) -> ls` (0-based here); the ids of ALL payload objects that are
    egal to a unique leaf (see `_egal_key`) map to its level set."""
    j_to_payload: List[Any] = []
    oid_to_j = LevelSetMap()
    key_to_j = oid_to_j.by_key
    for leaf in leaves(tree):
        payload = leaf.data[0]
        oid = id(payload)
        if oid in oid_to_j:
            continue
        key = _egal_key(payload)
        if key not in key_to_j:
            key_to_j[key] = len(j_to_payload)
            j_to_payload.append(payload)
        oid_to_j[oid] = key_to_j[key]
    return j_to_payload, oid_to_j


def compile_postfix(tree: Node, oid_to_ls: Dict[int, int]) -> List[int]:
    """Flatten a (sub)tree into the postfix boolean program consumed by the selector
    kernels: token >= 0 pushes the state row of that level set, negative tokens are
    operators.  Replaces the recursive broadcast of `compute_inoutcut`
    (`src/Interfaces/EmbeddedDiscretizations.jl:107-134`)."""
    out: List[int] = []

    def rec(n: Node):
        if n.is_leaf:
            out.append(oid_to_ls.level_set_of(n.data[0]) if isinstance(oid_to_ls, LevelSetMap) else oid_to_ls[id(n.data[0])])
            return
        for c in n.children():
            rec(c)
        op = n.data[0]
        if op not in _OP_TOKEN:
            raise ValueError(f"operation {op} not implemented")
        if (op == COMPLEMENT) != n.is_unary:
            raise ValueError(f"operation {op} has the wrong arity")
        out.append(_OP_TOKEN[op])

    rec(tree)
    return out
