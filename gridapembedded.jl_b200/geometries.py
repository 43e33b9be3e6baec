"""Analytical and discrete level-set geometries of the host-side mirror.

Reference: `src/LevelSetCutters/AnalyticalGeometries.jl` (factories, leaf formulas,
bounding boxes) and `src/LevelSetCutters/DiscreteGeometries.jl`.

Every closed-form leaf is stored as a `LeafFunction` record `(kind, x0, v, R, r)`
-- the "bytecode" the fused sm_100a classification kernel evaluates.  Parameters
are prepared on the host with exactly the arithmetic the reference constructors
use (e.g. `d = v/norm(v)` for cylinders, `x0 -+ 0.5*L*e` for cube faces) so that
the device formulas see bit-identical constants.  Arbitrary Python callables and
`popcorn` (12 `exp` calls, not bit-reproducible across libm implementations) are
kept as host functions and reach the device as NODAL values.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np

from .csg import Geometry, Leaf, Node, find_unique_leaves, get_geometry, intersect, replace_data, union

# leaf kinds shared with the C ABI (include/gecut.h)
LEAF_SPHERE, LEAF_CYLINDER, LEAF_PLANE, LEAF_DOUGHNUT, LEAF_NODAL = 0, 1, 2, 3, 4


def _vec(x) -> tuple:
    return tuple(x)


def _add(a, b):
    return tuple(ai + bi for ai, bi in zip(a, b))


def _sub(a, b):
    return tuple(ai - bi for ai, bi in zip(a, b))


def _scale(s, a):
    return tuple(s * ai for ai in a)


def _neg(a):
    return tuple(-ai for ai in a)


def _shift(a, s):
    """Point +- scalar (componentwise), as `x0 - e*R` in `_sphere_box`."""
    return tuple(ai + s for ai in a)


def _norm(v):
    s = v[0] * v[0]
    for vi in v[1:]:
        s = s + vi * vi
    return math.sqrt(s)


class BoundingBox:
    """`BoundingBox{D,T}` (`AnalyticalGeometries.jl:41-44`)."""

    def __init__(self, pmin, pmax):
        self.pmin = tuple(float(x) for x in pmin)
        self.pmax = tuple(float(x) for x in pmax)

    def __repr__(self):
        return f"BoundingBox({self.pmin},{self.pmax})"


class LeafFunction:
    """Payload of an analytical leaf.  Identity (`id`) plays the role of Julia's
    `objectid(f)` in `_find_unique_leaves` (`DiscreteGeometries.jl:65-74`)."""

    __slots__ = ("kind", "x0", "v", "R", "r", "fun", "dim")

    def __init__(self, kind: int, dim: int, x0=(0.0, 0.0, 0.0), v=(0.0, 0.0, 0.0), R=0.0, r=0.0,
                 fun: Optional[Callable] = None):
        self.kind = kind
        self.dim = dim
        self.x0 = tuple(float(c) for c in x0) + (0.0,) * (3 - len(x0))
        self.v = tuple(float(c) for c in v) + (0.0,) * (3 - len(v))
        self.R = float(R)
        self.r = float(r)
        self.fun = fun  # host callable x(N,D) -> (N,) for LEAF_NODAL payloads

    def __call__(self, x: np.ndarray) -> np.ndarray:
        """Host evaluation (numpy, same expression order as the reference formulas).
        Used for documentation/tests of the mirror only -- `cut` evaluates analytical
        leaves on the device."""
        x = np.asarray(x, dtype=np.float64)
        D = x.shape[1]
        if self.kind == LEAF_NODAL:
            return np.asarray(self.fun(x), dtype=np.float64)
        w = [x[:, d] - self.x0[d] for d in range(D)]

        def dot(a, b):
            s = a[0] * b[0]
            for d in range(1, D):
                s = s + a[d] * b[d]
            return s

        if self.kind == LEAF_SPHERE:
            return dot(w, w) - self.R * self.R
        if self.kind == LEAF_CYLINDER:
            A = dot(w, self.v)
            H2 = dot(w, w)
            B = H2 - A * A
            return B - self.R * self.R
        if self.kind == LEAF_PLANE:
            return dot(w, self.v)
        if self.kind == LEAF_DOUGHNUT:
            t = self.R - np.sqrt(w[0] * w[0] + w[1] * w[1])
            return (t * t + w[2] * w[2]) - self.r * self.r
        raise ValueError("unknown leaf kind")


class AnalyticalGeometry(Geometry):
    """`struct AnalyticalGeometry <: CSG.Geometry` (`AnalyticalGeometries.jl:31-49`).

    `AnalyticalGeometry(f, name=..., dim=...)` wraps a host callable `f(x[N,D]) -> [N]`
    (negative inside)."""

    def __init__(self, tree_or_fun, name: Optional[str] = None, dim: int = 3):
        if isinstance(tree_or_fun, Node):
            self.tree = tree_or_fun
        else:
            f = tree_or_fun
            nm = name if name is not None else getattr(f, "__name__", "f")
            self.tree = Leaf((LeafFunction(LEAF_NODAL, dim, fun=f), nm, None))

    def get_tree(self):
        return self.tree

    def similar_geometry(self, tree):
        return AnalyticalGeometry(tree)

    def compatible_geometries(self, other):
        if isinstance(other, AnalyticalGeometry):
            return self, other
        if isinstance(other, DiscreteGeometry):
            return discretize(self, other.model), other
        raise TypeError("incompatible geometries")


# ---- factories ---------------------------------------------------------------------------

def doughnut(R, r, x0=(0.0, 0.0, 0.0), name="doughnut"):
    """`doughnut` + `_doughnut_box` (`AnalyticalGeometries.jl:53-78`)."""
    m = 0.1 * r
    A = (R + r) + m
    B = r + m
    box = BoundingBox(_add((-A, -A, -B), x0), _add((A, A, B), x0))
    return AnalyticalGeometry(Leaf((LeafFunction(LEAF_DOUGHNUT, 3, x0=x0, R=R, r=r), name, box)))


def _popcorn_fun(x, x0, r0, sigma, A):
    """`_popcorn_fun` (`AnalyticalGeometries.jl:105-126`), host-evaluated."""
    def point_k(k):
        if 0 <= k <= 4:
            a = 2 * k * math.pi / 5
            s = r0 / math.sqrt(5)
            return (s * (2 * math.cos(a)), s * (2 * math.sin(a)), s * 1.0)
        if 5 <= k <= 9:
            a = (2 * (k - 5) - 1) * math.pi / 5
            s = r0 / math.sqrt(5)
            return (s * (2 * math.cos(a)), s * (2 * math.sin(a)), s * -1.0)
        if k == 10:
            return (0.0, 0.0, r0)
        return (0.0, 0.0, -r0)

    X = x[:, 0] - x0[0]
    Y = x[:, 1] - x0[1]
    Z = x[:, 2] - x0[2]
    val = np.sqrt(X * X + Y * Y + Z * Z) - r0
    for k in range(12):
        xk, yk, zk = point_k(k)
        val = val - A * np.exp(-((X - xk) ** 2 + (Y - yk) ** 2 + (Z - zk) ** 2) / sigma ** 2)
    return val


def popcorn(r0=0.6, sigma=0.2, A=2, x0=(0.0, 0.0, 0.0), name="popcorn"):
    """`popcorn` (`AnalyticalGeometries.jl:80-103`).  Host-evaluated (NODAL on the device)."""
    e = 1.5
    box = BoundingBox(_shift(x0, -(e * r0)), _shift(x0, e * r0))
    lf = LeafFunction(LEAF_NODAL, 3, fun=lambda x: _popcorn_fun(x, x0, r0, sigma, A))
    return AnalyticalGeometry(Leaf((lf, name, box)))


def _sphere_box(R, x0):
    """`_sphere_box` (`AnalyticalGeometries.jl:139-144`)."""
    e = 1.01
    return BoundingBox(_shift(x0, -(e * R)), _shift(x0, e * R))


def sphere(R, x0=(0, 0, 0), name="sphere"):
    """`sphere` (`AnalyticalGeometries.jl:128-137`); leaf `_sphere` (`:146-150`)."""
    return AnalyticalGeometry(Leaf((LeafFunction(LEAF_SPHERE, 3, x0=x0, R=R), name, _sphere_box(R, x0))))


def disk(R, x0=(0, 0), name="disk"):
    """`disk` (`AnalyticalGeometries.jl:152-161`): `_sphere` in 2-D."""
    return AnalyticalGeometry(Leaf((LeafFunction(LEAF_SPHERE, 2, x0=x0, R=R), name, _sphere_box(R, x0))))


def cylinder(R, x0=(0, 0, 0), v=(1, 0, 0), name="cylinder"):
    """`cylinder` (`AnalyticalGeometries.jl:163-174`): axis normalised as `d = v/norm(v)`."""
    nv = _norm(v)
    d = tuple(vi / nv for vi in v)
    return AnalyticalGeometry(Leaf((LeafFunction(LEAF_CYLINDER, 3, x0=x0, v=d, R=R), name, None)))


def plane(x0=(0, 0, 0), v=(1, 0, 0), name="plane"):
    """`plane` (`AnalyticalGeometries.jl:184-198`): `(x-x0).v`, `v` NOT normalised."""
    return AnalyticalGeometry(Leaf((LeafFunction(LEAF_PLANE, len(v), x0=x0, v=v), name, None)))


def square(L=1, x0=(0, 0), name="square", edges=None):
    """`square` (`AnalyticalGeometries.jl:200-215`)."""
    edges = edges or [f"edge{i}" for i in range(1, 5)]
    e1, e2 = (1, 0), (0, 1)
    plane1 = plane(x0=_sub(x0, _scale(0.5 * L, e2)), v=_neg(e2), name=edges[0])
    plane2 = plane(x0=_add(x0, _scale(0.5 * L, e2)), v=e2, name=edges[1])
    plane3 = plane(x0=_sub(x0, _scale(0.5 * L, e1)), v=_neg(e1), name=edges[2])
    plane4 = plane(x0=_add(x0, _scale(0.5 * L, e1)), v=e1, name=edges[3])
    geo12 = intersect(plane1, plane2)
    geo34 = intersect(plane3, plane4)
    return intersect(geo12, geo34, name=name)


def quadrilateral(x0=(0, 0), d1=(1, 0), d2=(0, 1), name="quadrilateral"):
    """`quadrilateral` (`AnalyticalGeometries.jl:217-260`), IEEE inf/nan semantics kept."""
    f = np.float64
    with np.errstate(divide="ignore", invalid="ignore"):
        x1 = _add(x0, d1)
        x2 = _add(x0, d2)
        slope1 = f(d1[1]) / f(d1[0])
        slope2 = f(d2[1]) / f(d2[0])
        if slope1 > slope2:
            slope1, slope2 = slope2, slope1
            x1, x2 = x2, x1
        slope_n1 = f(-1.0) / slope1
        slope_n2 = f(-1.0) / slope2
        den1 = np.sqrt(f(1) + slope_n1 * slope_n1)
        den2 = np.sqrt(f(1) + slope_n2 * slope_n2)
        n1 = (float(f(1) / den1), float(slope_n1 / den1))
        n2 = (float(f(1) / den2), float(slope_n2 / den2))
        if slope_n1 == -np.inf:
            n1 = (0.0, -1.0)
        if slope_n2 == -np.inf:
            n2 = (0.0, -1.0)
    plane1 = plane(x0=x0, v=n1, name="edge1")
    plane2 = plane(x0=x2, v=_neg(n1), name="edge2")
    plane3 = plane(x0=x0, v=_neg(n2), name="edge3")
    plane4 = plane(x0=x1, v=n2, name="edge4")
    geo12 = intersect(plane1, plane2)
    geo34 = intersect(plane3, plane4)
    return intersect(geo12, geo34, name=name)


def cube(L=1, x0=(0, 0, 0), name="cube"):
    """`cube` (`AnalyticalGeometries.jl:262-281`): six planes, DFS order face1..face6."""
    e1, e2, e3 = (1, 0, 0), (0, 1, 0), (0, 0, 1)
    plane1 = plane(x0=_sub(x0, _scale(0.5 * L, e3)), v=_neg(e3), name="face1")
    plane2 = plane(x0=_add(x0, _scale(0.5 * L, e3)), v=e3, name="face2")
    plane3 = plane(x0=_sub(x0, _scale(0.5 * L, e2)), v=_neg(e2), name="face3")
    plane4 = plane(x0=_add(x0, _scale(0.5 * L, e2)), v=e2, name="face4")
    plane5 = plane(x0=_sub(x0, _scale(0.5 * L, e1)), v=_neg(e1), name="face5")
    plane6 = plane(x0=_add(x0, _scale(0.5 * L, e1)), v=e1, name="face6")
    geo12 = intersect(plane1, plane2)
    geo34 = intersect(plane3, plane4)
    geo56 = intersect(plane5, plane6)
    return intersect(intersect(geo12, geo34), geo56, name=name)


def tube(R, L, x0=(0.0, 0.0, 0.0), v=(1, 0, 0), name="tube"):
    """`tube` (`AnalyticalGeometries.jl:283-299`): inlet ∩ outlet ∩ walls."""
    nv = _norm(v)
    d = tuple(vi / nv for vi in v)
    box = BoundingBox(_shift(x0, -R), _shift(_add(x0, _scale(L, d)), R))
    walls = cylinder(R, x0=x0, v=d, name="walls")
    inlet = plane(x0=x0, v=_neg(d), name="inlet")
    outlet = plane(x0=_add(x0, _scale(L, d)), v=d, name="outlet")
    return intersect(intersect(inlet, outlet), walls, name=name, meta=box)


def olympic_rings(R, r, name="olympic_rings"):
    """`olympic_rings` (`AnalyticalGeometries.jl:301-329`)."""
    m = 0.1 * r
    A = 2 * (R + r) + r + m
    B = (R + r) + m
    C = r + m
    box = BoundingBox((-A, -B - (R - r), -C), (A + 2 * (r + r + R), B, C))
    z = 0.0 * R
    geo1 = doughnut(R, r, name="ring1", x0=(-(r + r + R), z, z))
    geo2 = doughnut(R, r, name="ring2", x0=(r + r + R, z, z))
    geo3 = doughnut(R, r, name="ring3", x0=(z, -R + r, z))
    geo4 = doughnut(R, r, name="ring4", x0=(2 * (r + r + R), -R + r, z))
    geo5 = doughnut(R, r, name="ring5", x0=(3 * (r + r + R), z, z))
    geo12 = union(geo1, geo2)
    geo123 = union(geo12, geo3)
    geo1234 = union(geo123, geo4)
    return union(geo1234, geo5, name=name, meta=box)


# ---- discrete geometries -------------------------------------------------------------------

class DiscreteGeometry(Geometry):
    """`struct DiscreteGeometry{D,T} <: CSG.Geometry` (`DiscreteGeometries.jl:14-17`).

    Leaves hold vectors of nodal values (host numpy arrays or CUDA torch tensors), one
    value per background node (lexicographic, x fastest) -- the reference takes the free
    dofs of a first-order FE function verbatim (`DiscreteGeometries.jl:88-98`).

    Constructors: `DiscreteGeometry(point_to_value, model, name="")` or
    `DiscreteGeometry(tree, model)`.
    """

    def __init__(self, values_or_tree, model, name: str = ""):
        self.model = model
        if isinstance(values_or_tree, Node):
            self.tree = values_or_tree
        elif callable(values_or_tree):
            # DiscreteGeometry(f::Function,model) (DiscreteGeometries.jl:76-79)
            geo = discretize(AnalyticalGeometry(values_or_tree, name=name, dim=model.D), model)
            self.tree = geo.tree
        else:
            vals = values_or_tree
            if not hasattr(vals, "shape"):
                vals = np.asarray(vals, dtype=np.float64)
            # the length is checked at cut time (whole model, or one z-slab of it for a distributed cut)
            self.tree = Leaf((vals, name, None))

    def get_tree(self):
        return self.tree

    def similar_geometry(self, tree):
        return DiscreteGeometry(tree, self.model)

    def compatible_geometries(self, other):
        if isinstance(other, DiscreteGeometry):
            if not self.model.same_nodes(other.model):
                raise ValueError("discrete geometries live on different node sets")
            return self, other
        if isinstance(other, AnalyticalGeometry):
            return self, discretize(other, self.model)
        raise TypeError("incompatible geometries")


def discretize(a: AnalyticalGeometry, model) -> DiscreteGeometry:
    """`discretize(a::AnalyticalGeometry, model)` (`DiscreteGeometries.jl:36-63`): every unique
    leaf evaluated at every node; the tree keeps its shape with value vectors as payloads.
    Closed-form leaves are evaluated by the device node-evaluation kernel."""
    from .cutters import evaluate_leaves_at_nodes  # device path
    tree = a.get_tree()
    j_to_fun, oid_to_j = find_unique_leaves(tree)
    j_to_ls = evaluate_leaves_at_nodes(model, j_to_fun)

    def conversion(data):
        f, name, meta = data
        return (j_to_ls[oid_to_j[id(f)]], name, meta)

    return DiscreteGeometry(replace_data(lambda d: d, conversion, tree), model)
