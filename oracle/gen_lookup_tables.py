#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE -- not part of the shipped product path.

CPU restatement of the reference's run-time marching-simplex table builder
(`src/LevelSetCutters/LookupTables.jl:46-79` `_build_lookup_table`), used to
produce the golden table dump `tests/golden/lookup_tables.json`.

The reference builds the tables with two un-vendored third-party pieces
(SURVEY.md section 8c):
  * MiniQhull (compat 0.1-0.4, qhull 2020.2): `delaunay(points)` with flags
    "qhull d Qt Qbb Qc Qz" (`LookupTables.jl:156-166`).  Restated here with
    scipy's bundled qhull_r 8.0.2 (= qhull 2020.2) and the same flags.
  * Gridap 0.20.x: polytope vertex/edge/facet numbering, the mini-model facet
    numbering and `InterfaceTriangulation` ordering (`LookupTables.jl:234-260`).
    Restated from the published conventions (assumptions A1-A3, A6 below).

PARITY STATUS: "parity unpinned" for subcell/subfacet connectivity inside a
case -- the reference tests hold no literal vector for them (only the
`compute_case` known answers, `test/LevelSetCuttersTests/LookupTablesTests.jl:8-11`).
Everything that IS pinned by the reference tests (volumes, orientation through
the divergence theorem, ...) is asserted in tests/.  If a Julia dump of
`LookupTable(TET)` etc. ever becomes available, drop it in place of the JSON
this script writes; both the oracle and the CUDA library consume the JSON.

Usage:  python oracle/gen_lookup_tables.py [out.json]
"""
import itertools
import json
import sys

import numpy as np
from scipy.spatial import ConvexHull, Delaunay

IN, OUT, CUT, INTERFACE = -1, 1, 0, 0

# ---- Gridap conventions (assumptions A1-A3 of SURVEY.md section 8c) ----------
# A1: n-cube vertices lexicographic x-fastest; simplex = origin then unit axes.
# A3: edge lists `get_faces(p,1,0)` and facet lists `get_faces(p,D-1,0)`.
POLYTOPES = {
    "SEGMENT": dict(
        dim=1,
        vertices=[[0.0], [1.0]],
        edges=[[1, 2]],
        simplex_facets=[[1], [2]],
    ),
    "TRI": dict(
        dim=2,
        vertices=[[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]],
        edges=[[1, 2], [1, 3], [2, 3]],
        simplex_facets=[[1, 2], [1, 3], [2, 3]],
    ),
    "TET": dict(
        dim=3,
        vertices=[[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
        edges=[[1, 2], [1, 3], [2, 3], [1, 4], [2, 4], [3, 4]],
        simplex_facets=[[1, 2, 3], [1, 2, 4], [1, 3, 4], [2, 3, 4]],
    ),
    "QUAD": dict(
        dim=2,
        vertices=[[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]],
        edges=[[1, 2], [3, 4], [1, 3], [2, 4]],
        # A2: simplexify(QUAD)
        simplexify=[[1, 2, 3], [2, 3, 4]],
    ),
    "HEX": dict(
        dim=3,
        vertices=[[float(i), float(j), float(k)] for k in (0, 1) for j in (0, 1) for i in (0, 1)],
        edges=[[1, 2], [3, 4], [5, 6], [7, 8], [1, 3], [2, 4], [5, 7], [6, 8],
               [1, 5], [2, 6], [3, 7], [4, 8]],
        # A2: simplexify(HEX)
        simplexify=[[1, 2, 3, 7], [1, 2, 5, 7], [2, 3, 4, 7], [2, 4, 7, 8], [2, 5, 6, 7], [2, 6, 7, 8]],
    ),
}
POLYTOPES["TRI"]["simplexify"] = [[1, 2, 3]]
POLYTOPES["TET"]["simplexify"] = [[1, 2, 3, 4]]
POLYTOPES["SEGMENT"]["simplexify"] = [[1, 2]]


def isout(v):
    """LookupTables.jl:40-42 (v == 0 counts as IN)."""
    return v > 0


def compute_case(values):
    """LookupTables.jl:19-38."""
    case = 1
    for i, v in enumerate(values):
        if isout(v):
            case += 2 ** i
    return case


def _shapefun_grads(D):
    """Gradients of the P1 shape functions of the D-simplex:
    grad N_1 = (-1,...,-1), grad N_{i+1} = e_i (LookupTables.jl:168-179)."""
    g = [-np.ones(D)]
    for i in range(D):
        e = np.zeros(D)
        e[i] = 1.0
        g.append(e)
    return g


def ensure_positive_jacobians(subcell_to_points, point_to_coords, D):
    """LookupTables.jl:181-197: swap the first two ids when det(J) < 0,
    J = sum_i outer(grad N_i, x_i)."""
    grads = _shapefun_grads(D)
    for pts in subcell_to_points:
        J = np.zeros((D, D))
        for i, p in enumerate(pts):
            J += np.outer(grads[i], np.asarray(point_to_coords[p - 1], float))
        if np.linalg.det(J) < 0:
            pts[0], pts[1] = pts[1], pts[0]


def compute_delaunay_points(vertex_to_value, vertex_to_coords, edge_to_vertices):
    """LookupTables.jl:139-154: vertices, then one midpoint per sign-changing
    edge in edge order, midpoint value = INTERFACE."""
    coords = [list(c) for c in vertex_to_coords]
    values = list(vertex_to_value)
    for a, b in edge_to_vertices:
        if isout(vertex_to_value[a - 1]) != isout(vertex_to_value[b - 1]):
            p = 0.5 * (np.asarray(vertex_to_coords[a - 1]) + np.asarray(vertex_to_coords[b - 1]))
            coords.append(p.tolist())
            values.append(INTERFACE)
    return coords, values


def delaunay(point_to_coords):
    """LookupTables.jl:156-166 -> MiniQhull.delaunay (assumption A7): qhull
    flags 'd Qt Qbb Qc Qz', cells in qhull facet-list order, vertices in
    facet->vertices order, 1-based.

    1-D input cannot go through scipy (it insists on >=2-D data although qhull
    itself lifts to 2-D); it is emulated with the lower convex hull of the
    lifted points (x, x^2) (assumption A7b: only the ORDER of the two children
    of a cut segment depends on it)."""
    P = np.asarray(point_to_coords, float)
    if P.shape[1] == 1:
        if len(P) == 2:
            return [[1, 2]]
        lifted = np.column_stack([P[:, 0], P[:, 0] ** 2])
        hull = ConvexHull(lifted, qhull_options="Qt Qbb Qc")
        cells = []
        for simplex, eq in zip(hull.simplices, hull.equations):
            if eq[1] < 0:  # lower hull = Delaunay
                cells.append([int(i) + 1 for i in simplex])
        return cells
    d = Delaunay(P, qhull_options="Qt Qbb Qc Qz")
    return [[int(i) + 1 for i in s] for s in d.simplices]


def compute_subcell_to_inout(subcell_to_points, point_to_value):
    """LookupTables.jl:218-232."""
    out = []
    for pts in subcell_to_points:
        s = IN
        for p in pts:
            if isout(point_to_value[p - 1]):
                s = OUT
                break
        out.append(s)
    return out


def _unit_normal_from_points(pts):
    """LookupTables.jl:304-372 restricted to D<=3 (orientation-carrying
    normal defined by the point ordering)."""
    D = len(pts[0])
    if D == 1:
        return np.array([1.0])
    if D == 2:
        v = np.asarray(pts[1]) - np.asarray(pts[0])
        w = np.array([v[1], -v[0]])
    else:
        v1 = np.asarray(pts[1]) - np.asarray(pts[0])
        v2 = np.asarray(pts[2]) - np.asarray(pts[0])
        w = np.cross(v1, v2)
    m = np.sqrt(w @ w)
    if m < np.finfo(float).eps:
        return np.zeros(D)
    return w / m


def find_subfacets(point_to_coords, subcell_to_points, subcell_to_inout, D, lfacet_to_lpoints):
    """LookupTables.jl:234-274.  Restates Gridap's mini-model topology
    (assumption A6): facet ids by first encounter over (cell, local facet);
    `InterfaceTriangulation` keeps, in ascending facet id, the facets with one
    IN and one OUT neighbour; its plus side is the IN cell, so the reference
    normal is the outward normal of the IN subcell."""
    facet_ids = {}
    facet_cells = []
    for c, pts in enumerate(subcell_to_points):
        for lf, lpts in enumerate(lfacet_to_lpoints):
            key = tuple(sorted(pts[i - 1] for i in lpts))
            if key not in facet_ids:
                facet_ids[key] = len(facet_cells)
                facet_cells.append([])
            facet_cells[facet_ids[key]].append((c, lf))
    subfacet_to_points = []
    subfacet_to_normal = []
    for cells in facet_cells:
        if len(cells) != 2:
            continue
        (c1, lf1), (c2, lf2) = cells
        s1, s2 = subcell_to_inout[c1], subcell_to_inout[c2]
        if s1 == s2:
            continue
        c, lf = (c1, lf1) if s1 == IN else (c2, lf2)
        pts = subcell_to_points[c]
        fpts = [pts[i - 1] for i in lfacet_to_lpoints[lf]]
        subfacet_to_points.append(fpts)
        # outward normal of subcell c on that facet
        X = [np.asarray(point_to_coords[p - 1], float) for p in fpts]
        opp = [p for p in pts if p not in fpts][0]
        xo = np.asarray(point_to_coords[opp - 1], float)
        if D == 1:
            n = np.array([1.0]) if X[0][0] > xo[0] else np.array([-1.0])
        else:
            n = _unit_normal_from_points(X)
            if n @ (xo - X[0]) > 0:
                n = -n
        subfacet_to_normal.append(n.tolist())
    return subfacet_to_points, subfacet_to_normal


def setup_subfacet_to_orientation(point_to_coords, subfacet_to_points, subfacet_to_normal):
    """LookupTables.jl:276-287: sign(v . n), v from the point ordering."""
    out = []
    for fpts, n in zip(subfacet_to_points, subfacet_to_normal):
        X = [np.asarray(point_to_coords[p - 1], float) for p in fpts]
        v = _unit_normal_from_points(X)
        out.append(float(np.sign(v @ np.asarray(n))))
    return out


def find_in_out_or_cut(vertex_to_value):
    """LookupTables.jl:289-302."""
    b = isout(vertex_to_value[0])
    for v in vertex_to_value:
        if b != isout(v):
            return CUT
    return OUT if b else IN


def build_lookup_table(name):
    """LookupTables.jl:46-79."""
    p = POLYTOPES[name]
    nv = len(p["vertices"])
    D = p["dim"]
    ncases = 2 ** nv
    table = dict(
        polytope=name, dim=D, num_cases=ncases,
        vertices=p["vertices"], edges=p["edges"],
        case_to_inoutcut=[None] * ncases,
    )
    full = name in ("SEGMENT", "TRI", "TET")
    if full:
        for k in ("case_to_subcell_to_points", "case_to_subcell_to_inout",
                  "case_to_num_subcells_in", "case_to_num_subcells_out",
                  "case_to_subfacet_to_points", "case_to_subfacet_to_orientation",
                  "case_to_point_to_coordinates"):
            table[k] = [None] * ncases
        table["simplex_facets"] = p["simplex_facets"]
    # CartesianIndices((2,...,2)): first vertex fastest; c==1 -> IN else OUT
    for ci in itertools.product(*([(1, 2)] * nv)):
        ci = ci[::-1]
        vertex_to_value = [IN if c == 1 else OUT for c in ci]
        case = compute_case(vertex_to_value)
        table["case_to_inoutcut"][case - 1] = find_in_out_or_cut(vertex_to_value)
        if not full:
            continue
        coords, values = compute_delaunay_points(vertex_to_value, p["vertices"], p["edges"])
        subcells = delaunay(coords)
        ensure_positive_jacobians(subcells, coords, D)
        inout = compute_subcell_to_inout(subcells, values)
        fpts, fnormals = find_subfacets(coords, subcells, inout, D, p["simplex_facets"])
        orient = setup_subfacet_to_orientation(coords, fpts, fnormals)
        table["case_to_subcell_to_points"][case - 1] = subcells
        table["case_to_subcell_to_inout"][case - 1] = inout
        table["case_to_num_subcells_in"][case - 1] = sum(1 for s in inout if s == IN)
        table["case_to_num_subcells_out"][case - 1] = sum(1 for s in inout if s == OUT)
        table["case_to_subfacet_to_points"][case - 1] = fpts
        table["case_to_subfacet_to_orientation"][case - 1] = orient
        table["case_to_point_to_coordinates"][case - 1] = coords
    return table


def build_simplexify(name):
    """`simplexify(p)` (assumption A2) followed by the reference-coordinate
    positive-Jacobian fix (`CutTriangulations.jl:512-514`)."""
    p = POLYTOPES[name]
    lt = [list(s) for s in p["simplexify"]]
    ensure_positive_jacobians(lt, p["vertices"], p["dim"])
    return dict(raw=p["simplexify"], positive=lt)


def check_invariants(t):
    """Sanity invariants listed in SURVEY.md Appendix B."""
    if "case_to_subcell_to_points" not in t:
        return
    D = t["dim"]
    nv = len(t["vertices"])
    import math
    vol_p = 1.0 / math.factorial(D)
    for c in range(t["num_cases"]):
        coords = np.asarray(t["case_to_point_to_coordinates"][c], float)
        vol = 0.0
        for pts in t["case_to_subcell_to_points"][c]:
            X = coords[[q - 1 for q in pts]]
            J = (X[1:] - X[0]).T if D > 1 else np.array([[X[1][0] - X[0][0]]])
            d = np.linalg.det(J)
            assert d > 0, (t["polytope"], c, pts, d)
            vol += d / math.factorial(D)
        assert abs(vol - vol_p) < 1e-14, (t["polytope"], c, vol)
        for fpts in t["case_to_subfacet_to_points"][c]:
            assert all(q > nv for q in fpts), "subfacet points must be midpoints"
        mirror = t["num_cases"] - 1 - c
        assert t["case_to_subcell_to_points"][c] == t["case_to_subcell_to_points"][mirror]
        assert [-s for s in t["case_to_subcell_to_inout"][c]] == t["case_to_subcell_to_inout"][mirror]
        assert len(t["case_to_subfacet_to_points"][c]) == len(t["case_to_subfacet_to_points"][mirror])
        # orientation * normal(point order) must point from the IN vertices to the OUT vertices
        V = np.asarray(t["vertices"], float)
        outs = [(c >> i) & 1 for i in range(nv)]
        if 0 < sum(outs) < nv:
            d_in_out = V[[i for i in range(nv) if outs[i]]].mean(0) - V[[i for i in range(nv) if not outs[i]]].mean(0)
            for fpts, o in zip(t["case_to_subfacet_to_points"][c], t["case_to_subfacet_to_orientation"][c]):
                n = _unit_normal_from_points([coords[q - 1] for q in fpts])
                assert o * (n @ d_in_out) > 0, (t["polytope"], c, fpts, o)


def main(out_path):
    doc = dict(
        generator="oracle/gen_lookup_tables.py",
        provenance=("restatement of GridapEmbedded LookupTables.jl:46-79 with scipy qhull_r 8.0.2 "
                    "'d Qt Qbb Qc Qz'; connectivity parity UNPINNED (see header)"),
        tables={n: build_lookup_table(n) for n in ("SEGMENT", "TRI", "TET", "QUAD", "HEX")},
        simplexify={n: build_simplexify(n) for n in ("SEGMENT", "TRI", "TET", "QUAD", "HEX")},
    )
    for t in doc["tables"].values():
        check_invariants(t)
    with open(out_path, "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    return doc


if __name__ == "__main__":
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(here, "..", "tests", "golden", "lookup_tables.json")
    doc = main(out)
    for n in ("SEGMENT", "TRI", "TET"):
        t = doc["tables"][n]
        print(n)
        for c in range(t["num_cases"]):
            print(" case", c + 1, t["case_to_subcell_to_points"][c], t["case_to_subcell_to_inout"][c],
                  "facets", t["case_to_subfacet_to_points"][c], t["case_to_subfacet_to_orientation"][c])
    print("simplexify", {k: v["positive"] for k, v in doc["simplexify"].items()})
