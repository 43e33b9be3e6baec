"""Prototype of the closed-form facet numbering of a simplexified Cartesian model (csrc/gecut_facets.cuh, GcSxFacets):
builds the slot table from the simplexify table, decodes every facet id and compares nodes / boundary flags with the
oracle's first-encounter numbering.  Host-only check of the index arithmetic the kernels use."""
import sys, itertools
sys.path.insert(0, 'oracle'); sys.path.insert(0, '.')
import numpy as np

RAW = {2: [[0, 1, 2], [1, 2, 3]], 3: [[0, 1, 2, 6], [0, 1, 4, 6], [1, 2, 3, 6], [1, 3, 6, 7], [1, 4, 5, 6], [1, 5, 6, 7]]}
LF = {2: [[0, 1], [0, 2], [1, 2]], 3: [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]]}


def build(D):
    nt, nlf = len(RAW[D]), D + 1
    ns = nt * nlf
    vtx = [[RAW[D][s // nlf][m] for m in LF[D][s % nlf]] for s in range(ns)]
    face, partner = [-1] * ns, [255] * ns
    for s in range(ns):
        for d in range(D):
            bits = [(v >> d) & 1 for v in vtx[s]]
            if all(b == 0 for b in bits): face[s] = (d << 1)
            if all(b == 1 for b in bits): face[s] = (d << 1) | 1
    for s in range(ns):
        if face[s] < 0:
            key = sorted(vtx[s])
        else:
            d = face[s] >> 1
            key = sorted(v ^ (1 << d) for v in vtx[s])
        for s2 in range(ns):
            if s2 != s and sorted(vtx[s2]) == key:
                partner[s] = s2
        assert partner[s] != 255
    order, rank, nown = [], [], []
    for cls in range(1 << D):
        o, r = [], [255] * ns
        for s in range(ns):
            if face[s] < 0: own = s < partner[s]
            elif face[s] & 1: own = True
            else: own = bool((cls >> (face[s] >> 1)) & 1)
            if own:
                r[s] = len(o); o.append(s)
        order.append(o); rank.append(r); nown.append(len(o))
    C0 = nown[0]
    cd = nown[1] - C0
    for cls in range(1 << D):
        assert nown[cls] == C0 + cd * bin(cls).count("1")
    return dict(D=D, ns=ns, vtx=vtx, face=face, partner=partner, order=order, rank=rank, C0=C0, cd=cd)


def per_row(t, n, j, k):
    D = t["D"]
    return n[0] * (t["C0"] + (t["cd"] if j == 0 else 0) + (t["cd"] if (D == 3 and k == 0) else 0)) + t["cd"]


def per_layer(t, n, k):
    return per_row(t, n, 0, k) + (n[1] - 1) * per_row(t, n, 1, k)


def num_facets(t, n):
    if t["D"] == 2: return per_row(t, n, 0, 0) + (n[1] - 1) * per_row(t, n, 1, 0)
    return per_layer(t, n, 0) + (n[2] - 1) * per_layer(t, n, 1)


def cls_of(D, i, j, k):
    return (1 if i == 0 else 0) | (2 if j == 0 else 0) | (4 if (D == 3 and k == 0) else 0)


def decode(t, n, f):
    D = t["D"]
    rem, k = f, 0
    if D == 3:
        L0 = per_layer(t, n, 0)
        if rem >= L0:
            L1 = per_layer(t, n, 1); rem -= L0; k = 1 + rem // L1; rem %= L1
    j = 0
    R0 = per_row(t, n, 0, k)
    if rem >= R0:
        R1 = per_row(t, n, 1, k); rem -= R0; j = 1 + rem // R1; rem %= R1
    per = t["C0"] + (t["cd"] if j == 0 else 0) + (t["cd"] if (D == 3 and k == 0) else 0)
    i = 0
    if rem >= per + t["cd"]:
        rem -= per + t["cd"]; i = 1 + rem // per; rem %= per
    return (i, j, k), t["order"][cls_of(D, i, j, k)][rem]


def row_base(t, n, j, k):
    D = t["D"]; b = 0
    if D == 3 and k > 0: b += per_layer(t, n, 0) + (k - 1) * per_layer(t, n, 1)
    if j > 0: b += per_row(t, n, 0, k) + (j - 1) * per_row(t, n, 1, k)
    return b


def id_of(t, n, i, j, k, slot):
    D = t["D"]; c = [i, j, k]; cls = cls_of(D, *c)
    if t["rank"][cls][slot] == 255:
        fc = t["face"][slot]
        if fc >= 0: c[fc >> 1] -= 1
        slot = t["partner"][slot]; cls = cls_of(D, *c)
    per = t["C0"] + (t["cd"] if c[1] == 0 else 0) + (t["cd"] if (D == 3 and c[2] == 0) else 0)
    return row_base(t, n, c[1], c[2]) + c[0] * per + (t["cd"] if c[0] > 0 else 0) + t["rank"][cls][slot]


def check(partition):
    import gecut, oracle_binding as ob
    from gecut import CartesianDiscreteModel, simplexify, disk, sphere
    D = len(partition)
    box = (0, 1) * D
    model = simplexify(CartesianDiscreteModel(box, partition))
    oc = ob.oracle_cut_facets(model, disk(0.3, x0=(0.5, 0.5)) if D == 2 else sphere(0.3, x0=(0.5, 0.5, 0.5)))
    t = build(D)
    n = list(partition) + [1] * (3 - D)
    nn = [x + 1 for x in n]
    assert num_facets(t, n) == oc.Nf, (num_facets(t, n), oc.Nf)
    fn = oc.array("facet_nodes").reshape(-1, D)
    isb = oc.array("facet_is_boundary")
    for f in range(oc.Nf):
        (i, j, k), s = decode(t, n, f)
        nodes = []
        for lv in t["vtx"][s]:
            a, b, c = i + (lv & 1), j + ((lv >> 1) & 1), (k + ((lv >> 2) & 1)) if D == 3 else 0
            nodes.append(a + nn[0] * (b + nn[1] * c))
        assert list(fn[f]) == nodes, (f, fn[f], nodes)
        fc = t["face"][s]
        b = False
        if fc >= 0:
            d = fc >> 1
            b = ((i, j, k)[d] == 0) if (fc & 1) == 0 else ((i, j, k)[d] == n[d] - 1)
        assert bool(isb[f]) == b
    # every (cube, slot) maps back to the id of its facet
    seen = {}
    for f in range(oc.Nf): seen[tuple(sorted(fn[f]))] = f
    for k in range(n[2]):
        for j in range(n[1]):
            for i in range(n[0]):
                for s in range(t["ns"]):
                    nodes = []
                    for lv in t["vtx"][s]:
                        a, b, c = i + (lv & 1), j + ((lv >> 1) & 1), (k + ((lv >> 2) & 1)) if D == 3 else 0
                        nodes.append(a + nn[0] * (b + nn[1] * c))
                    assert id_of(t, n, i, j, k, s) == seen[tuple(sorted(nodes))]
    print("ok", partition, oc.Nf)


if __name__ == "__main__":
    for p in [(4, 3), (1, 1), (1, 5), (7, 1), (3, 4, 2), (1, 1, 1), (1, 3, 1), (2, 1, 4), (5, 3, 3)]:
        check(p)
