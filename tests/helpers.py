"""numpy helpers shared by the CPU (oracle) and GPU parity tests: generic integrands evaluated on the
layout arrays (SubCellData / SubFacetData), used for the reference's own known-answer checks
(divergence theorem, areas) -- test/InterfacesTests/EmbeddedDiscretizationsTests.jl:48-107."""
import numpy as np

SQ3 = np.sqrt(3.0)
GAUSS2 = np.array([0.5 - 0.5 / SQ3, 0.5 + 0.5 / SQ3])


def simplex_rule(ds, degree=2):
    if ds == 1:
        if degree <= 3:
            return GAUSS2[:, None], np.array([0.5, 0.5])
        x, w = np.polynomial.legendre.leggauss(4)
        return (0.5 * (x + 1))[:, None], 0.5 * w
    if ds == 2:
        if degree <= 2:
            return np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.full(3, 1 / 6)
        # degree-4 6-point rule (Dunavant)
        a1, b1, w1 = 0.445948490915965, 0.108103018168070, 0.223381589678011
        a2, b2, w2 = 0.091576213509771, 0.816847572980459, 0.109951743655322
        pts = np.array([[a1, a1], [a1, b1], [b1, a1], [a2, a2], [a2, b2], [b2, a2]])
        return pts, 0.5 * np.array([w1, w1, w1, w2, w2, w2])
    a, b = (5 - np.sqrt(5)) / 20, (5 + 3 * np.sqrt(5)) / 20
    return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 1 / 24)


def p1_shape(xi):
    """(nq, ds) -> (nq, ds+1)"""
    return np.concatenate([1 - xi.sum(1, keepdims=True), xi], axis=1)


def q1_shape_and_grad(eta):
    """eta (N,D) reference points in the n-cube -> phi (N,nv), dphi (N,nv,D), vertices x fastest."""
    N, D = eta.shape
    nv = 2 ** D
    phi = np.ones((N, nv))
    dphi = np.ones((N, nv, D))
    for a in range(nv):
        for e in range(D):
            bit = (a >> e) & 1
            f = eta[:, e] if bit else 1 - eta[:, e]
            df = np.full(N, 1.0 if bit else -1.0)
            phi[:, a] *= f
            for d in range(D):
                dphi[:, a, d] *= df if d == e else f
    return phi, dphi


def bg_field_at(model, cell_nodes, bgcell, eta, vnodal):
    """value and physical gradient of the bg FE function (Q1 on n-cubes / P1 on simplices) of cells `bgcell`
    (1-based, (N,)) at reference points eta (N,D)."""
    D = model.D
    nodes = cell_nodes[bgcell - 1] - 1  # (N, nv)
    vals = vnodal[nodes]
    if not model.simplexified:
        phi, dphi = q1_shape_and_grad(eta)
        h = np.asarray(model.sizes)
        v = (phi * vals).sum(1)
        g = (dphi * vals[:, :, None]).sum(1) / h[None, :]
        return v, g
    X = model.node_coordinates()[nodes]  # (N, D+1, D)
    phi = np.concatenate([1 - eta.sum(1, keepdims=True), eta], axis=1)
    v = (phi * vals).sum(1)
    Jt = X[:, 1:, :] - X[:, :1, :]  # (N, D, D): row i = x_{i+1}-x_1
    gref = np.concatenate([-np.ones((1, D)), np.eye(D)], axis=0)  # (D+1, D)
    gv = np.einsum("na,ad->nd", vals, gref)  # d v / d xi
    g = np.linalg.solve(Jt, gv[:, :, None])[:, :, 0]
    return v, g


def simplex_meas(X):
    """X (N, ds+1, D) -> measure of the affine map (|det J| or sqrt(det JtJ))."""
    J = X[:, 1:, :] - X[:, :1, :]
    ds, D = J.shape[1], J.shape[2]
    if ds == D:
        return np.abs(np.linalg.det(J))
    G = np.einsum("nid,njd->nij", J, J)
    return np.sqrt(np.abs(np.linalg.det(G)))


def integrate_on_simplices(model, cell_nodes, c2p, c2bg, X, R, fun, degree=2):
    """sum over the simplices (c2p rows, 1-based ids into X/R) of ∫ fun(bgcell, eta, x) with the reference rule."""
    if len(c2p) == 0:
        return 0.0
    ds = c2p.shape[1] - 1
    xi, w = simplex_rule(ds, degree)
    N = p1_shape(xi)  # (nq, ds+1)
    Xc = X[c2p - 1]
    Rc = R[c2p - 1]
    meas = simplex_meas(Xc)
    total = 0.0
    for q in range(len(w)):
        eta = np.einsum("i,nid->nd", N[q], Rc)
        x = np.einsum("i,nid->nd", N[q], Xc)
        total += (w[q] * meas * fun(c2bg, eta, x)).sum()
    return total


def integrate_on_bgcells(model, cell_nodes, ids, fun):
    """sum over uncut bg cells (1-based ids) of ∫ fun with tensor Gauss (n-cubes) / degree-2 rule (simplices)."""
    if len(ids) == 0:
        return 0.0
    D = model.D
    ids = np.asarray(ids)
    if not model.simplexified:
        vol = np.prod(model.sizes)
        total = 0.0
        for q in range(2 ** D):
            eta = np.tile(np.array([GAUSS2[(q >> d) & 1] for d in range(D)]), (len(ids), 1))
            total += (vol / 2 ** D * fun(ids, eta, None)).sum()
        return total
    xi, w = simplex_rule(D, 2)
    Xn = model.node_coordinates()[cell_nodes[ids - 1] - 1]
    meas = simplex_meas(Xn)
    total = 0.0
    for q in range(len(w)):
        eta = np.tile(xi[q], (len(ids), 1))
        total += (w[q] * meas * fun(ids, eta, None)).sum()
    return total


def monte_carlo_volume(model, inside_fun, n=400000, seed=0):
    rng = np.random.default_rng(seed)
    lo, hi = np.asarray(model.pmin), np.asarray(model.pmax)
    x = lo + rng.random((n, model.D)) * (hi - lo)
    return np.prod(hi - lo) * inside_fun(x).mean()
