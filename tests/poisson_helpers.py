"""A downstream consumer of the cut in numpy: the Nitsche + ghost-penalty CutFEM Poisson problem of the reference's
test/GridapEmbeddedTests/PoissonCutFEMTests.jl (manufactured solution u = x + y - z on the sphere R = 0.7, Q1 background
space on the ACTIVE cells), assembled from exactly what this path hands to Gridap:

  * ∫ ∇v·∇u dΩ            -- the element matrices of the path (cut cells: K_cut, uncut IN cells: K_bg),
  * the Nitsche terms on Γ -- from the SubFacetData layout (points, reference points through the glue, normals, bg cell),
  * the ghost penalty      -- on the interior facets between two active cells one of which is cut (GhostSkeleton).

The reference asserts el2/ul2 < 1e-8 and eh1/uh1 < 1e-7 (PoissonCutFEMTests.jl:116-117); u lies in the discrete space, so
the discrete solution reproduces it as long as the layout, the glue, the normals and the matrices are mutually consistent.
The same function serves the CPU oracle (tests without a GPU) and the device results (gpu tests)."""
import numpy as np

import helpers as H

IN, OUT, CUT = -1, 1, 0


def u_exact(x):
    return x[:, 0] + x[:, 1] - x[:, 2]


GRAD_U = np.array([1.0, 1.0, -1.0])


def solve_poisson_cutfem(model, bg_state, cutcells, K_cut, K_bg, sub, fac, gamma_d=10.0, gamma_g=0.1):
    """bg_state: Int8 per bg cell (state of the geometry); cutcells: 1-based ids (rows of K_cut); K_bg: nv x nv matrix of an
    uncut cell; sub = dict(c2p, c2bg, X, R) of the PHYSICAL subcells; fac = dict(c2p, c2bg, X, R, N) of the embedded boundary
    (normals already oriented).  Returns (el2/ul2, eh1/uh1, number of ghost facets)."""
    D = model.D
    assert D == 3 and not model.simplexified
    nv = 1 << D
    cn = model.cell_node_ids() - 1
    xn = model.node_coordinates()
    hvec = np.asarray(model.sizes, float)
    h = hvec[0]
    nn = model.num_nodes
    A = np.zeros((nn, nn))
    rhs = np.zeros(nn)

    def add(cells0, Ke):
        for c, K in zip(cells0, Ke):
            A[np.ix_(cn[c], cn[c])] += K

    # ---- ∫ ∇v·∇u dΩ from the element matrices of the path
    add(cutcells - 1, K_cut)
    in_cells = np.flatnonzero(bg_state == IN)
    add(in_cells, np.broadcast_to(K_bg, (len(in_cells), nv, nv)))
    # ---- Nitsche terms on the embedded boundary (degree-4 rule: v restricted to a plane is cubic)
    xi, w = H.simplex_rule(D - 1, 4)
    Nq = H.p1_shape(xi)  # (nq, D)
    P = fac["X"][fac["c2p"] - 1]   # (nf, D, D)
    Rr = fac["R"][fac["c2p"] - 1]
    meas = H.simplex_meas(P)
    nrm = fac["N"]
    fc = fac["c2bg"] - 1
    for q in range(len(w)):
        eta = np.einsum("i,nid->nd", Nq[q], Rr)
        x = np.einsum("i,nid->nd", Nq[q], P)
        phi, dphi = H.q1_shape_and_grad(eta)
        dn = np.einsum("nad,nd->na", dphi / hvec[None, None, :], nrm)  # n·∇φ_a
        wq = w[q] * meas
        Ke = wq[:, None, None] * ((gamma_d / h) * phi[:, :, None] * phi[:, None, :] - phi[:, :, None] * dn[:, None, :]
                                  - dn[:, :, None] * phi[:, None, :])
        ud = u_exact(x)
        re = wq[:, None] * ((gamma_d / h) * phi * ud[:, None] - dn * ud[:, None])
        for f in range(len(fc)):
            nd = cn[fc[f]]
            A[np.ix_(nd, nd)] += Ke[f]
            rhs[nd] += re[f]
    # ---- ghost penalty on the interior facets between two active cells, at least one of them cut
    n = model.partition
    active = (bg_state != OUT).reshape(n[2], n[1], n[0])
    iscut = (bg_state == CUT).reshape(n[2], n[1], n[0])
    cid = np.arange(model.num_cells).reshape(n[2], n[1], n[0])
    g2 = H.GAUSS2
    nghost = 0
    for d in range(3):
        ax = 2 - d  # array axis of direction d
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax] = slice(0, -1)
        hi[ax] = slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        sel = active[lo] & active[hi] & (iscut[lo] | iscut[hi])
        c1 = cid[lo][sel]
        c2 = cid[hi][sel]
        nghost += len(c1)
        others = [e for e in range(3) if e != d]
        area = hvec[others[0]] * hvec[others[1]]
        G = np.zeros((2 * nv, 2 * nv))
        for qa in range(2):
            for qb in range(2):
                eta1 = np.zeros((1, 3))
                eta2 = np.zeros((1, 3))
                eta1[0, d], eta2[0, d] = 1.0, 0.0
                eta1[0, others[0]] = eta2[0, others[0]] = g2[qa]
                eta1[0, others[1]] = eta2[0, others[1]] = g2[qb]
                _, d1 = H.q1_shape_and_grad(eta1)
                _, d2 = H.q1_shape_and_grad(eta2)
                g = np.concatenate([d1[0, :, d], -d2[0, :, d]]) / hvec[d]  # jump of n·∇φ over the two cells
                G += 0.25 * area * np.outer(g, g)
        G *= gamma_g * h
        for a, b in zip(c1, c2):
            nd = np.concatenate([cn[a], cn[b]])  # the two cells share the nodes of the facet: accumulate duplicates
            np.add.at(A, np.ix_(nd, nd), G)
    # ---- solve on the nodes of the active cells
    act_nodes = np.unique(cn[np.flatnonzero(bg_state != OUT)])
    uh = np.zeros(nn)
    uh[act_nodes] = np.linalg.solve(A[np.ix_(act_nodes, act_nodes)], rhs[act_nodes])
    # ---- errors over Ω (PHYSICAL subcells + IN cells), degree-2 rules as in the reference
    en = u_exact(xn) - uh
    cn1 = model.cell_node_ids()

    def sq(vn, with_grad):
        def f(bgcell, eta, x):
            v, g = H.bg_field_at(model, cn1, bgcell, eta, vn)
            return v * v + ((g * g).sum(1) if with_grad else 0.0)
        return (H.integrate_on_simplices(model, cn1, sub["c2p"], sub["c2bg"], sub["X"], sub["R"], f)
                + H.integrate_on_bgcells(model, cn1, in_cells + 1, f))

    el2, ul2 = np.sqrt(sq(en, False)), np.sqrt(sq(uh, False))
    eh1, uh1 = np.sqrt(sq(en, True)), np.sqrt(sq(uh, True))
    return el2 / ul2, eh1 / uh1, nghost


# --------------------------------------------------------------------------------------------------------------------
def solve_bimaterial_poisson(model, state1, cutcells, K1_cut, K2_cut, K_bg, meas_K1, meas_K2, sub1, sub2, fac,
                             alpha1=4.0, alpha2=3.0, gamma_hat=2.0):
    """test/GridapEmbeddedTests/BimaterialPoissonCutFEMTests.jl: two unfitted P1 fields u1 (disk, ACTIVE geo1 cells) and u2
    (the rest, ACTIVE !geo1 cells, Dirichlet on the box boundary) coupled by the Nitsche interface terms of the CutFEM
    paper, manufactured solution u = x - y/2 in both materials (fluxes jump by (alpha1 - alpha2)∇u).

    state1: Int8 per bg simplex (state of geo1); K1_cut / K2_cut: P1 matrices of the PHYSICAL IN / OUT subcells per cut
    cell (rows = cutcells, 1-based ids); K_bg[type]: matrix of an uncut simplex of that type; meas_K1 / meas_K2:
    get_cell_measure(Ω1 / Ω2, Ω_bg) per bg cell; sub1 / sub2: dict(c2p, c2bg, X, R) of the PHYSICAL subcells of each side;
    fac: dict(c2p, c2bg, X, R, N) of EmbeddedBoundary(cutgeo, geo1, geo2).  Returns (el2/ul2, eh1/uh1)."""
    D = model.D
    assert D == 2 and model.simplexified
    cn1 = model.cell_node_ids()
    cn = cn1 - 1
    xn = model.node_coordinates()
    nn = model.num_nodes
    ntypes = model.simplices_per_cube

    def u_ex(x):
        return x[:, 0] - 0.5 * x[:, 1]

    gradu = np.array([1.0, -0.5])
    A = np.zeros((2 * nn, 2 * nn))
    rhs = np.zeros(2 * nn)
    # ---- bulk terms from the element matrices of the path
    for side, (alpha, Kc, off, st) in enumerate(((alpha1, K1_cut, 0, IN), (alpha2, K2_cut, nn, OUT))):
        for k, c in enumerate(cutcells - 1):
            nd = cn[c] + off
            A[np.ix_(nd, nd)] += alpha * Kc[k]
        for c in np.flatnonzero(state1 == st):
            nd = cn[c] + off
            A[np.ix_(nd, nd)] += alpha * K_bg[c % ntypes]
    # ---- interface terms (degree-2 rule on the segments: products of two linear functions)
    fc = fac["c2bg"] - 1
    meas_KG = np.bincount(fc, weights=0.5 * H.simplex_meas(fac["X"][fac["c2p"] - 1]) * 2.0, minlength=len(state1))
    k1 = (alpha2 * meas_K1) / np.where(state1 == CUT, alpha2 * meas_K1 + alpha1 * meas_K2, 1.0)
    k2 = (alpha1 * meas_K2) / np.where(state1 == CUT, alpha2 * meas_K1 + alpha1 * meas_K2, 1.0)
    beta = (gamma_hat * meas_KG) / np.where(state1 == CUT, meas_K1 / alpha1 + meas_K2 / alpha2, 1.0)
    # constant physical gradients of the P1 shape functions of every bg cell
    Xc = xn[cn]                                   # (nc, 3, 2)
    Jt = Xc[:, 1:, :] - Xc[:, :1, :]              # rows: x_i - x_0
    gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    gphys = np.einsum("nij,aj->nai", np.linalg.inv(Jt), gref)   # (nc, 3, 2): ∇φ_a
    xi, w = H.simplex_rule(1, 2)
    Nq = H.p1_shape(xi)
    P = fac["X"][fac["c2p"] - 1]
    Rr = fac["R"][fac["c2p"] - 1]
    meas = H.simplex_meas(P)
    nrm = fac["N"]
    for f in range(len(fc)):
        c = fc[f]
        dn = gphys[c] @ nrm[f]                    # n·∇φ_a (3,)
        nd1, nd2 = cn[c], cn[c] + nn
        nd = np.concatenate([nd1, nd2])
        for q in range(len(w)):
            eta = Nq[q] @ Rr[f]
            phi = np.array([1.0 - eta[0] - eta[1], eta[0], eta[1]])
            wq = w[q] * meas[f]
            jump = np.concatenate([phi, -phi])                                  # [v] = v1 - v2
            meanq = np.concatenate([k1[c] * alpha1 * dn, k2[c] * alpha2 * dn])  # n·mean_q
            meanu = np.concatenate([k2[c] * phi, k1[c] * phi])                  # mean_u
            A[np.ix_(nd, nd)] += wq * (beta[c] * np.outer(jump, jump) - np.outer(jump, meanq) - np.outer(meanq, jump))
            rhs[nd] += wq * meanu * ((alpha1 - alpha2) * (nrm[f] @ gradu))
    # ---- spaces: u1 on the nodes of the cells active for geo1, u2 on those active for !geo1, Dirichlet on the box boundary
    act1 = np.unique(cn[np.flatnonzero(state1 != OUT)])
    act2 = np.unique(cn[np.flatnonzero(state1 != IN)])
    lo, hi = np.asarray(model.pmin), np.asarray(model.pmax)
    on_bnd = np.any((np.abs(xn - lo) < 1e-12) | (np.abs(xn - hi) < 1e-12), axis=1)
    free2 = act2[~on_bnd[act2]]
    fixed2 = act2[on_bnd[act2]]
    U = np.zeros(2 * nn)
    U[fixed2 + nn] = u_ex(xn[fixed2])
    free = np.concatenate([act1, free2 + nn])
    b = rhs[free] - A[np.ix_(free, fixed2 + nn)] @ U[fixed2 + nn]
    U[free] = np.linalg.solve(A[np.ix_(free, free)], b)
    # ---- errors
    ue = u_ex(xn)

    def sq(vn, sub, st, with_grad):
        def fun(bgcell, eta, x):
            v, g = H.bg_field_at(model, cn1, bgcell, eta, vn)
            return v * v + ((g * g).sum(1) if with_grad else 0.0)
        return (H.integrate_on_simplices(model, cn1, sub["c2p"], sub["c2bg"], sub["X"], sub["R"], fun)
                + H.integrate_on_bgcells(model, cn1, np.flatnonzero(state1 == st) + 1, fun))

    u1, u2 = U[:nn], U[nn:]
    el2 = np.sqrt(sq(ue - u1, sub1, IN, False) + sq(ue - u2, sub2, OUT, False))
    ul2 = np.sqrt(sq(u1, sub1, IN, False) + sq(u2, sub2, OUT, False))
    eh1 = np.sqrt(sq(ue - u1, sub1, IN, True) + sq(ue - u2, sub2, OUT, True))
    uh1 = np.sqrt(sq(u1, sub1, IN, True) + sq(u2, sub2, OUT, True))
    return el2 / ul2, eh1 / uh1


# --------------------------------------------------------------------------------------------------------------------
def solve_poisson_agfem_2d(model, bg_state, cutcells, K_cut, K_bg, cell_to_cellin, sub, fac, gamma_d=10.0):
    """test/GridapEmbeddedTests/PoissonAgFEMTests.jl: Nitsche Poisson WITHOUT ghost penalty on the aggregated Q1 space
    (AgFEMSpace, src/AgFEM/AgFEMSpaces.jl:57-150): a node all of whose active cells are cut (aggregated) cells is
    constrained to the extension of the shape functions of the root cell of its owner (the touching active cell with the
    largest id); u = x - y.  Returns (el2/ul2, eh1/uh1, number of constrained nodes)."""
    D = model.D
    assert D == 2 and not model.simplexified
    nv = 4
    cn1 = model.cell_node_ids()
    cn = cn1 - 1
    xn = model.node_coordinates()
    hvec = np.asarray(model.sizes, float)
    h = hvec[0]
    nn = model.num_nodes

    def u_ex(x):
        return x[:, 0] - x[:, 1]

    A = np.zeros((nn, nn))
    rhs = np.zeros(nn)
    for k, c in enumerate(cutcells - 1):
        A[np.ix_(cn[c], cn[c])] += K_cut[k]
    in_cells = np.flatnonzero(bg_state == IN)
    for c in in_cells:
        A[np.ix_(cn[c], cn[c])] += K_bg
    xi, w = H.simplex_rule(1, 2)
    Nq = H.p1_shape(xi)
    P = fac["X"][fac["c2p"] - 1]
    Rr = fac["R"][fac["c2p"] - 1]
    meas = H.simplex_meas(P)
    nrm = fac["N"]
    fc = fac["c2bg"] - 1
    for q in range(len(w)):
        eta = np.einsum("i,nid->nd", Nq[q], Rr)
        x = np.einsum("i,nid->nd", Nq[q], P)
        phi, dphi = H.q1_shape_and_grad(eta)
        dn = np.einsum("nad,nd->na", dphi / hvec[None, None, :], nrm)
        wq = w[q] * meas
        Ke = wq[:, None, None] * ((gamma_d / h) * phi[:, :, None] * phi[:, None, :] - phi[:, :, None] * dn[:, None, :]
                                  - dn[:, :, None] * phi[:, None, :])
        ud = u_ex(x)
        re = wq[:, None] * ((gamma_d / h) * phi * ud[:, None] - dn * ud[:, None])
        for f in range(len(fc)):
            nd = cn[fc[f]]
            A[np.ix_(nd, nd)] += Ke[f]
            rhs[nd] += re[f]
    # ---- AgFEM constraints
    active = np.flatnonzero(bg_state != OUT)
    node_isagg = np.ones(nn, bool)
    node_owner = np.full(nn, -1)
    touched = np.zeros(nn, bool)
    for c in active:  # ascending ids: the last writer is the touching active cell with the largest id
        aggregated = cell_to_cellin[c] != c + 1
        for nd in cn[c]:
            touched[nd] = True
            node_isagg[nd] &= aggregated
            node_owner[nd] = c
    act_nodes = np.flatnonzero(touched)
    masters = act_nodes[~node_isagg[act_nodes]]
    col = -np.ones(nn, int)
    col[masters] = np.arange(len(masters))
    T = np.zeros((nn, len(masters)))
    T[masters, col[masters]] = 1.0
    slaves = act_nodes[node_isagg[act_nodes]]
    for nd in slaves:
        root = cell_to_cellin[node_owner[nd]] - 1
        rn = cn[root]
        assert np.all(col[rn] >= 0)  # the nodes of a root cell are never constrained
        eta = ((xn[nd] - xn[rn[0]]) / hvec)[None, :]  # reference point of the node in the ROOT cell (outside [0,1]^2)
        phi, _ = H.q1_shape_and_grad(eta)
        T[nd, col[rn]] = phi[0]
    Ar = T.T @ A @ T
    br = T.T @ rhs
    uh = T @ np.linalg.solve(Ar, br)
    # ---- errors over Ω
    en = u_ex(xn) - uh

    def sq(vn, with_grad):
        def fun(bgcell, eta, x):
            v, g = H.bg_field_at(model, cn1, bgcell, eta, vn)
            return v * v + ((g * g).sum(1) if with_grad else 0.0)
        return (H.integrate_on_simplices(model, cn1, sub["c2p"], sub["c2bg"], sub["X"], sub["R"], fun)
                + H.integrate_on_bgcells(model, cn1, in_cells + 1, fun))

    return (np.sqrt(sq(en, False)) / np.sqrt(sq(uh, False)), np.sqrt(sq(en, True)) / np.sqrt(sq(uh, True)), len(slaves))


# --------------------------------------------------------------------------------------------------------------------
def solve_tracefem_projection_2d(model, bg_state, fac, gamma_d=10.0, gamma_g=0.1):
    """test/GridapEmbeddedTests/TraceFEMTests.jl: L2 projection of ud = x - y onto the P1 space of the CUT cells, posed on
    the embedded boundary Γ and stabilised by the gradient jumps over the interior edges between two CUT cells
    (GhostSkeleton(cutgeom, CUT, geom)).  Returns (l2 error on Γ, number of ghost edges)."""
    assert model.D == 2 and model.simplexified
    cn1 = model.cell_node_ids()
    cn = cn1 - 1
    xn = model.node_coordinates()
    nn = model.num_nodes
    h = float(model.sizes[0])

    def ud(x):
        return x[:, 0] - x[:, 1]

    A = np.zeros((nn, nn))
    rhs = np.zeros(nn)
    xi, w = H.simplex_rule(1, 2)
    Nq = H.p1_shape(xi)
    P = fac["X"][fac["c2p"] - 1]
    Rr = fac["R"][fac["c2p"] - 1]
    meas = H.simplex_meas(P)
    fc = fac["c2bg"] - 1
    for f in range(len(fc)):
        nd = cn[fc[f]]
        for q in range(len(w)):
            eta = Nq[q] @ Rr[f]
            x = (Nq[q] @ P[f])[None, :]
            phi = np.array([1.0 - eta[0] - eta[1], eta[0], eta[1]])
            A[np.ix_(nd, nd)] += w[q] * meas[f] * (gamma_d / h) * np.outer(phi, phi)
            rhs[nd] += w[q] * meas[f] * (gamma_d / h) * phi * ud(x)[0]
    # ---- ghost penalty over the interior edges shared by two CUT cells
    cut_cells = np.flatnonzero(bg_state == CUT)
    Xc = xn[cn]
    Jt = Xc[:, 1:, :] - Xc[:, :1, :]
    gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    gphys = np.einsum("nij,aj->nai", np.linalg.inv(Jt), gref)
    edge_to_cells = {}
    for c in cut_cells:
        for a, b in ((0, 1), (0, 2), (1, 2)):
            key = tuple(sorted((cn[c][a], cn[c][b])))
            edge_to_cells.setdefault(key, []).append(c)
    nghost = 0
    for (na, nb), cells in edge_to_cells.items():
        if len(cells) != 2:
            continue
        nghost += 1
        c1, c2 = cells
        t = xn[nb] - xn[na]
        length = np.linalg.norm(t)
        nrm = np.array([t[1], -t[0]]) / length
        g = np.concatenate([gphys[c1] @ nrm, -(gphys[c2] @ nrm)])  # jump of n·∇φ (constant along the edge)
        nd = np.concatenate([cn[c1], cn[c2]])
        np.add.at(A, np.ix_(nd, nd), gamma_g * h * length * np.outer(g, g))
    act = np.unique(cn[cut_cells])
    uh = np.zeros(nn)
    uh[act] = np.linalg.solve(A[np.ix_(act, act)], rhs[act])
    en = uh - ud(xn)
    err2 = 0.0
    for f in range(len(fc)):
        nd = cn[fc[f]]
        for q in range(len(w)):
            eta = Nq[q] @ Rr[f]
            phi = np.array([1.0 - eta[0] - eta[1], eta[0], eta[1]])
            err2 += w[q] * meas[f] * float(phi @ en[nd]) ** 2
    return np.sqrt(err2), nghost
