"""CPU: the oracle against the reference's golden vectors and analytic known answers
(SURVEY 8c).  tqli is the only item the reference pins numerically (python_tests/tqli.py)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import gll, mesh as om, operator as oo, solvers as osol

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_tqli_golden_vectors():
    g = json.load(open(os.path.join(GOLD, "tqli.json")))
    d, e = np.array(g["d"]), np.array(g["e"])
    assert osol.tqli(d, e) == 0
    assert np.allclose(np.sort(d), g["eigs"])          # python_tests/tqli.py:94-99


@pytest.mark.parametrize("P", range(1, 9))
def test_gll_tables(P):
    x, w, D = gll.tables(P)
    assert abs(w.sum() - 1.0) < 1e-14 and np.all(np.diff(x) > 0) and x[0] == 0.0 and x[-1] == 1.0
    for k in range(0, 2 * P):                          # P+1 GLL points integrate degree 2P-1 exactly
        assert abs(np.dot(w, x ** k) - 1.0 / (k + 1)) < 1e-13
    for k in range(1, P + 1):                          # nodal derivative exact on degree <= P
        assert np.allclose(D @ x ** k, k * x ** (k - 1), atol=1e-11)
    assert np.allclose(D.sum(axis=1), 0, atol=1e-12)


def test_gll_known_values():
    x, w = gll.gll_points_weights(3)
    assert np.allclose(x, [0, 0.5, 1]) and np.allclose(w, [1 / 6, 2 / 3, 1 / 6])
    x, w = gll.gll_points_weights(4)
    a = 0.5 * (1 - 1 / np.sqrt(5))
    assert np.allclose(x, [0, a, 1 - a, 1]) and np.allclose(w, [1 / 12, 5 / 12, 5 / 12, 1 / 12])


@pytest.mark.parametrize("P", [1, 2, 3, 4])
@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_matfree_equals_assembled_and_diag(P, perturb):
    """examples/mat_free/main.cpp:270-289 restated as an assert; D = diag(A_csr) (quirk Q4)."""
    m = om.create_box(3, 2, 3, perturb=perturb)
    dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
    G, detJ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    assert (detJ > 0).all()
    kap = np.full(m.ncells, 2.0)
    x = np.random.default_rng(42).uniform(-1, 1, nd)
    A = oo.assemble_csr(P, dm, G, kap, bc, nd)
    y = oo.apply(P, dm, G, kap, bc, x)
    assert np.linalg.norm(y - A @ x) <= 1e-12 * np.linalg.norm(A @ x)
    assert np.allclose(oo.diagonal(P, dm, G, kap, bc, nd), A.diagonal(), rtol=1e-13)
    assert abs(A - A.T).max() < 1e-13
    nobc = np.zeros_like(bc)
    X = om.dof_coords(m, P)
    for f in (np.ones(nd), X[:, 0] - 2 * X[:, 1] + X[:, 2]):
        assert np.abs(oo.apply(P, dm, G, kap, nobc, f)[bc == 0]).max() < 1e-12


def test_uniform_cube_geometry_known_answer():
    n, P = 5, 3
    m = om.create_box(n, n, n)
    G, detJ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    w = oo.weights_3d(P)
    assert np.allclose(detJ, 1.0 / n ** 3)
    assert np.abs(G[..., [1, 2, 4]]).max() < 1e-15 * np.abs(G).max()
    assert np.allclose(G[..., 0], w[None] / n) and np.allclose(G[..., 3], w[None] / n)
    # literal reference detJ coincides on the axis-aligned cube (quirk Q17)
    G2, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P, literal_detj=True)
    assert np.allclose(G, G2, rtol=1e-13, atol=1e-18)


def test_transfer_properties():
    """python_tests/interpolation_matrix.py:65,78 restated: element-local P with 1/multiplicity ==
    global interpolation, and R = P^T."""
    m = om.create_box(3, 3, 2)
    Pc, Pf = 1, 3
    dc, df = om.dofmap(m, Pc), om.dofmap(m, Pf)
    nc, nf = om.num_dofs(m, Pc), om.num_dofs(m, Pf)
    Xc, Xf = om.dof_coords(m, Pc), om.dof_coords(m, Pf)
    lin = lambda X: 2 + X[:, 0] + 3 * X[:, 1] - X[:, 2]
    assert np.allclose(oo.prolong(Pc, Pf, dc, df, lin(Xc), nf), lin(Xf), atol=1e-13)
    M = oo.local_interp_matrix(Pc, Pf)
    b = np.arange(nc, dtype=float)
    w = np.zeros(nf)
    np.add.at(w, df.reshape(-1), (b[dc] @ M.T).reshape(-1))
    w /= oo.multiplicity(df, nf)
    assert np.allclose(w, oo.prolong(Pc, Pf, dc, df, b, nf))
    rng = np.random.default_rng(0)
    a, f = rng.normal(size=nc), rng.normal(size=nf)
    assert abs(np.dot(oo.prolong(Pc, Pf, dc, df, a, nf), f) - np.dot(a, oo.restrict(Pc, Pf, dc, df, f, nc))) < 1e-11


def test_cg_cpp_semantics_and_eigs():
    """Break before store (src/cg.hpp:206-218); Lanczos estimates bracket the true spectrum."""
    m = om.create_box(4, 4, 4)
    P = 2
    dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
    G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    kap = np.full(m.ncells, 2.0)
    A = oo.assemble_csr(P, dm, G, kap, bc, nd)
    dinv = 1.0 / A.diagonal()
    x, k, al, be, hist, r0 = osol.cg(lambda v: A @ v, dinv, np.zeros(nd), np.ones(nd), 500, 1e-8)
    assert k < 500 and len(al) == k - 1 and len(hist) == k
    assert np.linalg.norm(A @ x - 1.0) < 1e-6 * np.sqrt(nd)
    eig = osol.lanczos_eigenvalues(al, be)
    true = np.sort(np.real(np.linalg.eigvals((dinv[:, None] * A.toarray()))))
    assert eig[-1] <= true[-1] * (1 + 1e-8) and eig[-1] > 0.9 * true[-1]


def test_vcycle_config1_converges():
    """Config 1 (python_tests/pmg.py:60-70: 10^3 cells, P3 -> P1, exact coarse solve)."""
    n, degs = 6, [1, 3]
    m = om.create_box(n, n, n)
    kap = np.full(m.ncells, 2.0)
    lev, dms = [], {}
    for P in degs:
        dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
        dms[P] = (dm, nd)
        G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
        A = (lambda P, dm, G, bc: (lambda x: oo.apply(P, dm, G, kap, bc, x)))(P, dm, G, bc)
        dinv = 1.0 / oo.diagonal(P, dm, G, kap, bc, nd)
        _, _, al, be, _, _ = osol.cg(A, dinv, np.zeros(nd), np.ones(nd), 20, 1e-6)
        lev.append(osol.Level(A, dinv, bc.astype(float), 1.1 * osol.lanczos_eigenvalues(al, be)[-1], 2))
    Pc, Pf = degs
    pro = [lambda xc: oo.prolong(Pc, Pf, dms[Pc][0], dms[Pf][0], xc, dms[Pf][1])]
    res = [lambda xf: oo.restrict(Pc, Pf, dms[Pc][0], dms[Pf][0], xf, dms[Pc][1])]
    A0 = oo.assemble_csr(Pc, dms[Pc][0], oo.geometry_factors(m.verts, m.geom_dofmap, Pc)[0], kap,
                         om.bc_marker(m, Pc), dms[Pc][1])
    lu = spla.splu(A0.tocsc())
    b = oo.rhs_collocated(m, Pf, oo.f_sines(1, 1, 1, 2.0), om.bc_marker(m, Pf))
    u = np.zeros_like(b)
    r0 = np.linalg.norm(b)
    for _ in range(8):
        u = osol.vcycle(lev, pro, res, b, u, coarse_solve=lambda u0, b0: lu.solve(b0))
    assert np.linalg.norm(b - lev[-1].A(u)) < 1e-5 * r0
    X = om.dof_coords(m, Pf)
    ue = np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2])
    assert np.abs(u - ue).max() < 5e-5


def test_partition_emulation_matches_single_rank():
    """R-rank algorithm run sequentially == 1-rank result: ghost cells are
    recomputed, so owned rows are complete after one forward scatter."""
    n, pg, P = (5, 4, 6), (2, 2, 2), 2
    m = om.create_box(*n, perturb=0.1)
    dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
    G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    kap = np.full(m.ncells, 2.0)
    x = np.random.default_rng(9).uniform(-1, 1, nd)
    y = oo.apply(P, dm, G, kap, bc, x)
    parts = om.partition(m, pg, [P])
    xs = []
    for p in parts:
        lv = p.levels[P]
        v = np.zeros(lv.n_owned + lv.n_ghost)
        v[: lv.n_owned] = x[lv.l2g[: lv.n_owned]]
        xs.append(v)
    om.scatter_fwd(parts, P, xs)
    yg = np.full(nd, np.nan)
    for p, xv in zip(parts, xs):
        lv = p.levels[P]
        assert np.array_equal(xv, x[lv.l2g])
        Gl, _ = oo.geometry_factors(p.verts, p.geom_dofmap, P)
        yl = np.zeros_like(xv)
        kl = np.full(len(p.cells), 2.0)
        oo.apply_cells(P, lv.dofmap, Gl, kl, lv.bc, xv, yl, p.lcells)
        oo.apply_cells(P, lv.dofmap, Gl, kl, lv.bc, xv, yl, p.bcells)
        yg[lv.l2g[: lv.n_owned]] = yl[: lv.n_owned]
    assert not np.isnan(yg).any()
    assert np.linalg.norm(yg - y) < 1e-12 * np.linalg.norm(y)


def test_oracle_pmg_preconditioned_cg_solves_variable_kappa_problem():
    """CPU: the oracle twin of the PMG-preconditioned CG with a non-constant kappa: the
    V-cycle as M^-1 converges to the direct solution in far fewer iterations than Jacobi."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    from oracle import solvers as osol
    mesh = om.create_box(4, 4, 4, perturb=0.1)
    c = mesh.verts[mesh.geom_dofmap].mean(axis=1)
    kap = 1.0 + 19.0 * c[:, 0] * c[:, 1]
    degrees, lev, dms = (1, 3), [], {}
    for P in degrees:
        dm, bc, nd = om.dofmap(mesh, P), om.bc_marker(mesh, P), om.num_dofs(mesh, P)
        G, _ = oo.geometry_factors(mesh.verts, mesh.geom_dofmap, P)
        A = (lambda P, dm, G, bc: lambda v: oo.apply(P, dm, G, kap, bc, v))(P, dm, G, bc)
        dinv = 1.0 / oo.diagonal(P, dm, G, kap, bc, nd)
        _, _, al, be, _, _ = osol.cg(A, dinv, np.zeros(nd), np.ones(nd), 20, 1e-6)
        lev.append(osol.Level(A, dinv, bc.astype(float), 1.1 * osol.lanczos_eigenvalues(al, be)[-1], 2))
        dms[P] = (dm, nd, G, bc)
    pro = [lambda xc: oo.prolong(1, 3, dms[1][0], dms[3][0], xc, dms[3][1])]
    res = [lambda xf: oo.restrict(1, 3, dms[1][0], dms[3][0], xf, dms[1][1])]
    A0 = sp.csc_matrix(oo.assemble_csr(1, dms[1][0], dms[1][2], kap, dms[1][3], dms[1][1]))
    lu = spla.splu(A0)
    nd = dms[3][1]
    b = oo.rhs_collocated(mesh, 3, oo.f_sines(1, 2, 1, 1.0), dms[3][3])
    M = lambda r: osol.vcycle(lev, pro, res, r, np.zeros(nd), coarse_solve=lambda u0, b0: lu.solve(b0))
    x, k, al, be, hist, r0 = osol.cg(lev[1].A, lev[1].dinv, np.zeros(nd), b, 40, 1e-10, M=M)
    _, kj, *_ = osol.cg(lev[1].A, lev[1].dinv, np.zeros(nd), b, 1000, 1e-10)
    At = oo.assemble_csr(3, dms[3][0], dms[3][2], kap, dms[3][3], nd)
    xd = spla.spsolve(sp.csc_matrix(At), b)
    assert k < 40 and 3 * k < kj
    assert np.linalg.norm(x - xd) <= 1e-8 * np.linalg.norm(xd)
    assert np.all(np.diff(hist) < 0)            # monotone in the M^-1 norm: M^-1 is SPD


@pytest.mark.parametrize("P", [1, 2, 4])
def test_oracle_lifting_reproduces_the_dirichlet_extension(P):
    """apply_lifting + set_bc (examples/pmg/main.cpp:293-295): with f = 0 and linear Dirichlet data on an
    affine mesh the discrete solution is that linear field (it is in the kernel of the unconstrained
    operator and in the discrete space), so A_bc u = b must hold for u = g."""
    import scipy.sparse as sp
    mesh = om.create_box(3, 4, 2)
    dm, bc, nd = om.dofmap(mesh, P), om.bc_marker(mesh, P), om.num_dofs(mesh, P)
    X = om.dof_coords(mesh, P)
    gfield = 0.7 + 2.0 * X[:, 0] - 1.5 * X[:, 1] + 0.25 * X[:, 2]
    kap = np.random.default_rng(0).uniform(1.0, 3.0, mesh.ncells) * 0 + 2.5   # constant: linear fields stay harmonic
    b = oo.rhs_collocated(mesh, P, lambda Y: np.zeros(len(Y)), bc, g=gfield, kappa=kap)
    G, _ = oo.geometry_factors(mesh.verts, mesh.geom_dofmap, P)
    r = oo.apply(P, dm, G, np.full(mesh.ncells, 2.5), bc, gfield) - b
    assert np.abs(r).max() <= 1e-12 * max(np.abs(b).max(), 1.0)
    assert np.array_equal(b[bc != 0], gfield[bc != 0])
