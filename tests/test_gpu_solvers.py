"""GPU parity: CG, tqli/eigen-estimate, Chebyshev, transfer, CSR, V-cycle.

Tolerances (north_star): Chebyshev/CG residual histories relative 1e-10, identical iteration
counts to convergence.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import mesh as om, operator as oo, solvers as osol
from helpers import OracleLevel, GpuLevel, rel

pytestmark = pytest.mark.gpu
HTOL = 1e-10


@pytest.mark.parametrize("P,n,perturb", [(3, (5, 5, 5), 0.0), (3, (4, 5, 3), 0.2), (6, (2, 3, 2), 0.0), (1, (8, 8, 8), 0.0)])
def test_cg_history_and_iteration_count(ctx, P, n, perturb):
    """examples/cg/main.cpp:238-249 settings: b = 1, x0 = 0, 20 its, rtol 1e-6, coefficients stored."""
    from pmg_dolfinx_b200 import api
    ol = OracleLevel(om.create_box(*n, perturb=perturb), P)
    gl = GpuLevel(ctx, ol)
    dinv = 1.0 / ol.diag()
    xo, ko, al, be, hist, r0 = osol.cg(ol.A, dinv, np.zeros(ol.nd), np.ones(ol.nd), 20, 1e-6)
    cg = api.CGSolver(ctx, ol.nd, 0)
    cg.set_max_iterations(20)
    cg.set_tolerance(1e-6)
    cg.store_coefficients(True)
    x, b = gl.vec(), gl.vec(np.ones(ol.nd))
    k = cg.solve(gl.op, x, b)
    assert k == ko
    g0, gh = cg.history()
    assert abs(g0 - r0) <= HTOL * r0
    assert len(gh) == len(hist) and np.all(np.abs(gh - hist) <= HTOL * np.abs(hist))
    assert len(cg.alphas()) == len(al)
    assert np.all(np.abs(cg.alphas() - al) <= HTOL * np.abs(al))
    assert np.all(np.abs(cg.betas() - be) <= HTOL * np.abs(be))
    e2, _ = rel(x.data_copy(), xo)
    assert e2 < 1e-9
    eig, eo = cg.compute_eigenvalues(), osol.lanczos_eigenvalues(al, be)
    assert np.allclose(eig, eo, rtol=1e-9)


def test_cg_converges_early_with_identical_count(ctx):
    """Break-before-store semantics (src/cg.hpp:206-218, quirk Q5)."""
    from pmg_dolfinx_b200 import api
    ol = OracleLevel(om.create_box(3, 3, 3), 2)
    gl = GpuLevel(ctx, ol)
    dinv = 1.0 / ol.diag()
    b = np.random.default_rng(1).uniform(0, 1, ol.nd)
    xo, ko, al, be, hist, r0 = osol.cg(ol.A, dinv, np.zeros(ol.nd), b, 200, 1e-8)
    assert ko < 200
    cg = api.CGSolver(ctx, ol.nd, 0)
    cg.set_max_iterations(200)
    cg.set_tolerance(1e-8)
    cg.store_coefficients(True)
    x = gl.vec()
    k = cg.solve(gl.op, x, gl.vec(b))
    assert k == ko and len(cg.alphas()) == ko - 1 == len(al)
    assert rel(x.data_copy(), xo)[0] < 1e-9


def test_tqli_golden(ctx):
    import json, os
    from pmg_dolfinx_b200 import api
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tqli.json")))
    d = api.tqli(g["d"], g["e"])
    assert np.allclose(np.sort(d), g["eigs"], rtol=1e-12)


@pytest.mark.parametrize("P,n,its", [(3, (4, 4, 4), 30), (4, (3, 3, 3), 2), (2, (5, 4, 3), 5)])
def test_chebyshev_history(ctx, P, n, its):
    """examples/cg/main.cpp:268-284: non-zero initial guess with BC values set, verbose norms."""
    from pmg_dolfinx_b200 import api
    mesh = om.create_box(*n, perturb=0.1)
    ol = OracleLevel(mesh, P)
    gl = GpuLevel(ctx, ol)
    dinv = 1.0 / ol.diag()
    X = om.dof_coords(mesh, P)
    f = lambda X: 1000.0 * np.exp(-((X[:, 0] - 0.5) ** 2 + (X[:, 1] - 0.5) ** 2) / 0.02)
    b = oo.rhs_collocated(mesh, P, f, ol.bc, g=1.3)
    x0 = np.ones(ol.nd)
    x0[ol.bc != 0] = 1.3
    _, _, al, be, _, _ = osol.cg(ol.A, dinv, np.zeros(ol.nd), np.ones(ol.nd), 20, 1e-6)
    lmax = 1.1 * osol.lanczos_eigenvalues(al, be)[-1]
    ho = []
    xo = osol.chebyshev(ol.A, dinv, x0, b, its, lmax, history=ho)
    ch = api.Chebyshev(ctx, ol.nd, 0, (0.1 * lmax / 1.1, lmax))
    ch.set_max_iterations(its)
    x = gl.vec(x0)
    h = ch.solve(gl.op, x, gl.vec(b), verbose=True)
    ho = np.array(ho)
    assert len(h) == its + 1 and np.all(np.abs(h - ho) <= HTOL * ho)
    assert rel(x.data_copy(), xo)[0] < 1e-11
    # silent path gives the same iterate
    x2 = gl.vec(x0)
    ch.solve(gl.op, x2, gl.vec(b))
    assert rel(x2.data_copy(), xo)[0] < 1e-11
    assert abs(ch.residual(gl.op, x2, gl.vec(b)) - np.linalg.norm(b - ol.A(xo))) <= 1e-9 * np.linalg.norm(b)


def test_rhs_assembly(ctx):
    mesh = om.create_box(3, 4, 3, perturb=0.2)
    P = 3
    ol = OracleLevel(mesh, P)
    gl = GpuLevel(ctx, ol)
    X = om.dof_coords(mesh, P)
    f = oo.f_sines(2, 3, 4, 2.0)
    bo = oo.rhs_collocated(mesh, P, f, ol.bc, g=0.7)
    b = gl.vec()
    gl.op.assemble_rhs(ctx.to_device(f(X)), 0.7, b)
    assert rel(b.data_copy(), bo)[1] < 1e-13


@pytest.mark.parametrize("Pc,Pf", [(1, 3), (1, 2), (2, 4), (3, 6), (1, 8), (2, 2)])
def test_prolong_restrict(ctx, Pc, Pf):
    from pmg_dolfinx_b200 import api
    mesh = om.create_box(3, 2, 4)
    oc, of = OracleLevel(mesh, Pc), OracleLevel(mesh, Pf)
    dmc, dmf = ctx.to_device(oc.dm), ctx.to_device(of.dm)
    rng = np.random.default_rng(Pc * 10 + Pf)
    perm = rng.permutation(mesh.ncells).astype(np.int32)
    it = api.Interpolator(ctx, Pc, Pf, dmc, dmf, oc.nd, of.nd, perm[:7], perm[7:])
    xc, xf = rng.uniform(-1, 1, oc.nd), rng.uniform(-1, 1, of.nd)
    vc, vf = api.Vector(ctx, oc.nd), api.Vector(ctx, of.nd)
    vc.copy_from_host(xc)
    it.interpolate(vc, vf)
    po = oo.prolong(Pc, Pf, oc.dm, of.dm, xc, of.nd)
    assert rel(vf.data_copy(), po)[1] < 1e-13
    vf.copy_from_host(xf)
    vc.set(123.0)  # output is zeroed first (interpolate.hpp:270)
    it.reverse_interpolate(vf, vc)
    ro = oo.restrict(Pc, Pf, oc.dm, of.dm, xf, oc.nd)
    assert rel(vc.data_copy(), ro)[1] < 1e-13
    # R = P^T
    assert abs(np.dot(po, xf) - np.dot(xc, ro)) < 1e-11 * np.abs(po).sum()
    # prolongating a linear field is exact at the fine nodes (test/test_csr.cpp:110-117 restated)
    Xc, Xf = om.dof_coords(mesh, Pc), om.dof_coords(mesh, Pf)
    lin = lambda X: 1.0 + X[:, 0] - 2 * X[:, 1] + 0.5 * X[:, 2]
    vc.copy_from_host(lin(Xc))
    it.interpolate(vc, vf)
    assert np.abs(vf.data_copy() - lin(Xf)).max() < 1e-13


@pytest.mark.parametrize("P,perturb", [(1, 0.0), (1, 0.2), (2, 0.2)])
def test_csr_assembly_and_spmv(ctx, P, perturb):
    """mat-free vs assembled CSR on the same vector (examples/mat_free/main.cpp:270-289 as an assert)."""
    from pmg_dolfinx_b200 import api
    ol = OracleLevel(om.create_box(4, 3, 5, perturb=perturb), P)
    gl = GpuLevel(ctx, ol)
    Ao = oo.assemble_csr(P, ol.dm, ol.G, ol.kappa, ol.bc, ol.nd)
    A = gl.op.to_csr()
    rp, co, va = A.to_host()
    Ag = sp.csr_matrix((va, co, rp), shape=(ol.nd, ol.nd))
    D = (Ag - Ao)
    assert abs(D).max() < 1e-13 * abs(Ao).max()
    assert A.nnz() == Ao.nnz or A.nnz() >= Ao.nnz
    x = np.random.default_rng(2).uniform(-1, 1, ol.nd)
    xv, yv, zv = gl.vec(x), gl.vec(), gl.vec()
    A(xv, yv)
    gl.op(xv, zv)
    assert rel(yv.data_copy(), Ao @ x)[0] < 1e-13
    assert rel(yv.data_copy(), zv.data_copy())[0] < 1e-12
    dv = gl.vec()
    A.get_diag_inverse(dv)
    assert np.allclose(dv.data_copy(), 1.0 / Ao.diagonal(), rtol=1e-13)
    # host-array constructor (MatrixOperator from CSR arrays)
    off = Ao.indptr[1:].astype(np.int32)
    B = api.MatrixOperator(ctx, Ao.indptr, off, Ao.indices, Ao.data)
    B(xv, zv)
    assert rel(zv.data_copy(), Ao @ x)[0] < 1e-14


def _build_hierarchy(ctx, mesh, degrees, nsmooth=2, kappa=2.0):
    from pmg_dolfinx_b200 import api
    ols = [OracleLevel(mesh, P, kappa=kappa) for P in degrees]
    gls = [GpuLevel(ctx, ol) for ol in ols]
    olev, smoothers = [], []
    for ol, gl in zip(ols, gls):
        dinv = 1.0 / ol.diag()
        _, _, al, be, _, _ = osol.cg(ol.A, dinv, np.zeros(ol.nd), np.ones(ol.nd), 20, 1e-6)
        lmax_o = 1.1 * osol.lanczos_eigenvalues(al, be)[-1]
        olev.append(osol.Level(ol.A, dinv, ol.bc.astype(float), lmax_o, nsmooth))
        cg = api.CGSolver(ctx, ol.nd, 0)
        cg.set_max_iterations(20)
        cg.set_tolerance(1e-6)
        cg.store_coefficients(True)
        x, b = gl.vec(), gl.vec(np.ones(ol.nd))
        cg.solve(gl.op, x, b)
        eig = cg.compute_eigenvalues()
        assert abs(1.1 * eig[-1] - lmax_o) < 1e-9 * lmax_o
        s = api.Chebyshev(ctx, ol.nd, 0, (0.1 * eig[-1], 1.1 * eig[-1]))
        s.set_max_iterations(nsmooth)
        smoothers.append(s)
    interps, pro, res = [], [], []
    allc = np.arange(mesh.ncells, dtype=np.int32)
    for i in range(len(degrees) - 1):
        a, b = ols[i], ols[i + 1]
        interps.append(api.Interpolator(ctx, a.P, b.P, gls[i].dofmap, gls[i + 1].dofmap, a.nd, b.nd, allc, allc[:0]))
        pro.append((lambda a, b: lambda xc: oo.prolong(a.P, b.P, a.dm, b.dm, xc, b.nd))(a, b))
        res.append((lambda a, b: lambda xf: oo.restrict(a.P, b.P, a.dm, b.dm, xf, a.nd))(a, b))
    return ols, gls, olev, smoothers, interps, pro, res


@pytest.mark.parametrize("degrees,n,coarse", [((1, 3), (10, 10, 10), False), ((1, 3), (12, 12, 12), False),
                                                ((1, 3), (6, 6, 6), True), ((1, 3), (6, 6, 6), "amg"),
                                                ((1, 3), (12, 12, 12), "amg"),
                                                ((1, 2, 4), (4, 4, 4), False), ((1, 2, 4), (4, 4, 4), True)])
def test_vcycle_history(ctx, degrees, n, coarse):
    """Config 1 (python_tests/pmg.py: 10^3 cells, P3->P1) and a 3-level P4->P2->P1 cycle: per-stage
    residual norms of 4 V-cycles against the oracle."""
    from pmg_dolfinx_b200 import api
    mesh = om.create_box(*n)
    ols, gls, olev, smoothers, interps, pro, res = _build_hierarchy(ctx, mesh, degrees)
    top = ols[-1]
    b = oo.rhs_collocated(mesh, top.P, oo.f_sines(1, 1, 1, 2.0), top.bc)
    cs_o, cs_g = None, None
    if coarse:
        A0 = oo.assemble_csr(ols[0].P, ols[0].dm, ols[0].G, ols[0].kappa, ols[0].bc, ols[0].nd)
        d0 = 1.0 / A0.diagonal()
        cs_o = lambda u0, b0: osol.cg(lambda v: A0 @ v, d0, u0, b0, 60, 1e-10)[0]
        # "amg": smoothed-aggregation PCG (min_coarse 100: a real multilevel hierarchy at 12^3); both coarse
        # solvers converge to 1e-10, so the cycle's iterates agree with the oracle's Jacobi-CG ones
        cs_g = api.CoarseSolverType(ctx, gls[0].op.to_csr(), 60, 1e-10, amg=(coarse == "amg"), min_coarse=100)
    pmg = api.MultigridPreconditioner(ctx, [g.bc for g in gls], flags=2)
    pmg.set_solvers(smoothers)
    pmg.set_operators([g.op for g in gls])
    pmg.set_interpolators(interps)
    pmg.set_coarse_solver(cs_g)
    u = gls[-1].vec()
    bv = gls[-1].vec(b)
    uo = np.zeros(top.nd)
    for it in range(4):
        ho = []
        uo = osol.vcycle(olev, pro, res, b, uo, coarse_solve=cs_o, history=ho)
        rn = pmg.apply(bv, u, verbose=True)
        hg = pmg.diagnostics()
        hov = np.array([h[2] for h in ho])
        assert len(hg) == len(hov)
        assert np.all(np.abs(hg - hov) <= 1e-9 * hov[0]), (it, hg, hov)
        assert abs(rn - hov[-1]) <= 1e-9 * hov[0]
        assert rel(u.data_copy(), uo)[0] < 1e-9
    if coarse:
        assert hov[-1] < 1e-2 * np.linalg.norm(b)


@pytest.mark.parametrize("degrees,n,coarse", [((1, 3), (6, 6, 6), True), ((1, 2, 4), (4, 4, 4), False),
                                                ((1, 2, 4), (4, 5, 3), True), ((1, 2, 4), (4, 5, 3), "amg")])
@pytest.mark.parametrize("flags", [0, 4])
def test_vcycle_default_and_literal_sequence_match_oracle(ctx, degrees, n, coarse, flags):
    """The production cycle (flags 0: the smoother's recurrence residual is restricted with the last
    r -= q folded into the gather, u += P u_c is formed in the prolongation's store, the reference's
    dead last smoothing iteration is dropped, the cycle works on the caller's vectors in place) and
    the literal reference sequence (flags 4) give the oracle's iterates."""
    from pmg_dolfinx_b200 import api
    mesh = om.create_box(*n, perturb=0.1)
    ols, gls, olev, smoothers, interps, pro, res = _build_hierarchy(ctx, mesh, degrees)
    top = ols[-1]
    b = oo.rhs_collocated(mesh, top.P, oo.f_sines(1, 2, 1, 2.0), top.bc)
    cs_o, cs_g = None, None
    if coarse:
        A0 = oo.assemble_csr(ols[0].P, ols[0].dm, ols[0].G, ols[0].kappa, ols[0].bc, ols[0].nd)
        d0 = 1.0 / A0.diagonal()
        cs_o = lambda u0, b0: osol.cg(lambda v: A0 @ v, d0, u0, b0, 200, 1e-12)[0]
        cs_g = api.CoarseSolverType(ctx, gls[0].op.to_csr(), 200, 1e-12, amg=(coarse == "amg"))
    pmg = api.MultigridPreconditioner(ctx, [g.bc for g in gls], flags=flags)
    pmg.set_solvers(smoothers)
    pmg.set_operators([g.op for g in gls])
    pmg.set_interpolators(interps)
    pmg.set_coarse_solver(cs_g)
    u = gls[-1].vec()
    bv = gls[-1].vec(b)
    uo = np.zeros(top.nd)
    b_before = bv.data_copy().copy()
    for it in range(3):
        ho = []
        uo = osol.vcycle(olev, pro, res, b, uo, coarse_solve=cs_o, history=ho)
        rn = pmg.apply(bv, u, verbose=True)
        assert abs(rn - ho[-1][2]) <= 1e-9 * ho[0][2]
        assert rel(u.data_copy(), uo)[0] < 1e-9
    assert np.array_equal(bv.data_copy(), b_before)      # the caller's b is never modified


def _variable_kappa(mesh):
    """A non-constant coefficient: one value per cell, a smooth factor-20 variation plus a seeded jitter."""
    c = mesh.verts[mesh.geom_dofmap].mean(axis=1)
    rng = np.random.default_rng(11)
    return (1.0 + 19.0 * c[:, 0] * c[:, 1] + np.sin(7.0 * c[:, 2]) ** 2) * rng.uniform(0.8, 1.25, mesh.ncells)


@pytest.mark.parametrize("degrees,n,perturb", [((1, 3), (10, 10, 10), 0.0), ((1, 2, 4), (4, 5, 3), 0.15)])
def test_pmg_preconditioned_cg_matches_oracle(ctx, degrees, n, perturb):
    """CGSolver with M^-1 = MultigridPreconditioner::apply (src/cg.hpp:147-222 with the
    V-cycle of src/pmg.hpp:56-155 where the reference multiplies by diag^-1), on config 1 (10^3 cells,
    P3->P1) and on a perturbed 3-level P4->P2->P1 hierarchy, with a NON-CONSTANT kappa per cell.
    alpha / beta / residual history to 1e-10, identical iteration count.  The coarse problem is solved
    exactly on both sides (oracle: sparse LU like python_tests/pmg.py:136-141; product: the AMG
    hierarchy's dense coarsest level), so the preconditioner is the same fixed linear operator."""
    from pmg_dolfinx_b200 import api
    mesh = om.create_box(*n, perturb=perturb)
    kap = _variable_kappa(mesh)
    assert kap.max() / kap.min() > 10
    ols, gls, olev, smoothers, interps, pro, res = _build_hierarchy(ctx, mesh, degrees, kappa=kap)
    top = ols[-1]
    b = oo.rhs_collocated(mesh, top.P, oo.f_sines(2, 1, 3, 1.0), top.bc)
    A0 = sp.csc_matrix(oo.assemble_csr(ols[0].P, ols[0].dm, ols[0].G, ols[0].kappa, ols[0].bc, ols[0].nd))
    lu = spla.splu(A0)
    cs_o = lambda u0, b0: lu.solve(b0)
    M_o = lambda r: osol.vcycle(olev, pro, res, r, np.zeros(top.nd), coarse_solve=cs_o)
    dinv = 1.0 / top.diag()
    xo, ko, al, be, hist, r0 = osol.cg(top.A, dinv, np.zeros(top.nd), b, 30, 1e-8, M=M_o)
    _, kj, *_ = osol.cg(top.A, dinv, np.zeros(top.nd), b, 500, 1e-8)
    assert ko < 30 and ko * 3 < kj          # the V-cycle is a much better preconditioner than Jacobi

    cs_g = api.CoarseSolverType(ctx, gls[0].op.to_csr(), 10, 1e-13, amg=True, min_coarse=5000)
    assert len(cs_g.levels()) == 1          # dense inverse of the whole P1 level: an exact coarse solve
    pmg = api.MultigridPreconditioner(ctx, [g.bc for g in gls])
    pmg.set_solvers(smoothers)
    pmg.set_operators([g.op for g in gls])
    pmg.set_interpolators(interps)
    pmg.set_coarse_solver(cs_g)
    cg = api.CGSolver(ctx, top.nd, 0)
    cg.set_max_iterations(30)
    cg.set_tolerance(1e-8)
    cg.store_coefficients(True)
    cg.set_preconditioner(pmg)
    x, bv = gls[-1].vec(), gls[-1].vec(b)
    k = cg.solve(gls[-1].op, x, bv)
    assert k == ko
    g0, gh = cg.history()
    assert abs(g0 - r0) <= HTOL * r0
    assert len(gh) == len(hist) and np.all(np.abs(gh - hist) <= 1e-9 * np.abs(hist[0]) + HTOL * np.abs(hist))
    assert np.all(np.abs(cg.alphas() - al) <= HTOL * np.abs(al))
    assert np.all(np.abs(cg.betas() - be) <= HTOL * np.abs(be))
    assert rel(x.data_copy(), xo)[0] < 1e-9
    # and it solves the problem: residual of the assembled operator
    Atop = oo.assemble_csr(top.P, top.dm, top.G, top.kappa, top.bc, top.nd)
    assert np.linalg.norm(Atop @ x.data_copy() - b) <= 1e-6 * np.linalg.norm(b)
    # back to Jacobi: the default path is untouched
    cg.set_preconditioner(None)
    x2 = gls[-1].vec()
    cg.set_max_iterations(5)
    assert cg.solve(gls[-1].op, x2, bv) == 5
