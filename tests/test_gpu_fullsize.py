"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot
reach 10^8 dofs in seconds): symmetry and linearity of the operator at config 3 (P3, 99.9 M dofs),
the known answer A.1 = 0 away from the boundary, D = diag(A) probed through A e, transfer
adjointness (R w, u_c) = (w, P u_c) at config 5 (P2 <-> P4 on 115x116x116 cells), and a V-cycle that
contracts the residual by the same factor as the small-mesh oracle runs.  One GPU, ~1 minute."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _level(ctx, api, mesh, P, want_coords=False):
    sp = mesh.space(P, want_coords=want_coords)
    d = dict(sp=sp, dm=ctx.to_device(sp.dofmap), bc=ctx.to_device(sp.bc))
    return d


@pytest.fixture(scope="module")
def big(ctx):
    from pmg_dolfinx_b200 import api
    n = api.boxmesh_fit(100_000_000, 3)
    assert n == (154, 154, 155)                      # examples/mat_free fit routine, SURVEY 8 header
    mesh = api.BoxMesh(n)
    lv = _level(ctx, api, mesh, 3, want_coords=True)
    xg, gd = ctx.to_device(mesh.xgeom), ctx.to_device(mesh.geom_dofmap)
    kap = torch.full((mesh.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    op = api.MatFreeLaplacian(ctx, 3, kap, lv["dm"], xg, gd, mesh.lcells, mesh.bcells, lv["bc"], lv["sp"].n_owned)
    yield dict(api=api, mesh=mesh, lv=lv, op=op, keep=(xg, gd, kap))
    op.destroy()
    torch.cuda.empty_cache()


def test_config3_symmetry_linearity_known_answer(ctx, big):
    api, op, sp = big["api"], big["op"], big["lv"]["sp"]
    n = sp.n_owned
    assert n == 99_895_954
    g = torch.Generator(device=ctx.device).manual_seed(42)
    x, y, ax, ay, t = (api.Vector(ctx, n) for _ in range(5))
    x.data.copy_(torch.rand(n, dtype=torch.float64, device=ctx.device, generator=g) * 2 - 1)
    y.data.copy_(torch.rand(n, dtype=torch.float64, device=ctx.device, generator=g) * 2 - 1)
    op(x, ax)
    op(y, ay)
    # symmetry (Dirichlet rows are identity rows, Dirichlet columns are zero: still symmetric)
    s1, s2 = api.inner_product(ax, y), api.inner_product(ay, x)
    assert abs(s1 - s2) <= 1e-12 * max(abs(s1), abs(s2))
    # positive definiteness on this sample and linearity A(2x - 3y) = 2Ax - 3Ay
    assert api.inner_product(ax, x) > 0
    t.data.copy_(2.0 * x.data - 3.0 * y.data)
    at = api.Vector(ctx, n)
    op(t, at)
    ref = 2.0 * ax.data - 3.0 * ay.data
    assert float((at.data - ref).norm() / ref.norm()) <= 1e-12
    # known answer: A.1 vanishes on every row none of whose cells touches the boundary (GLL rule is
    # exact for the constant), Dirichlet rows return x (src/laplacian.hpp:273-274)
    x.set(1.0)
    op(x, ax)
    X = torch.from_numpy(sp.coords).to(ctx.device)
    h = 1.0 / 154
    inner = ((X > 1.01 * h) & (X < 1 - 1.01 * h)).all(dim=1)
    bc = torch.from_numpy(sp.bc.astype(np.bool_)).to(ctx.device)
    assert float(ax.data[:n][inner].abs().max()) <= 1e-12
    assert bool((ax.data[:n][bc] == 1.0).all())
    # the value examples/mat_free prints for u = 1 (norm of the result) is mesh-determined
    assert abs(api.norm(ax) - 1134.1484801530153) <= 1e-9 * 1134.1484801530153


def test_config3_diagonal_is_the_operator_diagonal(ctx, big):
    """D^-1 from the matrix-free diagonal kernel vs (A e_i)_i for a few unit vectors."""
    api, op, sp = big["api"], big["op"], big["lv"]["sp"]
    n = sp.n_owned
    dinv = api.Vector(ctx, n)
    op.get_diag_inverse(dinv)
    e, ae = api.Vector(ctx, n), api.Vector(ctx, n)
    for i in (0, 12345, n // 2 + 7, n - 1, 31_415_926):
        e.set(0.0)
        e.data[i] = 1.0
        op(e, ae)
        assert abs(float(ae.data[i]) * float(dinv.data[i]) - 1.0) <= 1e-12


def test_config5_transfers_are_adjoint_and_vcycle_contracts(ctx):
    from pmg_dolfinx_b200 import api
    import bench
    n = api.boxmesh_fit(100_000_000, 4)
    assert n == (115, 116, 116)
    mesh = api.BoxMesh(n)
    pmg, ops, b, eigs, keep = bench.build_problem(ctx, api, torch, mesh, False)
    lv, interps = keep[0], keep[6]
    g = torch.Generator(device=ctx.device).manual_seed(7)
    for it, lc, lf in ((interps[1], lv[1], lv[2]), (interps[0], lv[0], lv[1])):
        nc, nf = lc["sp"].n_owned, lf["sp"].n_owned
        uc, w, pu, rw = api.Vector(ctx, nc), api.Vector(ctx, nf), api.Vector(ctx, nf), api.Vector(ctx, nc)
        uc.data.copy_(torch.rand(nc, dtype=torch.float64, device=ctx.device, generator=g) - 0.5)
        w.data.copy_(torch.rand(nf, dtype=torch.float64, device=ctx.device, generator=g) - 0.5)
        it.interpolate(uc, pu)
        it.reverse_interpolate(w, rw)
        a, c = api.inner_product(rw, uc), api.inner_product(w, pu)
        assert abs(a - c) <= 1e-12 * max(abs(a), abs(c))          # R = P^T (multiplicity-scaled gather)
        # prolongation reproduces the coarse space: a constant stays the constant
        uc.set(3.25)
        it.interpolate(uc, pu)
        assert float((pu.data[:nf] - 3.25).abs().max()) <= 1e-13
    # lambda_max of D^-1 A is mesh-size independent for these elements (oracle, small meshes: same values)
    assert abs(eigs[2] - 2.29) < 0.02 and abs(eigs[1] - 2.14) < 0.02 and abs(eigs[0] - 1.935) < 0.02
    sp = lv[-1]["sp"]
    u = api.Vector(ctx, sp.n_owned)
    r0 = api.norm(b)
    hist = [pmg.apply(b, u, verbose=True) / r0 for _ in range(4)]
    assert all(h1 < 0.5 * h0 for h0, h1 in zip([1.0] + hist[:-1], hist)), hist   # every cycle at least halves it
    assert hist[-1] < 1e-3


def test_config4_cg_p6_200m_dofs(ctx):
    """examples/cg at BASELINE size (P6, 94x98x100 cells, 200 003 785 dofs): 20 Jacobi-CG iterations with
    b = 1 run to the iteration cap like on the small meshes, the Lanczos lambda_max equals the mesh-independent value of the oracle
    runs (2.39) and x is the CG iterate: ||b - A x||_{D^-1} equals the solver's last r.D^-1 r."""
    from pmg_dolfinx_b200 import api
    n = api.boxmesh_fit(200_000_000, 6)
    assert n == (94, 98, 100)
    mesh = api.BoxMesh(n)
    sp = mesh.space(6)
    assert sp.n_owned == 200_003_785
    dm, bc = ctx.to_device(sp.dofmap), ctx.to_device(sp.bc)
    xg, gd = ctx.to_device(mesh.xgeom), ctx.to_device(mesh.geom_dofmap)
    kap = torch.full((mesh.n_cells,), 2.0, dtype=torch.float64, device=ctx.device)
    op = api.MatFreeLaplacian(ctx, 6, kap, dm, xg, gd, mesh.lcells, mesh.bcells, bc, sp.n_owned)
    x, b = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    b.set(1.0)
    cg = api.CGSolver(ctx, sp.n_owned, 0)
    cg.set_max_iterations(20)
    cg.set_tolerance(1e-6)
    cg.store_coefficients(True)
    assert cg.solve(op, x, b) == 20
    rn0, hist = cg.history()
    assert len(hist) == 20 and all(h > 0 for h in hist)   # (r.D^-1 r is not monotone for CG; b = 1 grows first)
    eig = cg.compute_eigenvalues()
    assert abs(eig[-1] - 2.39) < 0.01
    # independent check of the final iterate
    ax, dinv = api.Vector(ctx, sp.n_owned), api.Vector(ctx, sp.n_owned)
    op(x, ax)
    op.get_diag_inverse(dinv)
    r = b.data - ax.data
    rdr = float((r * r * dinv.data).sum())
    assert abs(rdr - hist[-1]) <= 1e-8 * hist[-1]
    op.destroy()
