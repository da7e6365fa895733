"""Pinning against outputs of the reference ITSELF: tests/golden/solvers_ref.npz holds what the
reference's own Python prototypes (python_tests/chebyshev.py: Chebyshev.cheb4; python_tests/cg.py:
CGSolver.solve / compute_eigs with python_tests/tqli.py) produce on a small SPD matrix; the fixture
and its generator (scripts/make_golden_solvers.py, run in the build container where /root/reference
exists) are committed.  CPU: the oracle restatement must reproduce them; GPU: so must the CUDA path
through the C ABI (Chebyshev / CG on a MatrixOperator built from the same CSR arrays)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import solvers as osol

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "solvers_ref.npz"))


def _matrix():
    n = len(G["indptr"]) - 1
    return sp.csr_matrix((G["data"], G["indices"], G["indptr"]), shape=(n, n))


def _close(a, b, tol):
    return np.abs(np.asarray(a) - np.asarray(b)).max() <= tol * np.abs(np.asarray(b)).max()


def test_oracle_cg_reproduces_reference_prototype():
    A = _matrix()
    dinv = 1.0 / A.diagonal()
    n = A.shape[0]
    x, k, al, be, _, _ = osol.cg(lambda v: A @ v, dinv, np.zeros(n), np.ones(n), 20, 0.0)
    assert k == 20 and _close(al, G["cg_ones_alphas"], 1e-12) and _close(be, G["cg_ones_betas"], 1e-12)
    assert _close(x, G["cg_ones_x"], 1e-12)
    assert _close(osol.lanczos_eigenvalues(al, be), G["cg_ones_eigs"], 1e-12)
    x, k, al, be, _, _ = osol.cg(lambda v: A @ v, dinv, G["x0"], G["b"], 12, 0.0)
    assert k == 12 and _close(al, G["cg_b_alphas"], 1e-12) and _close(be, G["cg_b_betas"], 1e-12)
    assert _close(x, G["cg_b_x"], 1e-12)


@pytest.mark.parametrize("its", [1, 2, 5])
def test_oracle_chebyshev_reproduces_reference_prototype(its):
    A = _matrix()
    dinv = 1.0 / A.diagonal()
    lmax = float(G["eig_range"][1])
    for tag, start in (("zero", np.zeros(A.shape[0])), ("x0", G["x0"])):
        x = osol.chebyshev(lambda v: A @ v, dinv, start, G["b"], its, lmax)
        assert _close(x, G[f"cheb{its}_{tag}_x"], 1e-13)


def _gpu_operator(ctx):
    from pmg_dolfinx_b200 import api
    A = _matrix()
    off = A.indptr[1:].astype(np.int32)          # no ghost columns on one rank
    return api, A, api.MatrixOperator(ctx, A.indptr, off, A.indices, A.data)


@pytest.mark.gpu
def test_cuda_cg_reproduces_reference_prototype(ctx):
    api, A, op = _gpu_operator(ctx)
    n = A.shape[0]
    for b, x0, its, key in ((np.ones(n), np.zeros(n), 20, "cg_ones"), (G["b"], G["x0"], 12, "cg_b")):
        cg = api.CGSolver(ctx, n, 0)
        cg.set_max_iterations(its)
        cg.set_tolerance(0.0)
        cg.store_coefficients(True)
        xv, bv = api.Vector(ctx, n), api.Vector(ctx, n)
        xv.copy_from_host(x0)
        bv.copy_from_host(b)
        assert cg.solve(op, xv, bv) == its
        assert _close(cg.alphas(), G[key + "_alphas"], 1e-10) and _close(cg.betas(), G[key + "_betas"], 1e-10)
        assert _close(xv.data_copy(), G[key + "_x"], 1e-10)
        if key == "cg_ones":
            assert _close(cg.compute_eigenvalues(), G["cg_ones_eigs"], 1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("its", [1, 2, 5])
def test_cuda_chebyshev_reproduces_reference_prototype(ctx, its):
    api, A, op = _gpu_operator(ctx)
    n = A.shape[0]
    er = G["eig_range"]
    for tag, start in (("zero", np.zeros(n)), ("x0", G["x0"])):
        for verbose in (False, True):     # False: the dead last iteration is dropped; True: literal sequence
            ch = api.Chebyshev(ctx, n, 0, (float(er[0]), float(er[1])))
            ch.set_max_iterations(its)
            xv, bv = api.Vector(ctx, n), api.Vector(ctx, n)
            xv.copy_from_host(start)
            bv.copy_from_host(G["b"])
            ch.solve(op, xv, bv, verbose=verbose)
            assert _close(xv.data_copy(), G[f"cheb{its}_{tag}_x"], 1e-10)
