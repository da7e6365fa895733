"""Golden vectors for the Chebyshev-4 smoother and the Jacobi-CG / Lanczos estimate, produced by the
REFERENCE'S OWN Python prototypes (python_tests/chebyshev.py: Chebyshev.cheb4, python_tests/cg.py:
CGSolver.solve / compute_eigs, python_tests/tqli.py) run here, in this container, on a small SPD matrix.

The prototypes are written against PETSc Mat/Vec objects and import dolfinx / mpi4py / petsc4py / ufl /
basix at module level (none installed).  Only those IMPORTS are stubbed; the algorithm code that runs is
the reference's, unmodified, on numpy-backed duck types that provide the handful of PETSc methods it
calls (Mat: @, getDiagonal, createVecRight; Vec: dot, norm, reciprocal, set, copy, size, arithmetic).

The matrix is the oracle's assembled P2 Laplacian on a perturbed 3x3x3 box (Dirichlet rows = identity),
stored in the fixture so the tests do not depend on the oracle's assembly.

    python scripts/make_golden_solvers.py        ->  tests/golden/solvers_ref.npz   (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/python_tests"
sys.path.insert(0, ROOT)


class Vec(np.ndarray):
    """numpy array with the PETSc.Vec methods the prototypes use."""

    def norm(self):
        return float(np.linalg.norm(np.asarray(self)))

    def reciprocal(self):
        np.divide(1.0, np.asarray(self), out=np.asarray(self))

    def set(self, v):
        np.asarray(self)[:] = v


def vec(a):
    return np.array(a, dtype=np.float64).view(Vec)


class Mat:
    """scipy CSR with the PETSc.Mat methods the prototypes use."""

    def __init__(self, A):
        self.A = sp.csr_matrix(A)

    def __matmul__(self, x):
        return vec(self.A @ np.asarray(x))

    def getDiagonal(self):
        return vec(self.A.diagonal())

    def createVecRight(self):
        return vec(np.zeros(self.A.shape[1]))


def stub_imports():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m
    dummy = object()
    mod("mpi4py", MPI=dummy)
    mod("petsc4py", PETSc=dummy)
    mod("basix")
    mod("ufl", TestFunction=dummy, TrialFunction=dummy, inner=dummy, grad=dummy, Measure=dummy)
    d = mod("dolfinx", fem=dummy, mesh=dummy)
    mod("dolfinx.mesh", exterior_facet_indices=dummy, create_unit_cube=dummy)
    f = mod("dolfinx.fem")
    mod("dolfinx.fem.petsc", assemble_matrix=dummy, assemble_vector=dummy, apply_lifting=dummy, set_bc=dummy)
    d.fem, f.petsc = f, sys.modules["dolfinx.fem.petsc"]


def main():
    assert os.path.isdir(REF), "the reference tree is needed to (re)generate the fixture"
    stub_imports()
    sys.path.insert(0, REF)
    import cg as ref_cg                  # noqa: E402  reference code
    import chebyshev as ref_cheb         # noqa: E402  reference code

    from oracle import mesh as om, operator as oo
    m = om.create_box(3, 3, 3, perturb=0.2)
    P = 2
    dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
    G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    A = sp.csr_matrix(oo.assemble_csr(P, dm, G, np.full(m.ncells, 2.0), bc, nd))
    A.sort_indices()
    rng = np.random.default_rng(2024)
    b = rng.uniform(-1, 1, nd)
    x0 = rng.uniform(-1, 1, nd)
    out = dict(indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), data=A.data, b=b, x0=x0)

    # ---- Jacobi-CG, 20 iterations, no early exit (python_tests/cg.py:31-59), Lanczos eigenvalues (:61-80)
    s = ref_cg.CGSolver(Mat(A), 20, 0.0, jacobi=True, verbose=False)
    x = vec(np.zeros(nd))
    s.solve(vec(np.ones(nd)), x)
    out["cg_ones_x"] = np.asarray(x)
    out["cg_ones_alphas"] = np.array(s.alphas)
    out["cg_ones_betas"] = np.array(s.betas)
    out["cg_ones_eigs"] = np.sort(np.asarray(s.compute_eigs()))
    s2 = ref_cg.CGSolver(Mat(A), 12, 0.0, jacobi=True, verbose=False)
    x = vec(x0.copy())
    s2.solve(vec(b.copy()), x)
    out["cg_b_x"] = np.asarray(x)
    out["cg_b_alphas"] = np.array(s2.alphas)
    out["cg_b_betas"] = np.array(s2.betas)

    # ---- Chebyshev, 4th kind, Jacobi (python_tests/chebyshev.py:67-91), eig range as in pmg.py:185-201
    lmax = float(out["cg_ones_eigs"][-1])
    out["eig_range"] = np.array([0.1 * lmax, 1.1 * lmax])
    for its in (1, 2, 5):
        for tag, start in (("zero", np.zeros(nd)), ("x0", x0)):
            c = ref_cheb.Chebyshev(Mat(A), its, (0.1 * lmax, 1.1 * lmax), 4, jacobi=True, verbose=False)
            x = vec(start.copy())
            c.solve(vec(b.copy()), x)
            out[f"cheb{its}_{tag}_x"] = np.asarray(x)
    path = os.path.join(ROOT, "tests", "golden", "solvers_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
