"""Jacobi-PCG, tqli, 4th-kind Chebyshev and the p-multigrid V-cycle (oracle).

TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's host algorithms;
operators are passed as callables ``A(x) -> y`` and ``dinv`` arrays so the same
code runs on the matrix-free oracle, on scipy CSR, or on per-rank emulations.
  * tqli / tqli_ml               src/cg.hpp:15-84   == python_tests/tqli.py:7-60
  * CGSolver::solve              src/cg.hpp:147-222 (C++ semantics: break BEFORE p update
                                 and BEFORE storing alpha/beta, test rnorm/rnorm0 < rtol^2)
  * CGSolver::compute_eigenvalues src/cg.hpp:121-142
  * Chebyshev::solve             src/chebyshev.hpp:46-91 == python_tests/chebyshev.py:67-91
  * MultigridPreconditioner::apply src/pmg.hpp:56-155 == python_tests/pmg.py:217-289
"""
import math
import numpy as np


def tqli(d, e):
    """In-place eigenvalues of the symmetric tridiagonal (d, e); returns 0 / -1."""
    n = len(d)

    def find_m(l):
        for m in range(l, n - 1):
            dd = abs(d[m]) + abs(d[m + 1])
            if abs(e[m]) + dd == dd:
                return m
        return n - 1

    for l in range(n):
        it = 0
        while True:
            m = find_m(l)
            if m == l:
                break
            if it == 30:
                return -1
            it += 1
            g = (d[l + 1] - d[l]) / (2.0 * e[l])
            r = math.sqrt(g * g + 1.0)
            g = d[m] - d[l] + e[l] / (g + r if g >= 0 else g - r)
            s = c = 1.0
            p = 0.0
            early = False
            for i in range(m - 1, l - 1, -1):
                f = s * e[i]
                b = c * e[i]
                r = math.sqrt(f * f + g * g)
                e[i + 1] = r
                if r == 0.0:
                    d[i + 1] -= p
                    e[m] = 0.0
                    early = True
                    break
                s = f / r
                c = g / r
                g = d[i + 1] - p
                r = (d[i] - g) * s + 2.0 * c * b
                p = s * r
                d[i + 1] = g + p
                g = c * r - b
            if not early:
                d[l] -= p
                e[l] = g
                e[m] = 0.0
        e[l] = 0.0
    return 0


def cg(A, dinv, x, b, max_iter, rtol, dot=np.dot, M=None):
    """Returns (x, iters, alphas, betas, rnorms, rnorm0).  ``rnorms`` holds the value of
    r.M^-1 r after every iteration (also those not stored by the reference).  ``M`` (callable
    r -> M^-1 r) stands in the two places where the reference multiplies by diag^-1
    (src/cg.hpp:162,192): the p-multigrid V-cycle as the preconditioner."""
    if M is None:
        M = lambda v: v * dinv
    y = A(x)
    r = b - y
    p = M(r)
    rnorm0 = dot(p, r)
    rnorm = rnorm0
    rtol2 = rtol * rtol
    alphas, betas, hist = [], [], []
    k = 0
    x = x.copy()
    while k < max_iter:
        k += 1
        y = A(p)
        alpha = rnorm / dot(p, y)
        x = x + alpha * p
        r = r - alpha * y
        y = M(r)
        rnorm_new = dot(r, y)
        beta = rnorm_new / rnorm
        rnorm = rnorm_new
        hist.append(rnorm)
        if rnorm / rnorm0 < rtol2:
            break
        p = beta * p + y
        alphas.append(alpha)
        betas.append(beta)
    return x, k, np.array(alphas), np.array(betas), np.array(hist), rnorm0


def lanczos_eigenvalues(alphas, betas):
    ne = len(alphas)
    if ne < 2:
        raise RuntimeError("Insufficient data to compute eigenvalues")
    d = np.zeros(ne)
    e = np.zeros(ne)
    for i in range(ne):
        d[i] = 1.0 / alphas[i]
    for i in range(ne - 1):
        d[i + 1] += betas[i] / alphas[i]
        e[i] = math.sqrt(betas[i]) / alphas[i]
    if tqli(d, e) == -1:
        raise RuntimeError("Eigenvalue estimate failed")
    return np.sort(d)


def chebyshev(A, dinv, x, b, max_iter, lmax, norm=np.linalg.norm, history=None):
    """4th-kind Chebyshev with Jacobi; only lmax = eig_range[1] is used (chebyshev.hpp:51)."""
    x = x.copy()
    r = b - A(x)
    if history is not None:
        history.append(norm(r))
    z = r * dinv * (4.0 / (3.0 * lmax))
    for i in range(1, max_iter + 1):
        x = x + z
        r = r - A(z)
        z = z * (float(2 * i - 1) / float(2 * i + 3))
        z = z + (float(8 * i + 4) / float(2 * i + 3) / lmax) * (r * dinv)
        if history is not None:
            history.append(norm(r))
    return x


class Level:
    """One p-level: operator callable, D^-1, smoother settings, BC marker."""
    def __init__(self, A, dinv, bc, lmax, nsmooth=2):
        self.A, self.dinv, self.bc, self.lmax, self.nsmooth = A, dinv, bc, lmax, nsmooth


def vcycle(levels, prolongs, restricts, b_top, u_top, coarse_solve=None, mask_all_levels=True,
           history=None, norm=np.linalg.norm):
    """One ``MultigridPreconditioner::apply``.  ``prolongs[i]``: level i -> i+1,
    ``restricts[i]``: level i+1 -> i.  ``mask_all_levels=False`` is the literal reference
    (b masked on level 0 only, src/pmg.hpp:100-103, quirk Q9); identical for two levels."""
    nl = len(levels)
    u = [None] * nl
    b = [None] * nl
    u[-1] = u_top.copy()
    b[-1] = b_top.copy()
    for i in range(nl - 1, 0, -1):
        L = levels[i]
        if u[i] is None:
            u[i] = np.zeros_like(b[i])
        if history is not None:
            history.append(("pre", i, norm(b[i] - L.A(u[i]))))
        u[i] = chebyshev(L.A, L.dinv, u[i], b[i], L.nsmooth, L.lmax)
        r = b[i] - L.A(u[i])
        if history is not None:
            history.append(("pre_smoothed", i, norm(r)))
        b[i - 1] = restricts[i - 1](r)
        if mask_all_levels or i - 1 == 0:
            b[i - 1] = b[i - 1] * (1 - levels[i - 1].bc)
    L0 = levels[0]
    u[0] = np.zeros_like(b[0])
    if coarse_solve is not None:
        u[0] = coarse_solve(u[0], b[0])
    else:
        u[0] = chebyshev(L0.A, L0.dinv, u[0], b[0], L0.nsmooth, L0.lmax)
    if history is not None:
        history.append(("coarse", 0, norm(b[0] - L0.A(u[0]))))
    for i in range(nl - 1):
        L = levels[i + 1]
        u[i + 1] = u[i + 1] + prolongs[i](u[i])
        if history is not None:
            history.append(("corrected", i + 1, norm(b[i + 1] - L.A(u[i + 1]))))
        u[i + 1] = chebyshev(L.A, L.dinv, u[i + 1], b[i + 1], L.nsmooth, L.lmax)
        if history is not None:
            history.append(("post", i + 1, norm(b[i + 1] - L.A(u[i + 1]))))
    return u[-1]
