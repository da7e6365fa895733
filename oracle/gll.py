"""1-D Gauss-Lobatto-Legendre tables (oracle; test infrastructure only).

Restates what the reference obtains from Basix (absent here; unpinned version
``fenics-basix@main``, hypre-cuda.yaml:48-49):
  * ``basix::quadrature::make_quadrature(gll, interval, m)``   src/laplacian.hpp:307-309
  * ``basix::create_element(P, interval, degree, gll_warped)`` + ``tabulate(1, pts)``
                                                               src/laplacian.hpp:302-317
  * ``basix::compute_interpolation_operator(Q1, Q2)``          src/interpolate.hpp:118
The GLL rule with n points and the nodal Lagrange basis on the same n points are
uniquely defined mathematically, so the restatement is exact up to rounding.
All tables use ascending node order on [0, 1].
"""
import numpy as np
from numpy.polynomial import legendre as L


def gll_points_weights(n):
    """n-point GLL rule on [0, 1]; weights sum to 1."""
    if n < 2:
        raise ValueError("GLL needs >= 2 points")
    N = n - 1
    c = np.zeros(N + 1)
    c[N] = 1.0
    if n == 2:
        t = np.array([-1.0, 1.0])
    else:
        dc = L.legder(c)
        t = np.sort(np.real(L.legroots(dc)))
        # Newton polish on P'_N
        ddc = L.legder(dc)
        for _ in range(4):
            t = t - L.legval(t, dc) / L.legval(t, ddc)
        t = np.concatenate([[-1.0], t, [1.0]])
    # symmetrise
    t = 0.5 * (t - t[::-1])
    w = 2.0 / (N * (N + 1) * L.legval(t, c) ** 2)
    return 0.5 * (t + 1.0), 0.5 * w


def _bary_weights(x):
    n = len(x)
    w = np.ones(n)
    for i in range(n):
        for j in range(n):
            if i != j:
                w[i] /= x[i] - x[j]
    return w


def lagrange_eval_matrix(xnodes, xeval):
    """M[e, i] = l_i(xeval[e]) for the Lagrange basis on ``xnodes``."""
    xn = np.asarray(xnodes, dtype=float)
    xe = np.asarray(xeval, dtype=float)
    bw = _bary_weights(xn)
    M = np.zeros((len(xe), len(xn)))
    for e, xv in enumerate(xe):
        d = xv - xn
        hit = np.where(np.abs(d) < 1e-14)[0]
        if len(hit):
            M[e, hit[0]] = 1.0
        else:
            t = bw / d
            M[e] = t / t.sum()
    return M


def lagrange_deriv_matrix(xnodes):
    """D[q, i] = l_i'(x_q) on the nodes themselves (src/laplacian.hpp:198: dphi[q*nd+i])."""
    x = np.asarray(xnodes, dtype=float)
    n = len(x)
    bw = _bary_weights(x)
    D = np.zeros((n, n))
    for q in range(n):
        for i in range(n):
            if i != q:
                D[q, i] = (bw[i] / bw[q]) / (x[q] - x[i])
        D[q, q] = -np.sum(D[q, np.arange(n) != q])
    return D


def tables(P):
    """(points, weights, D) for degree P: P+1 GLL points, nodes == points."""
    x, w = gll_points_weights(P + 1)
    return x, w, lagrange_deriv_matrix(x)


def interp_1d(Pc, Pf):
    """M1d[i_f, i_c] = l^c_{i_c}(x^f_{i_f}) (src/interpolate.hpp:118)."""
    xc, _ = gll_points_weights(Pc + 1)
    xf, _ = gll_points_weights(Pf + 1)
    return lagrange_eval_matrix(xc, xf)
