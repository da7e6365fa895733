"""CPU prototype (numpy/scipy, NOT product code) for the next round's coarse solver: smoothed-aggregation
AMG as the preconditioner of the P1 coarse PCG, against the Jacobi-PCG that the V-cycle uses today.
It answers two sizing questions for DESIGN.md section 8, item 1: how many PCG iterations to rtol 1e-5 does
each need as the mesh grows, and what does one AMG application cost in fine-level SpMV equivalents.

    python scripts/prototype_sa_amg.py [n ...]       (P1 Laplacian on an n^3 box, Dirichlet boundary)
"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mesh as om, operator as oo  # noqa: E402


def p1_matrix(n):
    m = om.create_box(n, n, n)
    dm, bc, nd = om.dofmap(m, 1), om.bc_marker(m, 1), om.num_dofs(m, 1)
    G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, 1)
    return sp.csr_matrix(oo.assemble_csr(1, dm, G, np.full(m.ncells, 2.0), bc, nd)), bc.astype(bool)


def aggregate(A, free):
    """Greedy (Vanek) aggregation on the graph of A restricted to the free rows."""
    n = A.shape[0]
    agg = -np.ones(n, dtype=np.int64)
    ip, ix = A.indptr, A.indices
    na = 0
    for i in np.flatnonzero(free):                      # pass 1: a node and all its free neighbours
        nb = ix[ip[i]:ip[i + 1]]
        nb = nb[free[nb]]
        if (agg[nb] < 0).all():
            agg[nb] = na
            na += 1
    for i in np.flatnonzero(free & (agg < 0)):          # pass 2: leftovers join a neighbouring aggregate
        nb = ix[ip[i]:ip[i + 1]]
        nb = nb[agg[nb] >= 0]
        if len(nb):
            agg[i] = agg[nb[0]]
        else:
            agg[i] = na
            na += 1
    return agg, na


def lam_max(A, dinv, its=15):
    x = np.random.default_rng(0).standard_normal(A.shape[0])
    for _ in range(its):
        x = dinv * (A @ x)
        lam = np.linalg.norm(x)
        x /= lam
    return lam


def build(A, free, levels=None, min_size=400):
    levels = levels if levels is not None else []
    dinv = 1.0 / A.diagonal()
    lmax = 1.1 * lam_max(A, dinv)
    lev = dict(A=A, dinv=dinv, lmax=lmax)
    levels.append(lev)
    if free.sum() <= min_size or len(levels) > 8:
        lev["dense"] = np.linalg.inv(A.toarray())
        return levels
    agg, na = aggregate(A, free)
    rows = np.flatnonzero(agg >= 0)
    T = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(A.shape[0], na))
    P = (T - sp.diags((4.0 / (3.0 * lmax)) * dinv) @ (A @ T)).tocsr()   # smoothed prolongator
    lev["P"] = P
    Ac = (P.T @ A @ P).tocsr()
    return build(Ac, np.ones(na, dtype=bool), levels, min_size)


def cheb(lev, x, b, its):
    """4th-kind Chebyshev / Jacobi, the smoother the library already has."""
    A, dinv, lmax = lev["A"], lev["dinv"], lev["lmax"]
    r = b - A @ x if x is not None else b.copy()
    x = np.zeros_like(b) if x is None else x
    z = r * dinv * (4.0 / (3.0 * lmax))
    for i in range(1, its + 1):
        x = x + z
        if i == its:
            break
        r = r - A @ z
        z = z * ((2 * i - 1) / (2 * i + 3)) + ((8 * i + 4) / (2 * i + 3) / lmax) * (r * dinv)
    return x


def vcycle(levels, li, b, nu=2):
    lev = levels[li]
    if "dense" in lev:
        return lev["dense"] @ b
    x = cheb(lev, None, b, nu)
    r = b - lev["A"] @ x
    x = x + lev["P"] @ vcycle(levels, li + 1, lev["P"].T @ r, nu)
    return cheb(lev, x, b, nu)


def pcg(A, b, M, rtol, maxit):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = r @ z
    rz0 = rz
    for k in range(1, maxit + 1):
        Ap = A @ p
        a = rz / (p @ Ap)
        x += a * p
        r -= a * Ap
        z = M(r)
        rz_new = r @ z
        if rz_new / rz0 < rtol * rtol:
            return x, k
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, maxit


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 24, 32, 48]
    print("   n     dofs  Jacobi-PCG its |  SA-AMG PCG its  levels  operator cx  SpMV-equiv per application")
    for n in sizes:
        A, bc = p1_matrix(n)
        b = np.random.default_rng(1).uniform(-1, 1, A.shape[0]) * (~bc)
        dinv = 1.0 / A.diagonal()
        _, kj = pcg(A, b, lambda r: dinv * r, 1e-5, 2000)
        t0 = time.time()
        levels = build(A, ~bc)
        nnz = [l["A"].nnz for l in levels]
        cx = sum(nnz) / nnz[0]
        # one V(2,2): per level 2*(nu - 1) smoother SpMVs + 1 residual + P, P^T (~ nnz(P)/nnz(A) each)
        work = sum((2 * 1 + 1) * l["A"].nnz + (2 * l["P"].nnz if "P" in l else 0) for l in levels) / nnz[0]
        _, ka = pcg(A, b, lambda r: vcycle(levels, 0, r), 1e-5, 200)
        print(f"{n:4d} {A.shape[0]:8d} {kj:12d}    | {ka:10d} {len(levels):9d} {cx:12.2f} {work:14.1f}"
              f"   (setup {time.time() - t0:.1f} s)")


if __name__ == "__main__":
    main()
