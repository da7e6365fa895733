"""Matrix-free hex Laplacian, diagonal, assembled CSR, transfer and RHS (oracle).

TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, vectorised over cells) of:
  * ``geometry_computation<T,P>``     src/laplacian.hpp:22-113   (G = w K K^T / detJ)
    host twin ``compute_scaled_geometrical_factor``  src/precompute.hpp:108-251
  * ``stiffness_operator<T,P>``       src/laplacian.hpp:143-278  (y += k D^T G D x, BC rows y=x)
  * assembled operator + BC diag=1    src/csr.hpp:66-112 (``set_diagonal`` :86), diag^-1 :101-112
  * ``interpolate_Q1Q2/Q2Q1``         src/interpolate.hpp:21-87,117-178
  * CSR SpMV                          src/csr.hpp:20-36
  * RHS with GLL collocation          examples/pmg/poisson.py:6-40, examples/pmg/main.cpp:289-295
"""
import numpy as np
import scipy.sparse as sp

from . import gll
from . import mesh as omesh


# ---------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------
def trilinear_dphi(P):
    """_dphi[d, q, k] at the (P+1)^3 GLL points, q = ix*nq^2+iy*nq+iz, k = (a*2+b)*2+c
    (layout of ``dphi_geometry`` consumed at src/laplacian.hpp:79)."""
    x1, _ = gll.gll_points_weights(P + 1)
    X, Y, Z = np.meshgrid(x1, x1, x1, indexing="ij")
    X, Y, Z = X.reshape(-1), Y.reshape(-1), Z.reshape(-1)
    nq = len(X)
    d = np.zeros((3, nq, 8))
    for a in range(2):
        for b in range(2):
            for c in range(2):
                k = (a * 2 + b) * 2 + c
                la, lb, lc = (X if a else 1 - X), (Y if b else 1 - Y), (Z if c else 1 - Z)
                da, db, dc = (1.0 if a else -1.0), (1.0 if b else -1.0), (1.0 if c else -1.0)
                d[0, :, k] = da * lb * lc
                d[1, :, k] = la * db * lc
                d[2, :, k] = la * lb * dc
    return d


def weights_3d(P):
    _, w1 = gll.gll_points_weights(P + 1)
    return np.einsum("i,j,k->ijk", w1, w1, w1).reshape(-1)


def geometry_factors(verts, geom_dofmap, P, literal_detj=False, cells=None):
    """G[c, q, 6] = (xx, xy, xz, yy, yz, zz) of w_q * K K^T / detJ, K = adj(J).

    ``literal_detj=True`` reproduces the reference's expression
    ``J00*K00 - J10*K01 + J02*K20`` (src/laplacian.hpp:97), which equals det J only
    when J01*J10*J22 terms vanish (true on the axis-aligned cube); the default is the
    exact determinant (quirk Q17 in DESIGN.md).  Also returns detJ[c, q].
    """
    gd = geom_dofmap if cells is None else geom_dofmap[cells]
    cv = verts[gd]                                   # [nc, 8, 3]
    dphi = trilinear_dphi(P)                         # [3, nq, 8]
    w = weights_3d(P)
    # J[c,q,i,j] = sum_k coord(k,i) * dphi(j,q,k)      (laplacian.hpp:81-87)
    J = np.einsum("cki,jqk->cqij", cv, dphi)
    K = np.empty_like(J)
    K[..., 0, 0] = J[..., 1, 1] * J[..., 2, 2] - J[..., 1, 2] * J[..., 2, 1]
    K[..., 0, 1] = -J[..., 0, 1] * J[..., 2, 2] + J[..., 0, 2] * J[..., 2, 1]
    K[..., 0, 2] = J[..., 0, 1] * J[..., 1, 2] - J[..., 0, 2] * J[..., 1, 1]
    K[..., 1, 0] = -J[..., 1, 0] * J[..., 2, 2] + J[..., 1, 2] * J[..., 2, 0]
    K[..., 1, 1] = J[..., 0, 0] * J[..., 2, 2] - J[..., 0, 2] * J[..., 2, 0]
    K[..., 1, 2] = -J[..., 0, 0] * J[..., 1, 2] + J[..., 0, 2] * J[..., 1, 0]
    K[..., 2, 0] = J[..., 1, 0] * J[..., 2, 1] - J[..., 1, 1] * J[..., 2, 0]
    K[..., 2, 1] = -J[..., 0, 0] * J[..., 2, 1] + J[..., 0, 1] * J[..., 2, 0]
    K[..., 2, 2] = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    if literal_detj:
        detJ = J[..., 0, 0] * K[..., 0, 0] - J[..., 1, 0] * K[..., 0, 1] + J[..., 0, 2] * K[..., 2, 0]
    else:
        detJ = J[..., 0, 0] * K[..., 0, 0] + J[..., 0, 1] * K[..., 1, 0] + J[..., 0, 2] * K[..., 2, 0]
    KKt = np.einsum("cqik,cqjk->cqij", K, K)
    s = (w[None, :] / detJ)
    G = np.stack([KKt[..., 0, 0], KKt[..., 1, 0], KKt[..., 2, 0],
                  KKt[..., 1, 1], KKt[..., 2, 1], KKt[..., 2, 2]], axis=-1) * s[..., None]
    return G, detJ


# ---------------------------------------------------------------------------
# matrix-free apply (src/laplacian.hpp:182-277)
# ---------------------------------------------------------------------------
def apply_cells(P, dofmap, G, kappa, bc, x, y, cells=None):
    """y (+)= contributions of ``cells`` (all if None); BC rows get y = x (last writer)."""
    nd = P + 1
    _, _, D = gll.tables(P)
    dm = dofmap if cells is None else dofmap[cells]
    Gc = G if cells is None else G[cells]
    kc = kappa if cells is None else kappa[cells]
    if dm.shape[0] == 0:
        return y
    isbc = bc[dm] != 0
    u = np.where(isbc, 0.0, x[dm]).reshape(-1, nd, nd, nd)          # :186-189
    vx = np.einsum("qi,cijk->cqjk", D, u)                           # :195-199
    vy = np.einsum("qj,cijk->ciqk", D, u)                           # :206-210
    vz = np.einsum("qk,cijk->cijq", D, u)                           # :214-218
    Gr = Gc.reshape(-1, nd, nd, nd, 6)
    k4 = kc.reshape(-1, 1, 1, 1)
    fx = k4 * (Gr[..., 0] * vx + Gr[..., 1] * vy + Gr[..., 2] * vz)  # :233-235
    fy = k4 * (Gr[..., 1] * vx + Gr[..., 3] * vy + Gr[..., 4] * vz)
    fz = k4 * (Gr[..., 2] * vx + Gr[..., 4] * vy + Gr[..., 5] * vz)
    out = (np.einsum("qi,cqjk->cijk", D, fx) + np.einsum("qj,ciqk->cijk", D, fy)
           + np.einsum("qk,cijq->cijk", D, fz)).reshape(dm.shape)   # :246-270
    np.add.at(y, dm[~isbc], out[~isbc])                             # :277
    y[dm[isbc]] = x[dm[isbc]]                                       # :273-274
    return y


def apply(P, dofmap, G, kappa, bc, x):
    """y = A x incl. the zero fill of src/laplacian.hpp:466."""
    y = np.zeros_like(x)
    return apply_cells(P, dofmap, G, kappa, bc, x, y)


def element_matrices(P, G, kappa):
    """A_e[c, i, j] = kappa * sum_q B[q,i,d] G[q,d,e] B[q,j,e] with the collocated
    tensor-product gradient table B (phi is the identity at the GLL points)."""
    nd = P + 1
    _, _, D = gll.tables(P)
    I = np.eye(nd)
    n3 = nd ** 3
    B = np.zeros((3, n3, n3))
    B[0] = np.einsum("qi,rj,sk->qrsijk", D, I, I).reshape(n3, n3)
    B[1] = np.einsum("qi,rj,sk->qrsijk", I, D, I).reshape(n3, n3)
    B[2] = np.einsum("qi,rj,sk->qrsijk", I, I, D).reshape(n3, n3)
    idx = [(0, 0, 0), (0, 1, 1), (0, 2, 2), (1, 1, 3), (1, 2, 4), (2, 2, 5)]
    Ae = np.zeros((G.shape[0], n3, n3))
    for (d, e, g) in idx:
        t = np.einsum("qi,cq,qj->cij", B[d], G[:, :, g], B[e], optimize=True)
        Ae += t
        if d != e:
            Ae += t.transpose(0, 2, 1)
    return Ae * kappa[:, None, None]


def assemble_csr(P, dofmap, G, kappa, bc, ndofs):
    """Assembled operator with BC rows/cols zeroed and diagonal 1 (src/csr.hpp:84-86)."""
    Ae = element_matrices(P, G, kappa)
    n3 = dofmap.shape[1]
    rows = np.repeat(dofmap, n3, axis=1).reshape(-1)
    cols = np.tile(dofmap, (1, n3)).reshape(-1)
    vals = Ae.reshape(-1).copy()
    keep = (bc[rows] == 0) & (bc[cols] == 0)
    A = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(ndofs, ndofs)).tocsr()
    A = A + sp.diags((bc != 0).astype(np.float64))
    A.sum_duplicates()
    A.sort_indices()
    return A.tocsr()


def diagonal(P, dofmap, G, kappa, bc, ndofs):
    """diag(A) of the assembled operator, 1 at BC dofs (src/csr.hpp:86,101-112)."""
    nd = P + 1
    _, _, D = gll.tables(P)
    D2 = D * D                                                  # [q, i]
    Gr = G.reshape(-1, nd, nd, nd, 6)
    # only same-direction terms survive on the diagonal (collocation): sum_q D[q,i]^2 G_dd(q,...)
    de = (np.einsum("qi,cqjk->cijk", D2, Gr[..., 0]) + np.einsum("qj,ciqk->cijk", D2, Gr[..., 3])
          + np.einsum("qk,cijq->cijk", D2, Gr[..., 5]))
    # cross terms: 2*G_xy*D[i,i]*D[j,j] etc. at q == dof
    Dd = np.diag(D)
    de += 2.0 * (Gr[..., 1] * Dd[:, None, None] * Dd[None, :, None]
                 + Gr[..., 2] * Dd[:, None, None] * Dd[None, None, :]
                 + Gr[..., 4] * Dd[None, :, None] * Dd[None, None, :])
    de = de.reshape(dofmap.shape) * kappa[:, None]
    d = np.zeros(ndofs)
    np.add.at(d, dofmap.reshape(-1), de.reshape(-1))
    d[bc != 0] = 1.0
    return d


def spmv(A, x):
    """y = A x, scalar-row CSR (src/csr.hpp:20-36)."""
    return A @ x


# ---------------------------------------------------------------------------
# p-transfer (src/interpolate.hpp)
# ---------------------------------------------------------------------------
def local_interp_matrix(Pc, Pf):
    """M[j_f, k_c] = kron(M1d, M1d, M1d); entries |v|<=1e-12 dropped (interpolate.hpp:120-135)."""
    M1 = gll.interp_1d(Pc, Pf)
    M = np.einsum("ai,bj,ck->abcijk", M1, M1, M1).reshape((Pf + 1) ** 3, (Pc + 1) ** 3)
    M[np.abs(M) <= 1e-12] = 0.0
    return M


def prolong(Pc, Pf, dm_c, dm_f, xc, nf, cells=None):
    """fine[dofs2[j]] = sum_k M[j,k] coarse[dofs1[k]] (overwrite; interpolate.hpp:21-45)."""
    M = local_interp_matrix(Pc, Pf)
    a = dm_c if cells is None else dm_c[cells]
    b = dm_f if cells is None else dm_f[cells]
    xf = np.zeros(nf)
    xf[b.reshape(-1)] = (xc[a] @ M.T).reshape(-1)
    return xf


def multiplicity(dm_f, nf):
    """Q2mult over all local (+ghost) cells (interpolate.hpp:172-178)."""
    m = np.zeros(nf)
    np.add.at(m, dm_f.reshape(-1), 1.0)
    return m


def restrict(Pc, Pf, dm_c, dm_f, xf, nc, mult=None):
    """coarse[dofs1[j]] += sum_k M^T[j,k] fine[d]/mult[d]; output zeroed first
    (interpolate.hpp:60-87,270)."""
    M = local_interp_matrix(Pc, Pf)
    if mult is None:
        mult = multiplicity(dm_f, len(xf))
    vals = (xf / np.where(mult > 0, mult, 1.0))[dm_f] @ M          # [nc_cells, ndc]
    xc = np.zeros(nc)
    np.add.at(xc, dm_c.reshape(-1), vals.reshape(-1))
    return xc


# ---------------------------------------------------------------------------
# RHS (lumped GLL collocation; examples/pmg/poisson.py:30-40, main.cpp:289-295)
# ---------------------------------------------------------------------------
def rhs_collocated(mesh, P, f, bc, g=0.0, kappa=None, A=None):
    """fem::assemble_vector + apply_lifting + set_bc with the GLL rule (examples/pmg/main.cpp:289-295,
    examples/cg/main.cpp:234-236):  b_i = sum_{K contains i} f(x_i) w_i |detJ_K(x_i)|, then the lifting
    b -= A_full g_bc (A_full: the operator WITHOUT Dirichlet rows/columns, g_bc = g at BC dofs, 0 elsewhere;
    DOLFINx apply_lifting with scale 1 and x0 = 0), then ``set_bc`` writes b = g at BC dofs.  ``g`` is a
    scalar (the reference's constant 1.3, examples/cg/main.cpp:158) or one value per dof; ``kappa`` (scalar or
    per cell) is only needed for g != 0 (default 2.0, examples/pmg/main.cpp:79)."""
    dm = omesh.dofmap(mesh, P)
    X = omesh.dof_coords(mesh, P)
    G, detJ = geometry_factors(mesh.verts, mesh.geom_dofmap, P)
    w = weights_3d(P)
    b = np.zeros(X.shape[0])
    fv = f(X)                                                       # [ndofs]
    np.add.at(b, dm.reshape(-1), (fv[dm] * (w[None, :] * np.abs(detJ))).reshape(-1))
    gv = np.full(X.shape[0], g, dtype=np.float64) if np.isscalar(g) else np.asarray(g, dtype=np.float64)
    if np.any(gv[bc != 0] != 0.0):
        kap = 2.0 if kappa is None else kappa
        kap = np.full(mesh.ncells, kap) if np.isscalar(kap) else np.asarray(kap, dtype=np.float64)
        gbc = np.where(bc != 0, gv, 0.0)
        b -= apply(P, dm, G, kap, np.zeros_like(bc), gbc)           # A_full g_bc
    b[bc != 0] = gv[bc != 0]
    return b


def f_sines(kx, ky, kz, kappa):
    """-kappa*div(grad(sin sin sin)) (examples/pmg/poisson.py:6-30, python_tests/pmg.py:70)."""
    def f(X):
        return (kappa * np.pi ** 2 * (kx * kx + ky * ky + kz * kz)
                * np.sin(kx * np.pi * X[:, 0]) * np.sin(ky * np.pi * X[:, 1]) * np.sin(kz * np.pi * X[:, 2]))
    return f
