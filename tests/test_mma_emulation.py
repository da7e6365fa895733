"""CPU: lane-level emulation of the FP64 tensor-core apply kernel (csrc/laplacian_mma.cuh, k_apply_affine_mma).

The kernel spreads one element over a warp in the fragment layouts of mma.sync.m8n8k4.f64 (A: lane holds
[lane/4][lane%4], B: [lane%4][lane/4], C: [lane/4][2(lane%4) + {0,1}]) and exchanges data between the three
directions only through swizzled shared-memory arrays.  This test runs exactly that index logic in numpy --
32 "lanes" as array elements, the mma as a small matrix product in the documented layout, the same swizzled
addresses (and a count of the shared-memory bank conflicts of every access) --
on a sheared (affine) cell against the oracle, so that an indexing error shows up without a GPU.  The CUDA
kernel itself is checked against the oracle by tests/test_gpu_operator.py (P6, P7 on affine and perturbed meshes).
The second test restates the streamed-G variant as one warp of a launch runs it (cell loop, Dirichlet handling,
G in layout Y, dof indices one cell ahead) and drives a whole mesh through it."""
import numpy as np
import pytest

from oracle import gll, mesh as om, operator as oo

SI = 72
lane = np.arange(32)
r, c, q = lane // 4, lane % 4, lane % 4
WAVEFRONTS = {"ideal": 0, "actual": 0}


def at(i, j, k):
    """MmaCfg::at: i-stride 72, j-stride 8, k XOR-swizzled with 4 ((i/2 + j/2) mod 2)"""
    return i * SI + j * 8 + (k ^ ((((i >> 1) ^ (j >> 1)) & 1) << 2))


def _count(addr, width):
    """shared-memory wavefronts of one warp access: 8-byte accesses go by half-warps over 16 8-byte banks,
    16-byte accesses by quarter-warps over 8 16-byte banks; same address = broadcast"""
    per = 16 if width == 8 else 8
    unit = addr if width == 8 else addr // 2
    for g in range(0, 32, per):
        u = np.unique(unit[g:g + per])
        banks = u % per
        WAVEFRONTS["ideal"] += 1
        WAVEFRONTS["actual"] += int(np.bincount(banks, minlength=per).max())


def lds(buf, addr):
    _count(addr, 8)
    return buf[addr]


def lds2(buf, addr):
    assert np.all(addr % 2 == 0)
    _count(addr, 16)
    return buf[addr], buf[addr + 1]


def sts2(buf, addr, v0, v1):
    assert np.all(addr % 2 == 0)
    _count(addr, 16)
    buf[addr], buf[addr + 1] = v0, v1


def mma(a, b, c0, c1):
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[lane // 4, lane % 4] = a
    B[lane % 4, lane // 4] = b
    C = A @ B
    return c0 + C[lane // 4, 2 * (lane % 4)], c1 + C[lane // 4, 2 * (lane % 4) + 1]


def _contract_cell(P, B0, B1, B2, flux):
    """forward contractions, re-layout of gx, flux(t, gx, gy, gz) -> (fx, fy, fz) pairs at the Y points, transposed
    contractions, re-layout of the X accumulator; U is in B0 on entry; returns the Y-layout result [t][e][lane]"""
    n = P + 1
    _, _, D1 = gll.tables(P)              # D1[q, i] = l_i'(x_q)
    D = np.zeros((8, 8)); D[:n, :n] = D1
    dA0, dA1 = D[r, c], D[r, c + 4]            # A = D (x, y forward); also B = D^T (z forward)
    tA0, tA1 = D[c, r], D[c + 4, r]            # A = D^T (x, y backward); also B = D (z backward)
    gz = np.zeros((8, 2, 32)); gy = np.zeros((8, 2, 32)); gx = np.zeros((8, 2, 32))
    z0 = np.zeros(32)
    for t in range(n):
        c0, c1 = mma(lds(B0, at(t, r, c)), dA0, z0, z0)                 # z: A = U[t][r][m]
        gz[t] = mma(lds(B0, at(t, r, c + 4)), dA1, c0, c1)
        c0, c1 = mma(dA0, lds(B0, at(t, c, r)), z0, z0)                 # y: B[m][k] = U[t][m][k]
        gy[t] = mma(dA1, lds(B0, at(t, c + 4, r)), c0, c1)
        c0, c1 = mma(dA0, lds(B0, at(c, t, r)), z0, z0)                 # x (tile j = t): B[m][k] = U[m][t][k]
        gx[t] = mma(dA1, lds(B0, at(c + 4, t, r)), c0, c1)
    for t in range(n):                                                  # gx: layout X -> shared memory
        sts2(B0, at(r, t, 2 * c), gx[t, 0], gx[t, 1])
    for t in range(n):                                                  # flux at the Y points
        o = at(t, r, 2 * c)
        f = flux(t, lds2(B0, o), gy[t], gz[t])
        sts2(B0, o, f[0][0], f[0][1])
        sts2(B1, o, f[1][0], f[1][1])
        sts2(B2, o, f[2][0], f[2][1])
    ayz = np.zeros((8, 2, 32)); ax = np.zeros((8, 2, 32))
    for t in range(n):
        c0, c1 = mma(lds(B2, at(t, r, c)), tA0, z0, z0)                 # z': A[j][q'] = fz[t][j][q'], B = D
        c0, c1 = mma(lds(B2, at(t, r, c + 4)), tA1, c0, c1)
        c0, c1 = mma(tA0, lds(B1, at(t, c, r)), c0, c1)                 # y': A = D^T, B[q'][k] = fy[t][q'][k]
        ayz[t] = mma(tA1, lds(B1, at(t, c + 4, r)), c0, c1)
        d0, d1 = mma(tA0, lds(B0, at(c, t, r)), z0, z0)                 # x' (tile j = t): B[q'][k] = fx[q'][t][k]
        ax[t] = mma(tA1, lds(B0, at(c + 4, t, r)), d0, d1)
    for t in range(n):
        sts2(B1, at(r, t, 2 * c), ax[t, 0], ax[t, 1])
    out = np.zeros((8, 2, 32))
    for t in range(n):
        a0, a1 = lds2(B1, at(t, r, 2 * c))
        out[t, 0], out[t, 1] = ayz[t, 0] + a0, ayz[t, 1] + a1
    return out


def apply_cell(P, u_cell, Gc, kappa):
    """affine variant on one cell: u_cell[n,n,n] (BC already zeroed) -> acc[n,n,n]"""
    n = P + 1
    _, w1, _ = gll.tables(P)
    w = np.zeros(8); w[:n] = w1
    B0 = np.zeros(8 * SI); B1 = np.zeros(8 * SI); B2 = np.zeros(8 * SI)
    up = np.zeros((8, 8, 8)); up[:n, :n, :n] = u_cell
    for t in range(n):                    # gather in layout Y: lane (r, c) holds (t, r, 2c), (t, r, 2c + 1)
        sts2(B0, at(t, r, 2 * c), up[t, r, 2 * c], up[t, r, 2 * c + 1])
    G00, G01, G02, G11, G12, G22 = Gc

    def flux(t, gxy, gy, gz):
        f = np.zeros((3, 2, 32))
        for e in range(2):
            ww = kappa * w[t] * w[r] * w[2 * c + e]
            f[0, e] = ww * (G00 * gxy[e] + G01 * gy[e] + G02 * gz[e])
            f[1, e] = ww * (G01 * gxy[e] + G11 * gy[e] + G12 * gz[e])
            f[2, e] = ww * (G02 * gxy[e] + G12 * gy[e] + G22 * gz[e])
        return f

    res = _contract_cell(P, B0, B1, B2, flux)
    acc = np.zeros((8, 8, 8))
    for t in range(n):
        acc[t, r, 2 * c] = res[t, 0]
        acc[t, r, 2 * c + 1] = res[t, 1]
    assert np.all(acc[n:] == 0) and np.all(acc[:, n:] == 0) and np.all(acc[:, :, n:] == 0)
    return acc[:n, :n, :n]


def warp_kernel(P, enc, G, kappa, x, y, first, count, gw, nw, ahead):
    """streamed-G variant as ONE warp of the launch runs it: the cell loop with stride nw, dof indices loaded in
    layout Y (one cell ahead if `ahead`), Dirichlet columns zeroed in the gather, G[p][6][n3] read in layout Y
    tile by tile, atomics / Dirichlet rows in the scatter.  enc, G flat as on the device."""
    n = P + 1
    n3 = n ** 3
    v = [(r < n) & (2 * c + e < n) for e in range(2)]          # the lane's two points exist (n = 7 pads)

    def load_idx(pl):
        e0 = (first + pl) * n3
        d = -np.ones((8, 2, 32), dtype=np.int64)
        for t in range(n):
            a = t * n * n + r * n + 2 * c                       # n = 8: 64 t + 2 lane, the linear order
            for e in range(2):
                d[t, e, v[e]] = enc[e0 + a[v[e]] + e]
        return d

    B0 = np.zeros(8 * SI); B1 = np.zeros(8 * SI); B2 = np.zeros(8 * SI)
    d = load_idx(gw) if ahead and gw < count else None
    for pl in range(gw, count, nw):
        p = first + pl
        if not ahead:
            d = load_idx(pl)
        for t in range(n):
            xv = [np.where(d[t, e] >= 0, x[np.maximum(d[t, e], 0)], 0.0) for e in range(2)]
            sts2(B0, at(t, r, 2 * c), xv[0], xv[1])
        Gq = p * 6 * n3 + r * n + 2 * c

        def flux(t, gxy, gy, gz):
            g = np.zeros((6, 2, 32))
            for cc in range(6):
                for e in range(2):
                    g[cc, e, v[e]] = G[Gq[v[e]] + cc * n3 + t * n * n + e]
            f = np.zeros((3, 2, 32))
            for e in range(2):
                f[0, e] = kappa[p] * (g[0, e] * gxy[e] + g[1, e] * gy[e] + g[2, e] * gz[e])
                f[1, e] = kappa[p] * (g[1, e] * gxy[e] + g[3, e] * gy[e] + g[4, e] * gz[e])
                f[2, e] = kappa[p] * (g[2, e] * gxy[e] + g[4, e] * gy[e] + g[5, e] * gz[e])
            return f

        res = _contract_cell(P, B0, B1, B2, flux)
        dn = load_idx(pl + nw) if ahead and pl + nw < count else None
        for t in range(n):
            for e in range(2):
                add = d[t, e] >= 0
                np.add.at(y, d[t, e][add], res[t, e][add])
                row = (d[t, e] < 0) & v[e]                      # Dirichlet row: y = x
                y[~d[t, e][row]] = x[~d[t, e][row]]
        if ahead:
            d = dn


@pytest.mark.parametrize("P", [7, 6, 4, 2])
def test_lane_level_mma_apply_matches_oracle(P):
    n = P + 1
    m = om.create_box(1, 1, 1)
    A = np.array([[1.0, 0.2, 0.1], [0.0, 0.9, 0.3], [0.1, 0.0, 1.2]])   # an affine (sheared) single cell
    m.verts[:] = m.verts @ A.T
    dm, nd = om.dofmap(m, P), om.num_dofs(m, P)
    bc = np.zeros(nd, dtype=np.int8)
    G, detJ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    kap = np.array([1.7])
    x = np.random.default_rng(0).uniform(-1, 1, nd)
    yo = oo.apply(P, dm, G, kap, bc, x)
    _, w1, _ = gll.tables(P)
    w3 = (w1[:, None, None] * w1[None, :, None] * w1[None, None, :]).reshape(-1)
    Gc = G[0, 0] / w3[0]                                                  # G(q) = w_q Gc on an affine cell
    assert np.allclose(G[0] / w3[:, None], Gc[None, :], rtol=1e-12)
    acc = apply_cell(P, x[dm[0]].reshape(n, n, n), Gc, kap[0])
    y = np.zeros(nd)
    np.add.at(y, dm[0], acc.reshape(-1))
    assert np.linalg.norm(y - yo) <= 1e-13 * np.linalg.norm(yo)
    # the swizzled layout is bank-conflict free for every access of the kernel
    assert WAVEFRONTS["actual"] == WAVEFRONTS["ideal"], WAVEFRONTS


@pytest.mark.parametrize("P,ahead", [(7, True), (6, True), (6, False), (3, True)])
def test_lane_level_streamed_kernel_on_a_mesh(P, ahead):
    """The streamed-G variant as the launch runs it: a perturbed 2x3x2 box with its Dirichlet boundary, an interior
    and a boundary launch (first / count), three "warps" sharing the cells so that each loops over several cells
    (the small GPU parity meshes give every warp at most one cell), dof indices one cell ahead or not."""
    m = om.create_box(2, 3, 2, perturb=0.2)
    n3 = (P + 1) ** 3
    dm, bc, nd = om.dofmap(m, P), om.bc_marker(m, P), om.num_dofs(m, P)
    G, _ = oo.geometry_factors(m.verts, m.geom_dofmap, P)
    kap = np.random.default_rng(5).uniform(0.5, 2.0, m.ncells)
    x = np.random.default_rng(1).uniform(-1, 1, nd)
    yo = oo.apply(P, dm, G, kap, bc, x)
    enc = np.where(bc[dm] != 0, ~dm.astype(np.int64), dm.astype(np.int64)).reshape(-1)     # enc = bc[d] ? ~d : d
    Gdev = np.ascontiguousarray(np.transpose(G, (0, 2, 1))).reshape(-1)                    # G[p][6][n3]
    assert enc.size == m.ncells * n3 and Gdev.size == m.ncells * 6 * n3
    y = np.zeros(nd)
    n_l = 7                                                                               # "interior" launch, then the rest
    for first, count in ((0, n_l), (n_l, m.ncells - n_l)):
        for gw in range(3):
            warp_kernel(P, enc, Gdev, kap, x, y, first, count, gw, 3, ahead)
    assert np.linalg.norm(y - yo) <= 1e-12 * np.linalg.norm(yo)
