"""Structured hex box mesh, dofmaps, BC markers and ghost-layer partition (oracle).

TEST INFRASTRUCTURE ONLY.  Restates, for a unit-cube box, the arrays the
reference obtains from DOLFINx (absent here):
  * ``mesh::create_box`` + tensor-product geometry dofmap  examples/pmg/main.cpp:442-451,
                                                           src/mesh.hpp:75-84 (z fastest)
  * ``V->dofmap()->map()`` for tp GLL spaces               examples/pmg/main.cpp:83-87,205-213
  * Dirichlet marker on all exterior facets                examples/pmg/main.cpp:122-124,173-185
  * ghost layer + lcells/bcells split                      src/mesh.hpp:16-143
  * owned-first / ghosts-after vector layout and the forward scatter lists
                                                           src/vector.hpp:86-95,186-238
DOLFINx's global numbering cannot be reproduced; parity is stated in the
canonical lexicographic numbering gid = (gx*Ny + gy)*Nz + gz on the GLL grid.
"""
from dataclasses import dataclass, field
import numpy as np

from . import gll


@dataclass
class BoxMesh:
    n: tuple                      # cells per direction (nx, ny, nz)
    verts: np.ndarray             # [nv, 3] float64
    geom_dofmap: np.ndarray       # [nc, 8] int32, local vertex = (a*2+b)*2+c

    @property
    def ncells(self):
        return int(np.prod(self.n))


def create_box(nx, ny, nz, perturb=0.0, seed=1234):
    """Unit cube, vertices exactly i/n; optional interior displacement
    U(-perturb*h, perturb*h)^3 with default_rng(seed) (parity only)."""
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    verts = np.stack([gx / nx, gy / ny, gz / nz], axis=-1).reshape(-1, 3).astype(np.float64)
    if perturb > 0.0:
        rng = np.random.default_rng(seed)
        h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
        d = rng.uniform(-perturb, perturb, size=verts.shape) * h
        interior = ((gx > 0) & (gx < nx) & (gy > 0) & (gy < ny) & (gz > 0) & (gz < nz)).reshape(-1)
        verts[interior] += d[interior]
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)
    gd = np.empty((nx * ny * nz, 8), dtype=np.int32)
    for a in range(2):
        for b in range(2):
            for c in range(2):
                gd[:, (a * 2 + b) * 2 + c] = ((cx + a) * (ny + 1) + (cy + b)) * (nz + 1) + (cz + c)
    return BoxMesh((nx, ny, nz), verts, gd)


def dof_grid(mesh, P):
    nx, ny, nz = mesh.n
    return (P * nx + 1, P * ny + 1, P * nz + 1)


def num_dofs(mesh, P):
    return int(np.prod(dof_grid(mesh, P)))


def dofmap(mesh, P):
    """[nc, (P+1)^3] int32; local index ix*nd^2+iy*nd+iz, x slowest (src/laplacian.hpp:173)."""
    nx, ny, nz = mesh.n
    Nx, Ny, Nz = dof_grid(mesh, P)
    nd = P + 1
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ix, iy, iz = np.meshgrid(np.arange(nd), np.arange(nd), np.arange(nd), indexing="ij")
    gx = cx.reshape(-1, 1) * P + ix.reshape(1, -1)
    gy = cy.reshape(-1, 1) * P + iy.reshape(1, -1)
    gz = cz.reshape(-1, 1) * P + iz.reshape(1, -1)
    return ((gx * Ny + gy) * Nz + gz).astype(np.int32)


def bc_marker(mesh, P):
    """int8[ndofs]: 1 on every exterior-facet dof (examples/pmg/main.cpp:122-124,173-185)."""
    Nx, Ny, Nz = dof_grid(mesh, P)
    m = np.zeros((Nx, Ny, Nz), dtype=np.int8)
    m[0], m[-1] = 1, 1
    m[:, 0], m[:, -1] = 1, 1
    m[:, :, 0], m[:, :, -1] = 1, 1
    return m.reshape(-1)


def dof_coords(mesh, P):
    """Physical coordinates of every dof (trilinear push-forward of the GLL nodes)."""
    x1, _ = gll.gll_points_weights(P + 1)
    nd = P + 1
    dm = dofmap(mesh, P)
    X = np.zeros((num_dofs(mesh, P), 3))
    ix, iy, iz = np.meshgrid(x1, x1, x1, indexing="ij")
    ref = np.stack([ix.reshape(-1), iy.reshape(-1), iz.reshape(-1)], axis=-1)  # [nd^3, 3]
    phi = np.zeros((nd ** 3, 8))
    for a in range(2):
        for b in range(2):
            for c in range(2):
                la = ref[:, 0] if a else 1 - ref[:, 0]
                lb = ref[:, 1] if b else 1 - ref[:, 1]
                lc = ref[:, 2] if c else 1 - ref[:, 2]
                phi[:, (a * 2 + b) * 2 + c] = la * lb * lc
    cv = mesh.verts[mesh.geom_dofmap]            # [nc, 8, 3]
    xc = np.einsum("qk,ckd->cqd", phi, cv)       # [nc, nd^3, 3]
    X[dm.reshape(-1)] = xc.reshape(-1, 3)
    return X


# ----------------------------------------------------------------------------
# Ghost-layer partition (src/mesh.hpp:16-143, src/vector.hpp:86-95)
# ----------------------------------------------------------------------------
def _splits(n, p):
    return [(i * n) // p for i in range(p + 1)]


def cell_owner(mesh, pgrid):
    """Block partition of the box over px*py*pz ranks; rank = (bx*py+by)*pz+bz."""
    nx, ny, nz = mesh.n
    px, py, pz = pgrid
    def blk(n, p):
        s = _splits(n, p)
        o = np.zeros(n, dtype=np.int32)
        for b in range(p):
            o[s[b]:s[b + 1]] = b
        return o
    bx, by, bz = blk(nx, px), blk(ny, py), blk(nz, pz)
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    return ((bx[cx] * py + by[cy]) * pz + bz[cz]).reshape(-1).astype(np.int32)


@dataclass
class LocalPart:
    rank: int
    cells: np.ndarray             # global cell ids, owned first then ghost
    n_owned_cells: int
    verts: np.ndarray             # [nlv, 3] local vertex coordinates
    geom_dofmap: np.ndarray       # [nlc, 8] local vertex ids
    lcells: np.ndarray            # local cell ids touching owned dofs only   (mesh.hpp:119-138)
    bcells: np.ndarray            # the rest + all ghost cells
    levels: dict = field(default_factory=dict)   # P -> LocalLevel


@dataclass
class LocalLevel:
    P: int
    dofmap: np.ndarray            # [nlc, nd^3] local dof ids
    l2g: np.ndarray               # local -> global dof id, owned first
    n_owned: int
    n_ghost: int
    bc: np.ndarray                # int8 [n_owned+n_ghost]
    nbr_send: list                # [(rank, local owned idx array)]  Scatterer local_indices
    nbr_recv: list                # [(rank, ghost slot array (0-based within ghost block))]


def partition(mesh, pgrid, degrees):
    """Return [LocalPart per rank].  Dof ownership rule (documented choice, DOLFINx's
    cannot be reproduced): the lowest rank owning a cell that contains the dof."""
    R = int(np.prod(pgrid))
    owner = cell_owner(mesh, pgrid)
    nv = mesh.verts.shape[0]
    # vertex -> set of ranks owning a touching cell (as bitmask, R <= 64)
    vmask = np.zeros(nv, dtype=np.uint64)
    np.bitwise_or.at(vmask, mesh.geom_dofmap.reshape(-1), np.repeat(np.uint64(1) << owner.astype(np.uint64), 8))
    parts = []
    dm = {P: dofmap(mesh, P) for P in degrees}
    bc = {P: bc_marker(mesh, P) for P in degrees}
    dof_owner = {}
    for P in degrees:
        o = np.full(num_dofs(mesh, P), R, dtype=np.int32)
        np.minimum.at(o, dm[P].reshape(-1), np.repeat(owner, dm[P].shape[1]))
        dof_owner[P] = o
    for r in range(R):
        owned = np.where(owner == r)[0]
        bit = np.uint64(1) << np.uint64(r)
        touches = ((vmask[mesh.geom_dofmap] & bit) != 0).any(axis=1)
        ghost = np.where(touches & (owner != r))[0]
        cells = np.concatenate([owned, ghost]).astype(np.int64)
        gv = mesh.geom_dofmap[cells]
        uv, inv = np.unique(gv.reshape(-1), return_inverse=True)
        part = LocalPart(r, cells, len(owned), mesh.verts[uv].copy(),
                         inv.reshape(gv.shape).astype(np.int32), None, None)
        for P in degrees:
            gd = dm[P][cells]
            ud = np.unique(gd.reshape(-1))
            own_mask = dof_owner[P][ud] == r
            owned_d = ud[own_mask]
            ghost_d = ud[~own_mask]
            order = np.lexsort((ghost_d, dof_owner[P][ghost_d]))
            ghost_d = ghost_d[order]
            l2g = np.concatenate([owned_d, ghost_d])
            g2l = -np.ones(num_dofs(mesh, P), dtype=np.int64)
            g2l[l2g] = np.arange(len(l2g))
            ldm = g2l[gd].astype(np.int32)
            part.levels[P] = LocalLevel(P, ldm, l2g, len(owned_d), len(ghost_d), bc[P][l2g].copy(), [], [])
        parts.append(part)
    # lcells/bcells from the finest space, reused on all levels (examples/pmg/main.cpp:95-97)
    Pf = max(degrees)
    for part in parts:
        lv = part.levels[Pf]
        has_ghost = (lv.dofmap >= lv.n_owned).any(axis=1)
        has_ghost[part.n_owned_cells:] = True
        part.lcells = np.where(~has_ghost)[0].astype(np.int32)
        part.bcells = np.where(has_ghost)[0].astype(np.int32)
    # forward-scatter lists
    for P in degrees:
        for part in parts:
            lv = part.levels[P]
            gids = lv.l2g[lv.n_owned:]
            own = dof_owner[P][gids]
            for q in np.unique(own):
                slots = np.where(own == q)[0]
                lv.nbr_recv.append((int(q), slots.astype(np.int32)))
                src = parts[int(q)].levels[P]
                pos = np.searchsorted(src.l2g[:src.n_owned], gids[slots])
                assert (src.l2g[pos] == gids[slots]).all()
                src.nbr_send.append((part.rank, pos.astype(np.int32)))
    return parts


def scatter_fwd(parts, P, vecs):
    """Owner -> ghost update of a list of local vectors (src/vector.hpp:186-238)."""
    for part, v in zip(parts, vecs):
        lv = part.levels[P]
        for (q, slots) in lv.nbr_recv:
            src = parts[q].levels[P]
            for (dst, idx) in src.nbr_send:
                if dst == part.rank:
                    v[lv.n_owned + slots] = vecs[q][idx]
